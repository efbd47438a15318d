"""The range-partitioned build's host logic (rmi_b200/sharded.py: layout planning, the three
collectives, halo planning + exchange, ownership) under torch.distributed/gloo with
world_size 2 and 3 on CPU.  The arithmetic engine is tests/shard_engine_numpy.py; the result
on every rank must equal the oracle's single-process build of the concatenated keys."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import datasets, parity


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_keys(kind, n):
    if kind == "uniform":
        return datasets.uniform_u64(n, seed=21)
    if kind == "dups":
        k = datasets.with_duplicates(datasets.uniform_u64(n, seed=22), frac=0.2)
        # a long run of equal keys that straddles every cut of an even split
        k[n // 2 - 40: n // 2 + 40] = k[n // 2 - 40]
        k[n // 3 - 5: n // 3 + 5] = k[n // 3 - 5]
        k.sort()
        return k
    return datasets.lognormal_u64(n, seed=23)


def _cuts(n, world, uneven):
    if not uneven:
        return [n * r // world for r in range(world + 1)]
    w = np.array([1.0 + 0.7 * r for r in range(world)])
    c = [0] + [int(x) for x in np.cumsum(w / w.sum() * n)]
    c[-1] = n
    return c


def _worker(rank, world, port, kind, n, spec, N, uneven, out_q, halo=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from rmi_b200 import sharded
        from tests.shard_engine_numpy import NumpyShardedData
        keys = _make_keys(kind, n)
        c = _cuts(n, world, uneven)
        data = NumpyShardedData(keys[c[rank]:c[rank + 1]].copy(), halo_capacity=n if halo is None else halo)
        g = sharded.train_sharded(data, spec, N)
        top = spec.split(",")[0]
        if top in ("linear", "robust_linear", "normal", "lognormal", "cubic"):
            # order-dependent sums (and libm's pow for cubic): coefficients within tolerance, and
            # with the same coefficients everything downstream bit for bit
            o_ref = oracle.train(keys, spec, N)
            parity.assert_top_equal(g, o_ref, exact=False, N=N)
            o = oracle.train(keys, spec, N, l0_override=g.l0_fparams)
        else:
            o = oracle.train(keys, spec, N)
            g.l0_model = o.l0.kind
            parity.assert_top_equal(g, o, exact=True)
        parity.assert_leaves_equal(g, o)
        assert g.model_max_error == o.max_error and g.model_max_error_idx == o.max_error_idx
        assert g.model_avg_error == o.avg_error
        # every rank must hold the same result
        t = torch.from_numpy(np.ascontiguousarray(g.l1_params).view(np.int64).reshape(-1).copy())
        ref = t.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(t, ref)
        out_q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        out_q.put((rank, "FAIL: " + "".join(traceback.format_exception(e))[-1500:]))
    finally:
        dist.destroy_process_group()


CASES = [
    (2, "uniform", "radix,linear", 64, False),
    (2, "uniform", "linear,linear", 64, True),
    (2, "dups", "linear_spline,linear", 48, False),
    (3, "dups", "radix,linear", 100, True),
    (3, "lognormal", "linear_spline,linear_spline", 64, False),
    (2, "uniform", "radix,cubic", 32, True),
    (2, "lognormal", "robust_linear,linear", 16, False),
    (2, "uniform", "cubic,linear", 64, True),
    (3, "dups", "cubic,linear", 48, False),
    (3, "lognormal", "cubic,linear_spline", 32, True),
    (2, "uniform", "normal,linear", 32, False),
    (3, "lognormal", "lognormal,linear", 32, True),
]


@pytest.mark.parametrize("world,kind,spec,N,uneven", CASES, ids=[f"w{c[0]}-{c[1]}-{c[2]}-{c[3]}" for c in CASES])
def test_sharded_build_equals_single_process_build(oracle, world, kind, spec, N, uneven):
    n = 6000
    try:
        oracle.train(_make_keys(kind, n), spec, N)
    except oracle.OraclePanic as e:
        pytest.skip(f"reference panics on this configuration: {e}")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, n, spec, N, uneven, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    bad = [r for r in results if r[1] != "ok"]
    assert not bad, bad


def test_halo_grows_when_a_leaf_reaches_past_the_prefetched_keys(oracle):
    """Skewed data, a deliberately tiny halo: the first build reports ST_HALO_TOO_SMALL on every rank, the
    orchestrator sizes the halo from the global boundaries, re-homes the slab and builds again."""
    world, kind, n, spec, N = 3, "lognormal", 6000, "linear_spline,linear", 24
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, n, spec, N, False, q, 4)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert not [r for r in results if r[1] != "ok"], results


def test_layout_planner_handles_runs_spanning_ranks():
    from rmi_b200 import sharded
    # rank 1 consists of one repeated key that began on rank 0 and continues into rank 2
    ends = np.array([[5, 9, 3, 10], [9, 9, 0, 4], [9, 12, 2, 6]], dtype=np.uint64)
    lay = sharded.plan_global_layout(ends, 0, 64)
    assert [d["base"] for d in lay] == [0, 10, 14]
    assert lay[1]["prev_F"] == 3 and lay[2]["prev_F"] == 3 and lay[2]["prev_key_bits"] == 9
    assert lay[0]["last_F"] == 16 and lay[2]["is_last"] == 1 and lay[0]["has_prev"] == 0
    # empty rank in the middle
    ends = np.array([[1, 4, 2, 3], [0, 0, 0, 0], [7, 8, 1, 2]], dtype=np.uint64)
    lay = sharded.plan_global_layout(ends, 0, 8)
    assert lay[2]["prev_key_bits"] == 4 and lay[2]["prev_F"] == 2 and lay[1]["has_prev"] == 1
    assert lay[2]["base"] == 3 and lay[2]["n_global"] == 5


def test_halo_planner():
    from rmi_b200 import sharded
    # rank 0's last leaf ends at 13 (inside rank 2): it needs [10, 14) = 2 keys of rank 1 + 2 of rank 2... rank 1 has 2 keys
    moves = sharded.plan_halo([0, 10, 12, 20], [13, 15, 20], 20)
    assert moves == [(0, 1, 0, 2), (0, 2, 0, 2), (1, 2, 0, 4)]
    assert sharded.plan_halo([0, 10], [10], 10) == []


def test_slab_reader_partitions_a_key_file(tmp_path):
    """read_slab: every rank reads only its contiguous slab; the slabs tile the file in order."""
    import struct
    from rmi_b200 import api, sharded
    for suffix, keys in (("uint64", datasets.uniform_u64(10_007, seed=61)), ("uint32", datasets.uniform_u32(5_003, seed=62)),
                         ("f64", datasets.uniform_f64(4_001, seed=63))):
        path = str(tmp_path / f"keys_{suffix}")
        with open(path, "wb") as f:
            f.write(struct.pack("<Q", keys.size))
            f.write(keys.tobytes())
        for world in (1, 2, 3, 8):
            parts = [sharded.read_slab(path, r, world) for r in range(world)]
            assert all(n == keys.size for _, n in parts)
            assert [p.size for p, _ in parts] == [b - a for a, b in (sharded.slab_bounds(keys.size, r, world) for r in range(world))]
            assert np.array_equal(np.concatenate([p for p, _ in parts]), keys)
            assert parts[0][0].dtype == keys.dtype
    assert sharded.key_type_of_path("/x/wiki_ts_200M_uint64") == api.KEY_U64
    with pytest.raises(api.RMIPanic):
        sharded.key_type_of_path("/x/keys.bin")
    short = str(tmp_path / "short_uint64")
    with open(short, "wb") as f:
        f.write(struct.pack("<Q", 100))
        f.write(b"\0" * 80)
    with pytest.raises(api.RMIPanic):
        sharded.read_slab(short, 0, 1)
