"""`--bounded <line_size>` (reference: cache_fix.rs, train_bounded in train/mod.rs:156-184, the
cache-fix parts of codegen.rs): the host pre-pass host/cache_fix.hpp against the Python
restatement oracle/cache_fix.py knot for knot, and the whole flow — knots -> two-layer RMI over
the knots -> generated code with the spline lookup — against the reference's own property
(tests/cache_fix_wiki/main.cpp: every key's lookup lands within one line of its lower bound).
CPU-only: the RMI over the knots comes from the oracle; the GPU test at the bottom runs the
`rmi` CLI with --bounded end to end."""
import os
import subprocess

import numpy as np
import pytest

from tests import datasets
from tests.test_codegen import build_and_check, dump_model, tool, write_keyfile  # noqa: F401  (tool is a fixture)


def host_cache_fix(tool_exe, work, keys, line):
    keyfile = os.path.join(work, "keys_cf.bin")
    write_keyfile(keyfile, keys)
    out = os.path.join(work, "spline.bin")
    r = subprocess.run([tool_exe, "cachefix", keyfile, str(line), out], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    raw = np.fromfile(out, dtype="<u8").reshape(-1, 2)
    return raw, out


DATA = {
    "uniform": lambda: datasets.uniform_u64(6_000, seed=5),
    "dups": lambda: datasets.with_duplicates(datasets.uniform_u64(6_000, seed=6), frac=0.3),
    "lognormal": lambda: datasets.lognormal_u64(6_000, seed=7),
    "dense": lambda: np.arange(10, 1_510, dtype=np.uint64),                        # consecutive keys: key-1 == last key
    "steps": lambda: np.sort(np.repeat(np.arange(1, 601, dtype=np.uint64) * 1000, 10)),   # long runs of equal keys
}


@pytest.mark.parametrize("dname", list(DATA))
@pytest.mark.parametrize("line", [1, 8, 64])
def test_host_cache_fix_equals_python_restatement(tool, tmp_path, dname, line):
    from oracle import cache_fix as ocf
    keys = DATA[dname]()
    want = ocf.cache_fix(keys.tolist(), line)
    got, _ = host_cache_fix(tool, str(tmp_path), keys, line)
    assert got.shape[0] == len(want)
    assert [tuple(map(int, r)) for r in got] == want
    # knots: strictly increasing keys, non-decreasing offsets, first knot = (first key - 1 or first key, 0)
    assert np.all(np.diff(got[:, 0].astype(object)) > 0) and np.all(np.diff(got[:, 1].astype(np.int64)) >= 0)
    assert int(got[-1, 0]) == int(keys[-1])


def test_cache_fix_panics_like_the_reference(tool, tmp_path):
    from oracle import cache_fix as ocf
    few = np.arange(1, 6, dtype=np.uint64)
    with pytest.raises(ocf.CacheFixPanic):
        ocf.cache_fix(few.tolist(), 8)              # fewer items than the line size (cache_fix.rs:107-108)
    with pytest.raises(RuntimeError, match="fewer items"):
        host_cache_fix(tool, str(tmp_path), few, 8)
    zero = np.arange(0, 100, dtype=np.uint64)       # key 0: key - 1 wraps, with_new_dest's assert fires
    with pytest.raises(ocf.CacheFixPanic):
        ocf.cache_fix(zero.tolist(), 4)
    with pytest.raises(RuntimeError):
        host_cache_fix(tool, str(tmp_path), zero, 4)


@pytest.mark.parametrize("spec,bf,dname,line", [("linear_spline,linear", 64, "uniform", 8), ("cubic,linear", 32, "dups", 8),
                                                ("radix,linear", 128, "uniform", 4), ("linear,linear", 64, "uniform", 16)])
def test_bounded_rmi_holds_the_reference_property(oracle, tool, tmp_path, spec, bf, dname, line):
    keys = DATA[dname]() if dname != "uniform" else datasets.uniform_u64(100_000, seed=8)
    work = str(tmp_path)
    os.makedirs(os.path.join(work, "rmi_data"))
    knots, spline_path = host_cache_fix(tool, work, keys, line)
    knot_keys = np.ascontiguousarray(knots[:, 0])
    try:
        o = oracle.train(knot_keys, spec, bf)       # train_bounded: the RMI indexes the knots (train/mod.rs:165-174)
    except oracle.OraclePanic as e:
        pytest.skip(f"reference panics: {e}")
    dump = os.path.join(work, "model.bin")
    dump_model(dump, o, spec)
    subprocess.run([tool, dump, "rmi", os.path.join(work, "rmi_data"), work, "1", "0", spline_path, str(line), str(keys.size)],
                   check=True)
    # artefacts: the knots are one more parameter array, byte for byte {key, offset} pairs
    blob = open(os.path.join(work, "rmi_data", "rmi_L2_PARAMETERS"), "rb").read()
    assert blob == knots.astype("<u8").tobytes()
    code = open(os.path.join(work, "rmi.cpp")).read()
    assert "#include <algorithm>" in code and "uint64_t _rmi_lookup_pre_cachefix(uint64_t key, size_t* err)" in code
    assert f"const uint64_t num_spline_pts = {knots.shape[0]};" in code and f"const uint64_t total_keys = {keys.size};" in code
    assert f"  *err = {line};" in code and f"/ {line}) * {line};" in code
    hdr = open(os.path.join(work, "rmi.h")).read()
    assert "uint64_t lookup(uint64_t key, size_t* err);" in hdr and "_rmi_lookup_pre_cachefix" not in hdr
    ppm = o.l1_params.shape[1]
    top_bytes = 8 * (len(o.l0.fp) + (len(o.l0.ip) if o.l0.kind in ("radix", "bradix") else 0))
    assert f"const size_t RMI_SIZE = {top_bytes + o.branching_factor * (8 * ppm + 8) + 16 * knots.shape[0]};" in hdr
    # the reference's cache-fix test: |lookup(key) - lower_bound(key)| <= line size, for every key
    out = build_and_check(work, keys)
    assert out.startswith("ok")
    assert int(out.split()[2]) <= line


@pytest.mark.gpu
def test_cli_bounded_end_to_end_on_gpu(tmp_path):
    """`rmi <file> rmi linear_spline,linear 4096 --bounded 8` (tests/cache_fix_wiki/Makefile:8 at small scale)."""
    from rmi_b200 import build
    cli = build.build_cli()
    keys = datasets.uniform_u64(1_000_000, seed=9)
    keys = keys[keys > 0]
    work = str(tmp_path)
    datafile = os.path.join(work, "synthetic_1M_uint64")
    write_keyfile(datafile, keys)
    r = subprocess.run([cli, datafile, "rmi", "linear_spline,linear", "4096", "--bounded", "8", "--zero-build-time"], cwd=work,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = build_and_check(work, keys)
    assert out.startswith("ok") and int(out.split()[2]) <= 8


@pytest.mark.parametrize("seed", range(12))
def test_host_cache_fix_randomised_against_python(tool, tmp_path, seed):
    """Small random key sets of very different shapes (clusters, huge gaps, keys up to 2^64 - 1, long runs of
    equal keys, consecutive integers) and line sizes: the host scan and the Python restatement must emit the
    same knots, or both must panic."""
    from oracle import cache_fix as ocf
    rng = np.random.Generator(np.random.MT19937(1000 + seed))
    n = int(rng.integers(40, 900))
    shape = seed % 6
    if shape == 0:
        keys = rng.integers(1, 1 << 63, size=n, dtype=np.uint64)
    elif shape == 1:      # clusters separated by huge gaps
        centres = rng.integers(1 << 20, 1 << 62, size=5, dtype=np.uint64)
        keys = (centres[rng.integers(0, 5, size=n)] + rng.integers(0, 1000, size=n, dtype=np.uint64)).astype(np.uint64)
    elif shape == 2:      # up to the top of the u64 range
        keys = (np.uint64((1 << 64) - 1) - rng.integers(0, 1 << 40, size=n, dtype=np.uint64)).astype(np.uint64)
    elif shape == 3:      # long runs of equal keys
        keys = np.repeat(rng.integers(1, 1 << 50, size=max(n // 20, 3), dtype=np.uint64), 20)
    elif shape == 4:      # consecutive integers with a few holes
        keys = np.arange(5, 5 + n, dtype=np.uint64)
        keys = np.delete(keys, rng.integers(0, n, size=n // 10))
    else:                 # tiny keys incl. 1 (key - 1 == 0 == the initial last_key)
        keys = rng.integers(1, 4 * n, size=n, dtype=np.uint64)
    keys = np.sort(keys)
    line = int(rng.choice([1, 2, 3, 8, 16, 37]))
    try:
        want = ocf.cache_fix(keys.tolist(), line)
    except ocf.CacheFixPanic:
        with pytest.raises(RuntimeError):
            host_cache_fix(tool, str(tmp_path), keys, line)
        return
    got, _ = host_cache_fix(tool, str(tmp_path), keys, line)
    assert [tuple(map(int, r)) for r in got] == want
    # the property the spline exists for: interpolating between the knots puts every key in its own line
    kk, vv = got[:, 0].astype(object), got[:, 1].astype(object)
    firsts = {}
    for i, k in enumerate(keys.tolist()):
        firsts.setdefault(k, i)
    for k, off in list(firsts.items())[:: max(1, len(firsts) // 200)]:
        j = int(np.searchsorted(got[:, 0], np.uint64(k), side="left"))
        if j == 0 or int(kk[j]) == k and j == 0:
            continue
        x0, y0, x1, y1 = int(kk[j - 1]), int(vv[j - 1]), int(kk[j]), int(vv[j])
        t = float(k - x0) / float(x1 - x0)
        pred = int(ocf._fma(1.0 - t, float(y0), t * float(y1)))
        assert pred // line == off // line, (k, off, pred, line)


def test_library_cache_fix_through_the_c_abi():
    """rmi_cache_fix (include/rmi_b200.h) is host-only: it runs without a GPU and must equal the Python restatement."""
    import rmi_b200
    from oracle import cache_fix as ocf
    keys = datasets.with_duplicates(datasets.uniform_u64(5_000, seed=31), frac=0.2)
    keys = keys[keys > 0]
    for line in (1, 8, 32):
        got = rmi_b200.cache_fix(keys, line)
        assert got.dtype == np.uint64 and got.shape[1] == 2
        assert [tuple(map(int, r)) for r in got] == ocf.cache_fix(keys.tolist(), line)
    with pytest.raises(rmi_b200.RMIPanic, match="fewer items"):
        rmi_b200.cache_fix(keys[:4], 8)
    with pytest.raises(rmi_b200.RMIPanic, match="u64"):
        rmi_b200.cache_fix(keys.astype(np.float64), 8)
