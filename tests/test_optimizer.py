"""The configuration search (host/optimizer.hpp, a restatement of rmi_lib/src/optimizer.rs): its
grids and Pareto helpers against a Python restatement on CPU, and `rmi --optimize` end to end on
the GPU — one replica and two replicas (worker threads building concurrently) must agree."""
import json
import os
import random
import subprocess

import pytest

from tests import datasets
from tests.test_codegen import ROOT, write_keyfile


@pytest.fixture(scope="module")
def opt_tool(tmp_path_factory):
    d = tmp_path_factory.mktemp("opt")
    exe = str(d / "optimizer_tool")
    subprocess.run(["g++", "-std=c++17", "-O1", "-pthread", os.path.join(ROOT, "tests", "cxx", "optimizer_tool.cpp"), "-o", exe], check=True)
    return exe


# ---- Python restatement of optimizer.rs ------------------------------------------------------
def top_only(profile):      # optimizer.rs:15-28
    return {"fast": ["robust_linear"], "disk": ["radix", "radix18", "radix22", "robust_linear", "normal", "lognormal", "loglinear"]
            }.get(profile, ["radix", "radix18", "radix22", "robust_linear"])


def anywhere(profile):      # :30-41
    return ["linear", "cubic"] if profile == "fast" else ["linear", "cubic", "linear_spline"]


def branching(profile):     # :43-57
    return [1 << i for i in range(6, 28 if profile == "disk" else 25, 2 if profile == "fast" else 1)]


def first_phase(profile):   # :110-125
    out = []
    for t in top_only(profile) + anywhere(profile):
        for b in anywhere(profile):
            out += [(f"{t},{b}", bf) for bf in branching(profile)[::5]]
    return out


def dominated_by(a, b):     # :173-187  (a, b) = (models, bf, avg, mx, size)
    if a[4] < b[4] or a[2] < b[2]:
        return False
    if a[4] == b[4] and a[2] <= b[2]:
        return False
    if a[4] <= b[4] and abs(a[2] - b[2]) < 2.220446049250313e-16:
        return False
    return True


def pareto(rs):             # :59-72
    return [x for x in rs if not any(dominated_by(x, v) for v in rs)]


def narrow(rs, desired):    # :74-108
    if len(rs) <= desired:
        return list(rs)
    tmp = sorted(rs, key=lambda r: r[4])
    best, tmp = tmp[0], tmp[1:]
    while len(tmp) > desired - 1:
        ratios = [tmp[i + 1][4] / tmp[i][4] for i in range(len(tmp) - 1)]
        gi = ratios.index(min(ratios))
        del tmp[gi if tmp[gi][2] > tmp[gi + 1][2] else gi + 1]
    return [best] + tmp


def second_phase(first, profile):   # :127-151
    out = []
    for m in sorted({r[0] for r in pareto(first)}):
        out += [(m, bf) for bf in branching(profile) if not any(v[0] == m and v[1] == bf for v in first)]
    return out


def run(tool, args, stats=None, profile=""):
    env = dict(os.environ, RMI_OPTIMIZER_PROFILE=profile)
    if not profile:
        env.pop("RMI_OPTIMIZER_PROFILE")
    text = "" if stats is None else "".join(f"{m} {bf} {a!r} {x!r} {s}\n" for m, bf, a, x, s in stats)
    r = subprocess.run([tool] + args, input=text, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    return [ln.split() for ln in r.stdout.splitlines()]


@pytest.mark.parametrize("profile", ["", "fast", "memory", "disk"])
def test_search_grids(opt_tool, profile):
    got = [(m, int(b)) for m, b in run(opt_tool, ["first"], profile=profile)]
    assert got == first_phase(profile)
    assert len(first_phase("")) == 84        # 84 phase-1 configurations by default
    rng = random.Random(3)
    stats = [(m, bf, rng.uniform(1, 12), rng.uniform(5, 20), bf * rng.choice([16, 24, 40])) for m, bf in first_phase(profile)]
    got2 = [(m, int(b)) for m, b in run(opt_tool, ["second"], stats, profile=profile)]
    assert got2 == second_phase(stats, profile)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_pareto_and_narrowing(opt_tool, seed):
    rng = random.Random(seed)
    stats = []
    for i in range(60):
        size = rng.choice([1 << k for k in range(10, 24)]) * rng.choice([16, 24, 40])
        avg = round(rng.uniform(0.5, 14.0), rng.choice([1, 3, 9]))      # ties in the error on purpose
        stats.append((f"m{i % 7},l{i % 3}", 1 << (6 + i % 15), avg, avg + 2.0, size))
    for restrict in (2, 5, 10, 1000):
        want = sorted(narrow(pareto(stats), restrict), key=lambda r: r[2])
        got = run(opt_tool, ["front", str(restrict)], stats)
        assert [(g[0], int(g[1]), int(g[2])) for g in got] == [(w[0], w[1], w[4]) for w in want]


@pytest.mark.gpu
def test_cli_optimize_on_gpu_one_and_two_replicas(tmp_path):
    from rmi_b200 import build
    cli = build.build_cli()
    keys = datasets.uniform_u64(400_000, seed=12)
    work = str(tmp_path)
    datafile = os.path.join(work, "synthetic_400k_uint64")
    write_keyfile(datafile, keys)
    env = dict(os.environ, RMI_OPTIMIZER_PROFILE="fast")
    outs = []
    for name, extra in (("one.json", []), ("two.json", ["--devices", "0,0"])):
        r = subprocess.run([cli, datafile, "--optimize", name] + extra, cwd=work, capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        cfgs = json.load(open(os.path.join(work, name)))["configs"]
        assert 2 <= len(cfgs) <= 10
        errs = [c["average log2 error"] for c in cfgs]
        assert errs == sorted(errs)
        # a Pareto front: sorted by error ascending means sizes descending
        sizes = [c["size"] for c in cfgs]
        assert all(a > b for a, b in zip(sizes, sizes[1:]))
        assert all(c["namespace"] == f"synthetic_400k_uint64_{i}" and c["binary"] is True for i, c in enumerate(cfgs))
        outs.append(cfgs)
        assert "Models" in r.stdout and "AvgLg2" in r.stdout
    assert outs[0] == outs[1]


@pytest.mark.gpu
def test_python_mirrors_of_the_public_api_on_gpu(tmp_path, monkeypatch):
    """find_pareto_efficient_configs / train_for_size / train_bounded / output_rmi through rmi_b200 (the same
    library entry points the CLI uses)."""
    import rmi_b200
    from tests.test_codegen import build_and_check
    monkeypatch.setenv("RMI_OPTIMIZER_PROFILE", "fast")
    keys = datasets.uniform_u64(400_000, seed=12)
    keys = keys[keys > 0]
    ds = rmi_b200.RMITrainingData(keys)
    front = rmi_b200.find_pareto_efficient_configs(ds, 10)
    assert 2 <= len(front) <= 10
    errs = [c["average_log2_error"] for c in front]
    sizes = [c["size"] for c in front]
    assert errs == sorted(errs) and all(a > b for a, b in zip(sizes, sizes[1:]))
    two = rmi_b200.find_pareto_efficient_configs([ds, ds.replicate(0)], 10)
    assert two == front
    # train_for_size: the most accurate configuration below the bound
    bound = sizes[len(sizes) // 2] + 1
    m = rmi_b200.train_for_size(ds, bound)
    assert rmi_b200.rmi_size(m) < bound
    with pytest.raises(rmi_b200.RMIPanic, match="smaller than"):
        rmi_b200.train_for_size(ds, 8)
    # train_bounded + output_rmi: the reference's cache-fix property on the generated code
    rmi, knots = rmi_b200.train_bounded(keys, "linear_spline,linear", 1024, 8)
    assert rmi.num_data_rows == keys.size and rmi.num_rmi_rows == knots.shape[0]
    work = str(tmp_path)
    rmi_b200.output_rmi("rmi", rmi, os.path.join(work, "rmi_data"), out_dir=work, build_time_ns=0, cache_fix_knots=knots,
                        line_size=8, num_data_rows=keys.size)
    out = build_and_check(work, keys)
    assert out.startswith("ok") and int(out.split()[2]) <= 8
    assert f"const size_t RMI_SIZE = {rmi_b200.rmi_size(rmi, num_spline_points=knots.shape[0])};" in open(os.path.join(work, "rmi.h")).read()


@pytest.mark.gpu
@pytest.mark.parametrize("top,bf", [("radix", 1024), ("radix18", 4096), ("robust_linear", 256), ("cubic", 1000), ("bradix", 512),
                                    ("histogram", 128)])
def test_stats_batch_equals_separate_builds(top, bf):
    """rmi_train_stats_batch (one top fit + one boundary pass for every leaf type of a (top, branching factor) group —
    the optimizer's unit of work) must report exactly what separate stats-only builds report, and the oracle's numbers."""
    import oracle
    import rmi_b200
    oracle.build()
    keys = datasets.with_duplicates(datasets.lognormal_u64(300_000, seed=21))
    ds = rmi_b200.RMITrainingData(keys)
    leaves = ["linear", "cubic", "linear_spline"]
    try:
        batch = rmi_b200.train_stats_batch(ds, top, leaves, bf)
    except rmi_b200.RMIPanic:
        with pytest.raises(rmi_b200.RMIPanic):
            for leaf in leaves:
                rmi_b200.train(ds, f"{top},{leaf}", bf, rmi_b200.FLAG_STATS_ONLY, counts=False)
        return
    for leaf, b in zip(leaves, batch):
        one = rmi_b200.train(ds, f"{top},{leaf}", bf, rmi_b200.FLAG_STATS_ONLY, counts=False)
        for f in ("model_max_error", "model_max_error_idx", "model_avg_error", "model_avg_l2_error", "model_avg_log2_error",
                  "model_max_log2_error", "branching_factor", "l0_model", "l1_model"):
            assert getattr(b, f) == getattr(one, f), (leaf, f, getattr(b, f), getattr(one, f))
        assert list(b.l0_fparams) == list(one.l0_fparams) and list(b.l0_iparams) == list(one.l0_iparams)
        assert rmi_b200.rmi_size(b) == rmi_b200.rmi_size(one)
        if top in ("radix", "radix18", "bradix", "histogram"):      # tops without an order-dependent float fit: the oracle's exact numbers
            o = oracle.train(keys, f"{top},{leaf}", bf)
            if leaf != "cubic":
                assert (b.model_max_error, b.model_avg_error) == (o.max_error, o.avg_error)
