"""Committed snapshots of the oracle's builds (tests/golden/oracle_builds.json, made by
tools/make_golden.py) against (a) the oracle as built on THIS box — guards the checker against
drift from another compiler / libm — and (b) the CUDA path on the GPU box.

The snapshots are oracle outputs, not outputs of the reference binary (no Rust toolchain in
this environment); what ties the oracle to the reference is tests/test_oracle_kats.py and
tests/test_oracle_property.py."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests import parity
from tools.make_golden import DATA

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_builds.json")))["cases"]
IDS = [f"{c['spec']}:{c['branching_factor']}:{c['data']}" for c in GOLD]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def unhex(xs):
    return np.array([float.fromhex(x) for x in xs], dtype=np.float64)


def check(case, kind, fp, ip, high, t32, pivots, l1_params, l1_errors, l1_counts, max_error, max_error_idx, avg_error,
          avg_l2, avg_log2, exact):
    N = case["branching_factor"]
    assert kind == case["l0_kind"]
    want_fp = unhex(case["l0_fparams"])
    if exact:
        assert np.array_equal(parity.bits(fp), parity.bits(want_fp)), (fp, want_fp)
    elif len(want_fp):
        parity.assert_coef_close(kind, fp, want_fp, N)
    assert [int(x) for x in ip] == case["l0_iparams"]
    if kind == "bradix":
        assert bool(high) == case["l0_bradix_high"]
    if case["l0_table_sha256"]:
        assert sha(np.asarray(t32, dtype=np.uint32)) == case["l0_table_sha256"]
    if case["l0_pivots_sha256"]:
        assert sha(np.asarray(pivots, dtype=np.uint64)) == case["l0_pivots_sha256"]
    head = unhex(case["l1_params_head"]).reshape(-1, l1_params.shape[1])
    if exact:
        assert sha(np.asarray(l1_params, dtype=np.float64)) == case["l1_params_sha256"], "leaf parameters differ from the snapshot"
        assert sha(np.asarray(l1_errors, dtype=np.uint64)) == case["l1_errors_sha256"], "leaf error bounds differ"
        assert int(max_error) == case["max_error"] and int(max_error_idx) == case["max_error_idx"]
        assert float(avg_error) == float.fromhex(case["avg_error"])
        for got, want in ((avg_l2, case["avg_l2_error"]), (avg_log2, case["avg_log2_error"])):
            w = float.fromhex(want)
            assert abs(float(got) - w) <= parity.STAT_RTOL * max(abs(w), 1e-300)
        assert sha(np.asarray(l1_counts, dtype=np.uint64)) == case["l1_counts_sha256"], "leaf key counts differ"
    elif np.array_equal(parity.bits(fp), parity.bits(want_fp)):
        # libm on the path (pow / ln): same top model -> same leaf ranges; leaf parameters within tolerance
        got = np.asarray(l1_params[: head.shape[0]], dtype=np.float64)
        assert np.allclose(got, head, rtol=1e-6, atol=1e-6 * N, equal_nan=True)
        assert sha(np.asarray(l1_counts, dtype=np.uint64)) == case["l1_counts_sha256"], "leaf key counts differ"


@pytest.mark.parametrize("case", GOLD, ids=IDS)
def test_oracle_reproduces_its_snapshots(oracle, case):
    keys = DATA[case["data"]]()
    assert sha(keys) == case["keys_sha256"], "the seeded data set changed: regenerate with tools/make_golden.py"
    o = oracle.train(keys, case["spec"], case["branching_factor"])
    check(case, o.l0.kind, o.l0.fp, o.l0.ip, o.l0.high, o.l0.t32, o.l0.a2, o.l1_params, o.l1_errors, o.l1_counts, o.max_error,
          o.max_error_idx, o.avg_error, o.avg_l2_error, o.avg_log2_error, exact=not case["libm"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLD, ids=IDS)
def test_cuda_build_matches_the_snapshots(case):
    import rmi_b200
    keys = DATA[case["data"]]()
    ds = rmi_b200.RMITrainingData(keys)
    # the serial ("exact") top fit where the top model has one: then nothing but libm separates
    # the CUDA path from the snapshot
    g = rmi_b200.train(ds, case["spec"], case["branching_factor"], rmi_b200.FLAG_TOP_FIT_EXACT)
    kind = "radix_table" if g.l0_model.startswith("radix") and g.l0_model not in ("radix",) and case["l0_kind"] == "radix_table" else g.l0_model
    check(case, kind, g.l0_fparams, g.l0_iparams if case["l0_iparams"] else [], g.l0_bradix_high, g.l0_table32, g.l0_pivots,
          g.l1_params, g.last_layer_max_l1s, g.l1_counts, g.model_max_error, g.model_max_error_idx, g.model_avg_error,
          g.model_avg_l2_error, g.model_avg_log2_error, exact=not case["libm"])
