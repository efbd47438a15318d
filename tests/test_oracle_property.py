"""The reference's integration-test property on the oracle's output: for every key of the
data set, |lookup(key) - lower_bound(key)| <= err  (tests/simple_model_wiki/main.cpp:26-42).
The six reference cases use 200M-key SOSD files that are not available offline; the same
model specs are run on seeded synthetic keys instead."""
import numpy as np
import pytest

from tests import datasets

SPECS = [
    ("cubic,linear", 1024),            # tests/simple_model_wiki/Makefile:8 (262144 on 200M)
    ("robust_linear,linear", 1024),    # tests/simple_model_osm/Makefile:8
    ("radix,linear", 1024),            # tests/radix_model_wiki/Makefile:8
    ("linear,linear", 100),            # BASELINE.json configs[0]
    ("linear,cubic", 512),
    ("linear_spline,linear_spline", 256),
    ("radix18,linear", 2048),
    ("bradix,linear", 512),
    ("histogram,linear", 256),
    ("normal,linear", 128),
    ("linear,loglinear", 64),
]
DATA = {
    "uniform_u64": lambda: datasets.uniform_u64(200_000),
    "lognormal_u64": lambda: datasets.lognormal_u64(200_000),
    "dups_u64": lambda: datasets.with_duplicates(datasets.uniform_u64(200_000)),
    "uniform_u32": lambda: datasets.uniform_u32(200_000),
    "uniform_f64": lambda: datasets.uniform_f64(200_000),
    "front_heavy_u64": lambda: datasets.front_heavy_u64(200_000),
}


@pytest.mark.parametrize("dname", list(DATA))
@pytest.mark.parametrize("spec,bf", SPECS, ids=[f"{s}:{b}" for s, b in SPECS])
def test_error_bound_holds_for_every_key(oracle, spec, bf, dname):
    keys = DATA[dname]()
    try:
        rmi = oracle.train(keys, spec, bf)
    except oracle.OraclePanic as e:
        pytest.skip(f"reference would panic: {e}")
    pos, err = rmi.lookup_batch(keys)
    true_idx = np.searchsorted(keys, keys, side="left").astype(np.uint64)
    diff = np.where(pos > true_idx, pos - true_idx, true_idx - pos)
    bad = np.flatnonzero(diff > err)
    assert bad.size == 0, (spec, bf, dname, int(bad[0]), int(diff[bad[0]]), int(err[bad[0]]))
