"""`--param-grid` host logic (host/param_grid.hpp; reference src/main.rs:171-261): the grid reader against
Python's json module, the `<file>_results` writer (field names, the reference's "average error %" quirk)."""
import json
import os
import subprocess

import pytest

from tests.test_codegen import ROOT


@pytest.fixture(scope="module")
def pg_tool(tmp_path_factory):
    d = tmp_path_factory.mktemp("pg")
    exe = str(d / "param_grid_tool")
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cxx", "param_grid_tool.cpp"), "-o", exe], check=True)
    return exe


def test_grid_reader_agrees_with_json(pg_tool, tmp_path):
    grid = {"configs": [
        {"layers": "linear,linear", "branching factor": 1024, "namespace": "ns_a", "size": 99, "binary": True},
        {"namespace": "with \"quotes\" and \\ slash", "branching factor": 64, "layers": "radix18,linear_spline"},
        {"layers": "cubic,linear", "branching factor": 16777216},
        {"layers": "bradix,linear", "branching factor": 2.0e3, "namespace": None, "extra": [1, 2, {"x": -1.5e-3}], "u": "Ab"},
    ], "other": {"nested": []}}
    for text in (json.dumps(grid), json.dumps(grid, indent=2), json.dumps(grid, separators=(",", ":"))):
        path = str(tmp_path / "grid.json")
        open(path, "w").write(text)
        r = subprocess.run([pg_tool, "parse", path], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        got = [ln.split("\t") for ln in r.stdout.splitlines()]
        want = [[c["layers"], str(int(c["branching factor"])), c.get("namespace") or "<none>"] for c in grid["configs"]]
        assert got == want
    # the file --optimize writes is itself a valid grid (optimizer.rs:208-217 to_grid_spec)
    opt = {"configs": [{"layers": "robust_linear,linear", "branching factor": 4096, "namespace": "f_0", "size": 98328,
                        "average log2 error": 3.25, "binary": True}]}
    open(str(tmp_path / "o.json"), "w").write(json.dumps(opt))
    r = subprocess.run([pg_tool, "parse", str(tmp_path / "o.json")], capture_output=True, text=True)
    assert r.stdout.strip().split("\t") == ["robust_linear,linear", "4096", "f_0"]


@pytest.mark.parametrize("text,msg", [
    ('{"configs": 3}', "Configs must have an array"), ('{"nope": []}', "Configs must have an array"),
    ('{"configs": [{"layers": 5, "branching factor": 1}]}', "unwrap"), ('{"configs": [{"layers": "a,b"}]}', "unwrap"),
    ('{"configs": [{"layers": "a,b", "branching factor": -4}]}', "unwrap"), ('{"configs": [', "end of JSON"),
    ('{"configs": []} trailing', "trailing"), ('{"configs": [}', "unexpected character"),
])
def test_grid_reader_rejects_what_the_reference_panics_on(pg_tool, tmp_path, text, msg):
    path = str(tmp_path / "bad.json")
    open(path, "w").write(text)
    r = subprocess.run([pg_tool, "parse", path], capture_output=True, text=True)
    assert r.returncode == 1 and msg in r.stderr, (r.returncode, r.stderr)


def test_results_writer(pg_tool):
    lines = "linear,linear 1024 ns_a 12.5 900.25 3.5 7.0 128 24584\ncubic,linear 64 - 0.1 0.02 1.0 2.0 3 2592\n"
    r = subprocess.run([pg_tool, "results", "1000"], input=lines, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    res = json.loads(r.stdout)["results"]
    assert list(res[0].keys()) == ["layers", "branching factor", "average error", "average error %", "average l2 error",
                                   "average log2 error", "max error", "max error %", "max log2 error", "size binary search",
                                   "namespace"]                      # main.rs:205-220, in this order
    assert res[0]["layers"] == "linear,linear" and res[0]["branching factor"] == 1024 and res[0]["namespace"] == "ns_a"
    assert res[0]["average error"] == 12.5 and res[0]["max error"] == 128 and res[0]["size binary search"] == 24584
    # main.rs:210-211: "average error %" is computed from the MAX error
    assert res[0]["average error %"] == res[0]["max error %"] == 128 / 1000 * 100.0
    assert res[1]["namespace"] is None and res[1]["max error %"] == 3 / 1000 * 100.0


@pytest.mark.gpu
def test_cli_param_grid_on_gpu(tmp_path):
    """`rmi <file> --param-grid grid.json`: every entry is built, `grid.json_results` reports it, and only the
    entries that name a namespace get source files + parameter blobs (main.rs:171-261)."""
    from rmi_b200 import build
    from tests import datasets
    from tests.test_codegen import write_keyfile
    cli = build.build_cli()
    keys = datasets.uniform_u64(500_000, seed=21)
    work = str(tmp_path)
    datafile = os.path.join(work, "synthetic_500k_uint64")
    write_keyfile(datafile, keys)
    grid = {"configs": [{"layers": "linear,linear", "branching factor": 2048, "namespace": "pg_a"},
                        {"layers": "radix,linear_spline", "branching factor": 512}]}
    gpath = os.path.join(work, "grid.json")
    open(gpath, "w").write(json.dumps(grid))
    r = subprocess.run([cli, datafile, "--param-grid", gpath, "--zero-build-time"], cwd=work, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    res = json.load(open(gpath + "_results"))["results"]
    assert [(x["layers"], x["branching factor"], x["namespace"]) for x in res] == [("linear,linear", 2048, "pg_a"),
                                                                                  ("radix,linear_spline", 512, None)]
    for x in res:
        assert x["max error"] > 0 and x["max error %"] == x["max error"] / keys.size * 100.0 == x["average error %"]
        assert x["average log2 error"] > 0 and x["size binary search"] == x["branching factor"] * 24 + 16
    assert os.path.exists(os.path.join(work, "pg_a.cpp")) and os.path.exists(os.path.join(work, "pg_a.h"))
    assert os.path.getsize(os.path.join(work, "rmi_data", "pg_a_L1_PARAMETERS")) == 2048 * 24
    assert not [f for f in os.listdir(work) if f.endswith(".cpp") and f != "pg_a.cpp"]
