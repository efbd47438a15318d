"""CPU-side checks of the C-ABI library: it loads and exports every symbol the header declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    from rmi_b200 import build
    path = build.build_library()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "rmi_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rmi_[a-z0-9_]+)\s*\(", header))
    assert {"rmi_train", "rmi_dataset_create", "rmi_result_free", "rmi_last_error"} <= declared
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/rmi_b200.h but not exported"


def test_result_struct_layout_matches_header():
    """The ctypes mirror must have the size the C compiler gives struct rmi_result."""
    import subprocess
    import tempfile
    from rmi_b200.api import _Result
    src = '#include <stdio.h>\n#include "rmi_b200.h"\nint main(){printf("%zu\\n", sizeof(rmi_result));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        size = int(subprocess.run([exe], capture_output=True, text=True, check=True).stdout)
    assert size == ctypes.sizeof(_Result)


def test_errors_without_a_device_are_reported_not_fatal():
    import numpy as np
    import rmi_b200
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    try:
        rmi_b200.RMITrainingData(np.arange(10, dtype=np.uint64))
    except rmi_b200.RMIError as e:
        assert "rmi_b200 error 3" in str(e)   # RMI_ERR_CUDA, with the CUDA message
    else:
        raise AssertionError("dataset creation must fail loudly without a GPU")


def test_shard_top_rounds_agree_with_the_orchestrator():
    """rmi_shard_top_rounds (the C side's statement of which collectives a top model needs) and
    rmi_b200/sharded.py's TOP_ROUNDS must describe the same protocol."""
    from rmi_b200 import build, sharded
    lib = ctypes.CDLL(build.build_library())
    lib.rmi_shard_top_rounds.argtypes = [ctypes.c_char_p]
    lib.rmi_shard_top_rounds.restype = ctypes.c_int
    code = {(): 0, ("sum",): 1, ("sum", "sum"): 2, ("min", "sum"): 3}
    for name in ("linear", "robust_linear", "linear_spline", "cubic", "loglinear", "normal", "lognormal", "radix",
                 "radix18", "bradix", "histogram", "no_such_model"):
        want = code[sharded.TOP_ROUNDS[name]] if name in sharded.TOP_ROUNDS else (4 if name in sharded.NATIVE_ONLY_TOPS else -1)
        assert lib.rmi_shard_top_rounds(name.encode()) == want, name
