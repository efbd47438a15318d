"""Pure-Python restatement of the reference's `--bounded` pre-pass for small inputs.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows rmi_lib/src/cache_fix.rs line by
line — Spline (:5-44), SplineFit (:46-104), cache_fix (:106-150) — over data.iter_unique()
(models/mod.rs:187-231, :286-288).  Release-build semantics: wrapping u64 subtraction,
saturating f64 -> usize casts, active assert!s.  Parity of this file is pinned only by the
reference's own property (tests/cache_fix_wiki/main.cpp: every lookup lands within one line of
the key's lower bound), which tests/test_bounded.py checks on the generated code; the Rust
binary cannot be built here (no cargo), so knot-for-knot parity with it is unpinned.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math

U64 = (1 << 64) - 1


class CacheFixPanic(RuntimeError):
    """The reference would have panicked (an assert! fired)."""


_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.fma.restype = ctypes.c_double
_libm.fma.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double]
_fma = _libm.fma          # f64::mul_add (one rounding)


def _as_usize(v: float) -> int:
    if math.isnan(v) or v <= 0:
        return 0
    return U64 if v >= 18446744073709551615.0 else int(v)


def _predict(sp, inp: int) -> int:
    """Spline::predict, cache_fix.rs:36-43"""
    fx, fy, tx, ty = sp
    den = float(tx - fx)
    num = float((inp - fx) & U64)
    t = num / den if den != 0.0 else (float("nan") if num == 0.0 else float("inf"))
    return _as_usize(_fma(1.0 - t, float(fy), t * float(ty)))


def cache_fix(keys, line_size: int):
    """cache_fix.rs:106-150; keys: sorted sequence of u64; returns the knots [(key, offset)]."""
    n = len(keys)
    if not n > line_size:
        raise CacheFixPanic("Cannot apply a cachefix with fewer items than the line size")
    spline = None          # (from_x, from_y, to_x, to_y)
    curr = []
    out = []

    def add_point(pt):
        nonlocal spline, curr
        if spline is None:                                    # :63-66
            spline = (pt[0], pt[1], pt[0], pt[1])
            return pt
        if pt[0] < spline[0] or pt[1] < spline[1]:            # with_new_dest asserts, :23-30
            raise CacheFixPanic(f"When source x is {spline[0]}, cannot set dest x to {pt[0]}")
        proposed = (spline[0], spline[1], pt[0], pt[1])
        curr.append((spline[2], spline[3]))                   # :71
        if all(_predict(proposed, x) // line_size == y // line_size for x, y in curr):   # :96-103
            spline = proposed
            return None
        prev = (spline[2], spline[3])                         # :76-86
        if not pt[0] > prev[0]:
            raise CacheFixPanic(f"new point: {pt} prev point: {prev}")
        spline = (prev[0], prev[1], pt[0], pt[1])
        curr = [pt]
        return prev

    last_key = 0
    prev_key = None
    for i, key in enumerate(keys):
        key = int(key)
        if prev_key is not None and key == prev_key:          # DedupIter: first item of each run
            continue
        prev_key = key
        km = (key - 1) & U64                                  # minus_epsilon (wrapping in release)
        if not km >= last_key:
            raise CacheFixPanic(f"key: {key} last key: {last_key}, key - e: {km}")
        if km != last_key:
            p = add_point((km, i))
            if p is not None:
                out.append(p)
        p = add_point((key, i))
        if p is not None:
            out.append(p)
        last_key = key
    if spline is not None:
        out.append((spline[2], spline[3]))                    # finish(), :91-93
    return out
