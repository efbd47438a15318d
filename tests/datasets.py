"""Seeded synthetic key sets shared by the tests and bench.py.
All arrays are sorted ascending, duplicates kept unless stated."""
import numpy as np


def uniform_u64(n: int, seed: int = 42) -> np.ndarray:
    rng = np.random.Generator(np.random.MT19937(seed))
    k = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
    k.sort()
    return k


def uniform_u32(n: int, seed: int = 7) -> np.ndarray:
    rng = np.random.Generator(np.random.MT19937(seed))
    k = rng.integers(0, 1 << 32, size=n, dtype=np.uint32)
    k.sort()
    return k


def lognormal_u64(n: int, seed: int = 3, sigma: float = 2.0) -> np.ndarray:
    """exp(N(0, sigma)) * 2^40, rounded: heavy skew, many empty leaves and a few huge ones."""
    rng = np.random.Generator(np.random.MT19937(seed))
    k = np.rint(np.exp(rng.normal(0.0, sigma, size=n)) * float(1 << 40)).astype(np.uint64)
    k.sort()
    return k


def with_duplicates(keys: np.ndarray, frac: float = 0.05, seed: int = 5) -> np.ndarray:
    """Overwrite ~frac of the keys with a copy of their left neighbour (runs of equal keys)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    k = keys.copy()
    idx = np.flatnonzero(rng.random(k.size) < frac)
    idx = idx[idx > 0]
    for i in idx:          # sequential so that runs longer than 2 appear
        k[i] = k[i - 1]
    k.sort()
    return k


def uniform_f64(n: int, seed: int = 11) -> np.ndarray:
    rng = np.random.Generator(np.random.MT19937(seed))
    k = rng.random(n) * float(1 << 52)
    k.sort()
    return k


def lognormal_f64(n: int, seed: int = 11, sigma: float = 2.0) -> np.ndarray:
    rng = np.random.Generator(np.random.MT19937(seed))
    k = np.exp(rng.normal(0.0, sigma, size=n))
    k.sort()
    return k


def front_heavy_u64(n: int, seed: int = 13) -> np.ndarray:
    """Three quarters of the keys packed into [0, 2^20), the rest spread over [2^20, 2^63):
    under a radix / spline top model the first leaf holds a very long run of keys while all its
    neighbours are short (exercises the long-leaf paths of the fused leaf kernel)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    a = rng.integers(0, 1 << 20, size=(3 * n) // 4, dtype=np.uint64)
    b = rng.integers(1 << 20, 1 << 63, size=n - a.size, dtype=np.uint64)
    k = np.concatenate([a, b])
    k.sort()
    return k
