"""Inputs of the benchmark-size HDBSCAN parity cases (tests/golden/hdbscan_big_*.npz).

FOR-instance-shaped synthetic embeddings (BASELINE configs[2] / SURVEY 8d C3: thing points of a forest cylinder,
5-D embeddings = instance centre + noise, a few percent background), generated with INTEGER arithmetic only
(numpy Generator.integers + exact dyadic scaling), so that every machine -- the container that froze the golden
answers with the CPU oracle and the GPU box that replays them -- gets bit-identical float32 inputs (no libm /
SIMD-dependent transcendental functions are involved).
"""
import hashlib

import numpy as np

CASES = {"50k": dict(n=50000, trees=24, seed=11), "350k": dict(n=350000, trees=45, seed=12)}


def make(name):
    c = CASES[name]
    rng = np.random.default_rng(c["seed"])
    n, T, D = c["n"], c["trees"], 5
    mu = rng.integers(-9 * 256, 9 * 256 + 1, (T, D)).astype(np.float64) / 256.0          # centres, sigma ~ 3 * sqrt(3)
    weight = rng.integers(1, 20, T)                                                        # tree sizes vary 20x
    owner = np.searchsorted(np.cumsum(weight), rng.integers(0, int(weight.sum()), n), side="right")
    noise = rng.integers(-2130, 2131, (n, D, 4)).sum(2).astype(np.float64) / 16384.0       # Irwin-Hall, sigma ~ 0.15
    X = mu[owner] + noise
    bg = rng.integers(0, 100, n) < 3                                                       # 3 % background scatter
    X[bg] = rng.integers(-12 * 256, 12 * 256 + 1, (int(bg.sum()), D)).astype(np.float64) / 256.0
    X32 = X.astype(np.float32)
    assert np.array_equal(X32.astype(np.float64), X)                                       # exactly representable
    return X32, np.where(bg, -1, owner)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()
