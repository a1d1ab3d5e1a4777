import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def synthetic_clouds(B, N, seed, kind="ball"):
    """SURVEY 8(d) input distributions, generated with NumPy so CPU and GPU tests share them exactly.
    ball  : gaussian cloud, centred, scaled into the unit ball, per-cloud scale U(2/3,3/2) + shift U(-.2,.2)
    sphere: noisy unit-sphere surface with ~1% exact duplicate points and a few points inside the
            |p|^2 <= 1e-3 skip region (exercises ties and the pointnet2 skip rule)."""
    rng = np.random.default_rng(seed)
    if kind == "ball":
        x = rng.standard_normal((B, N, 3))
        x -= x.mean(axis=1, keepdims=True)
        x /= np.maximum(np.linalg.norm(x, axis=-1).max(axis=1), 1e-12)[:, None, None]
        x = x * rng.uniform(2 / 3, 3 / 2, (B, 1, 3)) + rng.uniform(-0.2, 0.2, (B, 1, 3))
    elif kind == "sphere":
        x = rng.standard_normal((B, N, 3))
        x /= np.linalg.norm(x, axis=-1, keepdims=True)
        x *= 1 + 0.01 * rng.standard_normal((B, N, 1))
        ndup = max(1, N // 100)
        for b in range(B):
            src = rng.integers(0, N, ndup)
            dst = rng.integers(0, N, ndup)
            x[b, dst] = x[b, src]
            near = rng.integers(1, max(N, 2), max(1, N // 200)) % N
            x[b, near] = rng.uniform(-0.015, 0.015, (len(near), 3))
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(x, dtype=np.float32)


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "reference_glue.npz")
    return np.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
