"""Shared fixtures.  Tests marked ``gpu`` need a B200 and call the CUDA path through the
C-ABI; everything else runs on CPU (oracle vs golden vectors, host logic, ABI surface)."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "multigpu: needs at least 2 CUDA devices")


def _npz(name):
    with np.load(GOLDEN / name) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_onestep():
    flat = _npz("onestep.npz")
    cases = {}
    for key, val in flat.items():
        case, field = key.split("/")
        cases.setdefault(case, {})[field] = val
    return cases


@pytest.fixture(scope="session")
def golden_traj():
    flat = _npz("trajectories.npz")
    cases = {}
    for key, val in flat.items():
        case, field = key.split("/")
        cases.setdefault(case, {})[field] = val
    return cases


@pytest.fixture(scope="session")
def golden_transform():
    return _npz("transform.npz")


@pytest.fixture(scope="session")
def datasets():
    z = _npz("datasets.npz")
    out = {}
    for name in ("animals", "lastfm", "paleo"):
        n = int(z[f"{name}_shape"][1])
        out[name] = np.unpackbits(z[f"{name}_bits"], axis=1, bitorder="little")[:, :n].astype(np.float64)
    n = out["animals"].shape[1]
    for key in ("train_mask", "val_mask", "test_mask"):
        out[f"animals_{key}"] = np.unpackbits(z[f"animals_{key}_bits"], axis=1, bitorder="little")[:, :n].astype(np.float64)
    return out


def cfg1_matrix():
    """Quick-start synthetic X of BASELINE.json config 1 (reference README.md:67-72)."""
    return (np.random.default_rng(0).random((100, 500)) < 0.25).astype(float)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def rel_err_elementwise(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor if floor > 0 else 1e-300)))
