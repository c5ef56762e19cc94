"""Shared pytest configuration: registers the `gpu` marker and exposes repo paths."""

from __future__ import annotations

import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def golden():
    def load(name: str):
        return np.load(os.path.join(GOLDEN, name + ".npz"))

    return load


def ulp_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units-in-the-last-place between two float32 arrays (sign-magnitude aware)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)

    def key(x):
        i = x.view(np.int32).astype(np.int64)
        return np.where(i < 0, -(i & 0x7FFFFFFF), i)

    return np.abs(key(a) - key(b))
