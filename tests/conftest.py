import os
import sys

import numpy as np
import pytest

# Two tests run two ranks of a multi-GPU group as two host threads on ONE GPU.  With CUDA's default lazy module loading the
# first launch of a kernel synchronises the device, which would wait for the other rank's barrier kernel (spinning until this
# rank arrives): load every kernel up front.  Must be set before CUDA initialises.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def ctx():
    """The device context.  No CPU fallback: fails loudly if the CUDA library or device is missing."""
    import euc_b200
    return euc_b200.default_context()


def channel_diff(a, b):
    """max abs difference per 8-bit channel between two u32 images."""
    a8 = np.ascontiguousarray(a).view(np.uint8).astype(np.int16)
    b8 = np.ascontiguousarray(b).view(np.uint8).astype(np.int16)
    return int(np.abs(a8 - b8).max()) if a8.size else 0


def assert_depth_bit_exact(gpu, ref, what=""):
    g, r = np.ascontiguousarray(gpu).view(np.uint32), np.ascontiguousarray(ref).view(np.uint32)
    bad = np.argwhere(g != r)
    assert bad.size == 0, f"{what}: {bad.shape[0]} depth texels differ, first at (y,x)={tuple(bad[0])}: gpu={gpu[tuple(bad[0])]!r} ref={ref[tuple(bad[0])]!r}"


def assert_colour_within_1lsb(gpu, ref, what=""):
    d = channel_diff(gpu, ref)
    assert d <= 1, f"{what}: colour differs by {d} LSB (> 1)"
