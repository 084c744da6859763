"""A fixed 200-scene slice of the differential soak (tools/fuzz_parity.py) in the `-m gpu` suite: random target sizes (incl. the
"renders nothing" quirk), 1 .. 3000 triangles from sub-pixel to screen-filling, perspective and w <= 0, snapped / NaN / Inf
vertices, every depth / cull / coordinate mode, MSAA levels 0 - 3, immediate and deferred pipelines, fused clears; textured
cubes with random textures, filters and wrap modes; line lists.  CUDA path against the oracle: depth and fragment counts
bit-exact, colour within 1 LSB."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [7, 8])
def test_fuzz_triangles_slice(seed):
    import fuzz_parity
    frags = sum(fuzz_parity.triangle_scene(k, seed) for k in range(1, 71))
    assert frags > 100000


def test_fuzz_samplers_and_lines_slice():
    import fuzz_parity
    kinds = [fuzz_parity.sampler_or_line_scene(k, 9) for k in range(1, 61)]
    assert kinds.count("cube") >= 10 and kinds.count("lines") >= 10
