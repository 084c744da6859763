"""Known-answer tests that pin the oracle to Rust language semantics and to hand-derivable results.
The reference ships no tests or golden vectors, so these (plus tests/test_oracle_vs_numpy.py) are what pins it."""
import math

import numpy as np
import pytest

import euc_b200 as e
from euc_b200 import abi
from oracle import oracle

L = oracle.lib()
nan, inf = float("nan"), float("inf")


def test_f32_as_usize_saturating_cast():
    # Rust `as`: truncate toward zero, saturate, NaN -> 0  (drives triangles.rs:117-138, linear.rs:36-37)
    assert L.oracle_f32_as_usize(3.99) == 3
    assert L.oracle_f32_as_usize(-0.5) == 0
    assert L.oracle_f32_as_usize(-1e30) == 0
    assert L.oracle_f32_as_usize(nan) == 0
    assert L.oracle_f32_as_usize(inf) == 2 ** 64 - 1
    assert L.oracle_f32_as_usize(1e30) == 2 ** 64 - 1
    assert L.oracle_f32_as_usize(16777216.0) == 16777216


def test_f32_as_u8():
    assert [L.oracle_f32_as_u8(x) for x in (-3.0, 0.99, 1.0, 254.999, 255.0, 300.0, nan, inf)] == [0, 0, 1, 254, 255, 255, 0, 255]


def test_min_max_ignore_nan():
    assert L.oracle_f32_min(nan, 2.0) == 2.0 and L.oracle_f32_min(2.0, nan) == 2.0
    assert L.oracle_f32_max(nan, -2.0) == -2.0 and L.oracle_f32_max(-2.0, nan) == -2.0
    assert math.isnan(L.oracle_f32_min(nan, nan))
    assert L.oracle_f32_min(1.0, 2.0) == 1.0 and L.oracle_f32_max(1.0, 2.0) == 2.0


def test_fract_and_rem_euclid():
    assert L.oracle_fract(1.75) == 0.75
    assert L.oracle_fract(-1.75) == -0.75  # negative stays negative
    assert L.oracle_fract(1.0) == 0.0
    assert L.oracle_rem_euclid(1.25, 1.0) == 0.25
    assert L.oracle_rem_euclid(-0.25, 1.0) == 0.75
    assert L.oracle_rem_euclid(-3.0, 1.0) == 0.0
    # a tiny negative input rounds to exactly 1.0 (r + 1.0 in f32)
    assert L.oracle_rem_euclid(-1e-9, 1.0) == 1.0
    assert L.oracle_rem_euclid(3.5, 2.0) == 1.5


def test_wrap_adaptors():
    # Clamped: max(0).min(1); Tiled: rem_euclid(1); Mirrored (sampler/mod.rs:159-168)
    assert L.oracle_wrap(abi.WRAP_CLAMP, -0.2) == 0.0 and L.oracle_wrap(abi.WRAP_CLAMP, 1.7) == 1.0
    assert L.oracle_wrap(abi.WRAP_CLAMP, nan) == 0.0  # NaN.max(0.0) == 0.0
    assert L.oracle_wrap(abi.WRAP_TILE, 2.25) == 0.25 and L.oracle_wrap(abi.WRAP_TILE, -0.25) == 0.75
    assert L.oracle_wrap(abi.WRAP_MIRROR, 0.25) == 0.25
    assert L.oracle_wrap(abi.WRAP_MIRROR, 1.25) == 0.75  # rem_euclid(2) = 1.25 >= 1 -> 1 - 0.25
    assert L.oracle_wrap(abi.WRAP_MIRROR, -0.25) == 0.25  # rem_euclid(2) = 1.75 -> 1 - 0.75
    assert L.oracle_wrap(abi.WRAP_NONE, 5.5) == 5.5


def test_linear_sampler_2x2():
    # Linear on a 2x2 texture at (0.25, 0.25): ix = 0.5 -> average of the 4 texels by the a13 formula
    t = np.array([[1.0, 2.0], [3.0, 5.0]], dtype=np.float32)
    s = lambda x, y, f=abi.FILTER_LINEAR, w=abi.WRAP_NONE: L.oracle_sample_f32(t.ctypes.data, 2, 2, f, w, x, y)
    assert s(0.25, 0.25) == 2.75
    assert s(0.0, 0.0) == 1.0
    assert s(0.5, 0.0) == 2.0          # ix = 1.0: p = 1, f = 0; tap p+1 clamps to size-1
    assert s(0.75, 0.0) == 2.0         # ix = 1.5: both taps clamp to texel 1
    assert s(1.0, 0.0) == 1.0          # fract(1.0) = 0: x == 1.0 wraps to texel 0 (linear.rs:31)
    assert s(0.25, 0.75) == 4.0        # y clamps: (3+5)/2
    # negative input: fract stays negative, p saturates to 0, weights extrapolate: t0*(1-f) + t1*f with f = -0.5
    assert s(-0.25, 0.0) == np.float32(1.0 * 1.5 + 2.0 * -0.5)
    # nearest: ((x*size).max(0) as usize).min(size-1)
    n = lambda x, y: s(x, y, abi.FILTER_NEAREST)
    assert n(0.49, 0.0) == 1.0 and n(0.5, 0.0) == 2.0 and n(7.0, 0.0) == 2.0 and n(-3.0, 0.99) == 3.0


def test_rgba8_map_sampler():
    # texture.map(|p| p as f32): 0..255, not normalised (texture_mapping.rs:119-121)
    t = np.array([[[10, 20, 30, 255], [50, 60, 70, 255]]], dtype=np.uint8)
    out = np.zeros(4, dtype=np.float32)
    L.oracle_sample_rgba8(t.ctypes.data, 2, 1, abi.FILTER_LINEAR, abi.WRAP_NONE, 0.25, 0.0, out.ctypes.data)
    assert out.tolist() == [30.0, 40.0, 50.0, 255.0]


@pytest.mark.parametrize("w,h,msaa,rows,bands", [
    (512, 512, 0, 39, 14), (640, 480, 0, 31, 16), (1920, 1080, 0, 10, 108), (2048, 2048, 0, 9, 228),
    (3840, 2160, 1, 10, 216), (3840, 2160, 2, 20, 108), (3840, 2160, 0, 5, 432), (256, 256, 0, 78, 4)])
def test_band_table(w, h, msaa, rows, bands):
    # SURVEY §8 band table, from pipeline.rs:328-330
    import ctypes
    nt = ctypes.c_uint64()
    g = L.oracle_band_rows(w, h, msaa, 1 << 20, ctypes.byref(nt))
    assert g == rows
    assert -(-h // g) == bands
    assert nt.value == h // g


def test_zero_thread_quirk_renders_nothing():
    # 1x1 and 32x32: needed_threads = h / group_rows = 0 (pipeline.rs:330, :337)
    verts = np.zeros(3, dtype=e.VERTEX_P4C4)
    verts["pos"] = [(-1, -1, 0, 1), (1, -1, 0, 1), (0, 1, 0, 1)]
    verts["rgba"] = 1.0
    for s in (1, 32):
        px = np.zeros((s, s), dtype=np.uint32)
        st = oracle.render(e.VertexColor(), verts, px, None)
        assert st["fragments"] == 0 and not px.any()
    px = np.zeros((625, 32), dtype=np.uint32)  # h == group_rows: one thread, everything renders
    assert oracle.render(e.VertexColor(), verts, px, None)["fragments"] > 0


def test_readme_triangle_known_answers():
    # README.md:43-50 / examples/triangle.rs:31-39 at 640x480, VULKAN coordinates (SURVEY §8c self-check recipe)
    verts = np.zeros(3, dtype=e.VERTEX_P4C4)
    verts["pos"] = [(-1, -1, 0, 1), (1, -1, 0, 1), (0, 1, 0, 1)]
    verts["rgba"] = [(1, 0, 0, 1), (0, 1, 0, 1), (0, 0, 1, 1)]
    px = np.zeros((480, 640), dtype=np.uint32)
    st = oracle.render(e.VertexColor(), verts, px, None)
    assert st["primitives"] == 1
    assert px[1, 320] != 0 and px[0, 0] == 0                      # apex row covered near x=320; corner not
    assert px[479, 0] == 0 and (px[479, 1:640] != 0).all()        # edges at x ~ 0.67 / 639.33 on the last row
    assert np.count_nonzero(px) == st["fragments"]                # no depth test, one triangle: one fragment per pixel
    assert abs(st["fragments"] - 640 * 480 / 2) < 640             # area of the triangle
    # winding = +4 >= 0 -> not culled under Back, culled under Front
    px2 = np.zeros((480, 640), dtype=np.uint32)
    assert oracle.render(e.VertexColor(cull=e.CullMode.Front), verts, px2, None)["fragments"] == 0
    # apex is blue (vertex 2), bottom-left red: RGBA little-endian bytes
    assert px[1, 320] >> 24 == 0xFF and (px[1, 320] >> 16) & 0xFF > 250
    assert px[479, 1] & 0xFF > 250


def test_thread_count_does_not_change_results():
    verts = np.zeros(3 * 50, dtype=e.VERTEX_P4C4)
    r = e.scenes.u01(5, verts.size * 8).reshape(-1, 8)
    verts["pos"][:, :2] = r[:, :2] * 2 - 1
    verts["pos"][:, 2] = r[:, 2]
    verts["pos"][:, 3] = 1.0
    verts["rgba"] = r[:, 3:7]
    outs = []
    for nt in (1, 2, 0):
        px = np.zeros((480, 640), dtype=np.uint32)
        z = np.ones((480, 640), dtype=np.float32)
        oracle.render(e.BlendTris(), verts, px, z, n_threads=nt)
        outs.append((px, z))
    for px, z in outs[1:]:
        assert np.array_equal(px, outs[0][0]) and np.array_equal(z, outs[0][1])


def test_row_restricted_equals_full():
    verts = np.zeros(3 * 80, dtype=e.VERTEX_P4C4)
    r = e.scenes.u01(9, verts.size * 8).reshape(-1, 8)
    verts["pos"][:, :2] = r[:, :2] * 2 - 1
    verts["pos"][:, 2] = r[:, 2]
    verts["pos"][:, 3] = 0.5 + r[:, 7]
    verts["rgba"] = r[:, 3:7]
    full = np.zeros((480, 640), dtype=np.uint32)
    fz = np.ones((480, 640), dtype=np.float32)
    oracle.render(e.BlendTris(), verts, full, fz)
    part = np.zeros((480, 640), dtype=np.uint32)
    pz = np.ones((480, 640), dtype=np.float32)
    # bands are 31 rows; row ranges that are not band-aligned still select whole bands that intersect them
    for r0, r1 in [(0, 155), (155, 310), (310, 480)]:
        oracle.render(e.BlendTris(), verts, part, pz, rows=(r0, r1))
    assert np.array_equal(full, part) and np.array_equal(fz, pz)
