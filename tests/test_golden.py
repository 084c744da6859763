"""Golden fixtures (tests/golden/golden.npz, made by tools/make_golden.py from the CPU oracle at BASELINE sizes).
`not gpu`: the oracle still reproduces them.  `gpu`: the CUDA path matches them at the full sizes."""
import os
import zlib

import numpy as np
import pytest

import euc_b200 as e
from euc_b200 import scenes
from oracle import oracle
from conftest import assert_colour_within_1lsb

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz"))


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


# ---- oracle vs golden (CPU) ----------------------------------------------------------------------------
def test_oracle_c1_teapot():
    stream, u = scenes.teapot_stream(), scenes.teapot_uniforms(640, 480, 512)
    shadow = np.full((512, 512), 1.0, np.float32)
    color = np.zeros((480, 640), np.uint32)
    depth = np.full((480, 640), 1.0, np.float32)
    f1 = oracle.render(e.TeapotShadow(u["shadow_mvp"]), stream, None, shadow, n_threads=2)["fragments"]
    f2 = oracle.render(e.Teapot(u["m"], u["v"], u["p"], u["light_pos"], e.Sampler(shadow, e.abi.TEXEL_F32, e.abi.FILTER_LINEAR).clamped(),
                                u["light_vp"], u["cam_pos"]), stream, color, depth, n_threads=2)["fragments"]
    assert [f1, f2] == G["c1_frags"].tolist()
    assert crc(shadow) == G["c1_shadow_crc"] and crc(depth) == G["c1_depth_crc"]
    assert np.array_equal(color, G["c1_color"])


def test_oracle_c2_cube():
    w, h = 1920, 1080
    verts, idx = scenes.cube_geometry(3.0)
    tex = scenes.rust_texture()
    color = np.full((h, w), 180, np.uint32)
    smp = e.Sampler(tex.view(np.uint32).reshape(tex.shape[0], tex.shape[1]), e.abi.TEXEL_RGBA8_TO_F32, e.abi.FILTER_LINEAR).tiled()
    f = oracle.render(e.Cube(scenes.cube_mvp(250, w, h), smp), e.IndexedVertices(idx, verts), color, None, n_threads=0)["fragments"]
    assert f == G["c2_frags"] and crc(color) == G["c2_color_crc"]


def test_oracle_c5_icons():
    verts, idx, draws, ubs = scenes.voxel_icon_batch(4)
    for k in range(4):
        color = np.zeros((256, 256), np.uint32)
        depth = np.full((256, 256), 1.0, np.float32)
        first, count, base, _ = draws[k]
        st = oracle.render(e.VoxelIcon(scenes.voxel_icon_mvp(k), scenes.VOXEL_LIGHT_DIR), e.IndexedVertices(idx, verts), color, depth,
                           draw=(first, count, base))
        assert st["fragments"] == G["c5_frags"][k]
        assert np.array_equal(color, G["c5_color"][k]) and crc(depth) == G["c5_depth_crc"][k]


# ---- CUDA path vs golden at full BASELINE sizes ----------------------------------------------------------
def _gpu_teapot(w, h, s, msaa):
    ctx = e.default_context()
    ctx.set_stats(True)
    stream, u = scenes.teapot_stream(), scenes.teapot_uniforms(w, h, s)
    geom = e.Geometry(stream)
    shadow = e.Buffer2d.fill([s, s], 1.0)
    color = e.Buffer2d.fill([w, h], 0, dtype=np.uint32)
    depth = e.Buffer2d.fill([w, h], 1.0)
    e.TeapotShadow(u["shadow_mvp"]).render(geom, e.Empty(), shadow)
    f1 = ctx.get_stats()["fragments"]
    aa = e.AaMode.Msaa(msaa) if msaa else None
    e.Teapot(u["m"], u["v"], u["p"], u["light_pos"], shadow.linear().clamped(), u["light_vp"], u["cam_pos"], aa=aa).render(geom, color, depth)
    f2 = ctx.get_stats()["fragments"]
    return shadow.raw(), color.raw(), depth.raw(), f1, f2


@pytest.mark.gpu
def test_gpu_c1_teapot_640x480():
    sh, c, d, f1, f2 = _gpu_teapot(640, 480, 512, 0)
    assert [f1, f2] == G["c1_frags"].tolist()
    assert crc(sh) == G["c1_shadow_crc"] and crc(d) == G["c1_depth_crc"]
    assert_colour_within_1lsb(c, G["c1_color"], "C1 colour")


@pytest.mark.gpu
def test_gpu_c3_teapot_4k_msaa():
    sh, c, d, f1, f2 = _gpu_teapot(3840, 2160, 2048, 1)
    assert [f1, f2] == G["c3_frags"].tolist()
    assert crc(sh) == G["c3_shadow_crc"] and crc(d) == G["c3_depth_crc"]
    assert crc(c != 0) == G["c3_coverage_crc"]
    assert_colour_within_1lsb(c[700:1212, 1500:2012], G["c3_color_crop"], "C3 colour crop")


@pytest.mark.gpu
def test_gpu_c2_cube_1080p():
    w, h = 1920, 1080
    ctx = e.default_context()
    ctx.set_stats(True)
    verts, idx = scenes.cube_geometry(3.0)
    tex = e.Buffer2d.from_array(scenes.rust_texture())
    color = e.Buffer2d.fill([w, h], 180, dtype=np.uint32)
    e.Cube(scenes.cube_mvp(250, w, h), tex.linear().tiled()).render(e.IndexedVertices(idx, verts), color, e.Empty())
    assert ctx.get_stats()["fragments"] == G["c2_frags"]
    c = color.raw()
    assert_colour_within_1lsb(c[400:656, 800:1056], G["c2_color_crop"], "C2 crop")
    assert crc(c) == G["c2_color_crc"]  # no transcendental in this shader: expected bit-exact


@pytest.mark.gpu
def test_gpu_c4_full_size():
    w, h = 3840, 2160
    ctx = e.default_context()
    ctx.set_stats(True)
    verts, idx = scenes.blend_tris(1 << 19, w, h)
    geom = e.Geometry(verts, idx)
    color = e.Buffer2d.fill([w, h], 0xFF000000, dtype=np.uint32)
    depth = e.Buffer2d.fill([w, h], 1.0)
    e.BlendTris().render(geom, color, depth)
    assert ctx.get_stats()["fragments"] == G["c4_frags"]
    assert crc(depth.raw()) == G["c4_depth_crc"], "depth buffer not bit-identical"
    c = color.raw()
    assert_colour_within_1lsb(c[1000:1128, 1900:2028], G["c4_color_crop"], "C4 crop")
    assert crc(c) == G["c4_color_crc"]  # + - * / only: expected bit-exact
    # idempotence of the whole path: a second frame after clears gives identical bits
    color.clear(0xFF000000)
    depth.clear(1.0)
    e.BlendTris().render(geom, color, depth)
    assert crc(color.raw()) == G["c4_color_crc"] and crc(depth.raw()) == G["c4_depth_crc"]


@pytest.mark.gpu
def test_gpu_c5_icons_golden():
    verts, idx, draws, ubs = scenes.voxel_icon_batch(4)
    geom = e.Geometry(verts, idx)
    color = e.Buffer2d.fill([256, 256], 0, dtype=np.uint32, layers=4)
    depth = e.Buffer2d.fill([256, 256], 1.0, layers=4)
    e.VoxelIcon(np.eye(4), scenes.VOXEL_LIGHT_DIR).render_batch(geom, draws, ubs, color, depth)
    c, d = color.raw(), depth.raw()
    for k in range(4):
        assert crc(d[k]) == G["c5_depth_crc"][k]
        assert_colour_within_1lsb(c[k], G["c5_color"][k], f"icon {k}")
