"""The C++ oracle against the independently written numpy-float32 restatement (oracle/np_oracle.py): coverage,
depth bits and colour must agree exactly (both are IEEE f32, unfused)."""
import numpy as np
import pytest

import euc_b200 as e
from euc_b200 import scenes
from oracle import np_oracle, oracle


def _tris(n, seed, nasty=False, size=0.25):
    r = scenes.u01(seed, n * 3 * 8).reshape(n, 3, 8)
    v = np.zeros((n, 3), dtype=e.VERTEX_P4C4)
    wv = 0.5 + r[:, :, 4] * 1.5
    x = (r[:, :1, 0] * 2.2 - 1.1) + (r[:, :, 2] - 0.5) * size * 2
    y = (r[:, :1, 1] * 2.2 - 1.1) + (r[:, :, 3] - 0.5) * size * 2
    z = r[:, :, 5] * 1.2 - 0.1
    v["pos"][:, :, 0], v["pos"][:, :, 1], v["pos"][:, :, 2], v["pos"][:, :, 3] = x * wv, y * wv, z * wv, wv
    v["rgba"][:, :, :3] = r[:, :, 5:8]
    v["rgba"][:, :, 3] = 0.25 + 0.5 * r[:, :, 6]
    if nasty:
        v["pos"][0::7, 1] = v["pos"][0::7, 0]
        v["pos"][1::9, 2, 3] = -0.5
        v["pos"][2::11, 0, 3] = 0.0
        v["pos"][3::13, 1, 0] = np.nan
        v["pos"][4::15, 2, 1] = np.inf
        v["pos"][5::17, :, :2] *= 30.0
    return v.reshape(-1)


@pytest.mark.parametrize("w,h,n,size,nasty", [(640, 64, 60, 0.25, False), (2000, 30, 40, 0.15, False), (640, 96, 80, 0.3, True),
                                              (100, 200, 30, 0.8, False)])
@pytest.mark.parametrize("cull", ["None", "Back", "Front"])
def test_blend_pipeline_matches_numpy(w, h, n, size, nasty, cull):
    verts = _tris(n, 1000 + w + n, nasty, size)
    px = np.full((h, w), 0xFF000000, dtype=np.uint32)
    z = np.ones((h, w), dtype=np.float32)
    cm = {"None": e.CullMode.NONE, "Back": e.CullMode.Back, "Front": e.CullMode.Front}[cull]
    st = oracle.render(e.BlendTris(cull=cm), verts, px, z)
    npx = np.full((h, w), 0xFF000000, dtype=np.uint32)
    nz = np.ones((h, w), dtype=np.float32)
    rec = []
    np_oracle.render(verts["pos"], verts["rgba"], npx, nz, np_oracle.blend_src_over, cull=cull, depth_test="Less", depth_write=True, record=rec)
    assert st["fragments"] == len(rec)
    assert np.array_equal(z.view(np.uint32), nz.view(np.uint32)), "depth bits differ"
    assert np.array_equal(px, npx), "colour differs"
    if not nasty and w >= 640:
        assert len(rec) > 100


def test_opengl_coords_and_no_depth():
    verts = _tris(40, 77)
    px = np.zeros((64, 640), dtype=np.uint32)
    oracle.render(e.VertexColor(coords=e.CoordinateMode.OPENGL), verts, px, None)
    npx = np.zeros((64, 640), dtype=np.uint32)
    np_oracle.render(verts["pos"], verts["rgba"], npx, None, np_oracle.blend_vertex_color, cull="Back", y_up=True, z_clip=(-1.0, 1.0))
    assert np.array_equal(px, npx)
    assert px.any()


def test_smart_span_path_is_exercised():
    # large triangles (bbox area >= 128 within a band) take the edge-intersection span path (triangles.rs:228-253)
    verts = _tris(12, 4242, size=0.9)
    px = np.full((124, 640), 0xFF000000, dtype=np.uint32)
    z = np.ones((124, 640), dtype=np.float32)
    st = oracle.render(e.BlendTris(cull=e.CullMode.NONE), verts, px, z)
    npx = np.full((124, 640), 0xFF000000, dtype=np.uint32)
    nz = np.ones((124, 640), dtype=np.float32)
    np_oracle.render(verts["pos"], verts["rgba"], npx, nz, np_oracle.blend_src_over, cull="None", depth_test="Less", depth_write=True)
    assert st["fragments"] > 20000
    assert np.array_equal(z.view(np.uint32), nz.view(np.uint32)) and np.array_equal(px, npx)
