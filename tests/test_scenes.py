"""Scene generators: deterministic, of the BASELINE shapes."""
import numpy as np

from euc_b200 import scenes


def test_splitmix64_known_values():
    # splitmix64 reference outputs for seed 0 (Vigna's splitmix64.c: first three outputs)
    out = scenes.splitmix64(0, 3)
    assert [int(x) for x in out] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
    u = scenes.u01(0xE0C40004, 1000)
    assert u.dtype == np.float32 and (u >= 0).all() and (u < 1).all()


def test_teapot_stream_shape():
    s = scenes.teapot_stream()
    assert s.shape == (6768,) and s.dtype.itemsize == 24  # 2256 faces x 3 face-vertices
    assert np.isfinite(s["pos"]).all() and np.isfinite(s["normal"]).all()


def test_blend_tris_shape_and_determinism():
    v, i = scenes.blend_tris(1 << 10, 3840, 2160)
    v2, i2 = scenes.blend_tris(1 << 10, 3840, 2160)
    assert v.shape == (4 << 10,) and i.shape == (6 << 10,) and i.dtype == np.uint32
    assert v.tobytes() == v2.tobytes() and i.tobytes() == i2.tobytes()
    assert i[:6].tolist() == [0, 1, 2, 2, 1, 3] and i[6:12].tolist() == [4, 5, 6, 6, 5, 7]
    w = v["pos"][:, 3]
    assert (w >= 0.5).all() and (w <= 2.0).all()
    a = v["rgba"][:, 3]
    assert (a >= 0.25).all() and (a <= 0.75).all()


def test_voxel_icons():
    for k in (0, 1, 7):
        v, i = scenes.voxel_icon_mesh(k)
        assert 500 <= i.size // 3 <= 4000, i.size // 3
        assert i.max() < v.size and v.dtype.itemsize == 32
    verts, idx, draws, ubs = scenes.voxel_icon_batch(3)
    assert len(draws) == 3 and len(ubs) == 3 * 80
    assert draws[1][0] == draws[0][1] and draws[1][2] > 0


def test_cube_geometry():
    v, i = scenes.cube_geometry(3.0)
    assert v.shape == (24,) and i.shape == (36,) and v["uv"].max() == 3.0
    m = scenes.cube_mvp(250, 1920, 1080)
    assert m.shape == (4, 4) and m.dtype == np.float32
