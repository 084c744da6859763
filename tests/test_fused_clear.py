"""euc_render_clear: `pixel.clear(a); depth.clear(b); pipe.render(..)` (benches/teapot.rs:183-204) in one call, the clears
fused into the render's kernels.  The fused form must leave exactly the bytes of the three-call form in both targets,
whatever they held before."""
import numpy as np
import pytest

import euc_b200 as e
from euc_b200 import scenes
from test_gpu_parity import _random_tris

pytestmark = pytest.mark.gpu

PX, Z = 0xFF102030, 1.0


def _garbage(w, h, seed, layers=1):
    r = np.random.default_rng(seed)
    c = r.integers(0, 2**32, size=(layers * h, w), dtype=np.uint64).astype(np.uint32)
    z = r.random((layers * h, w), dtype=np.float32)
    return c, z


def _both(render, w, h, px=PX, z=Z, layers=1, clear_px=True, clear_z=True, rows=None):
    """render(pixel, depth, clear) -> None.  Returns ((colour, depth) of the three-call form, (..) of the fused form)."""
    out = []
    for fused in (False, True):
        gc, gz = _garbage(w, h, 1, layers)  # the same for both forms: a target that is not cleared must end up identical too
        c = e.Buffer2d([w, h], np.uint32, layers=layers); c.upload(gc)
        d = e.Buffer2d([w, h], np.float32, layers=layers); d.upload(gz)
        if fused:
            render(c, d, (px if clear_px else None, z if clear_z else None))
            got = (c.raw().copy(), d.raw().copy())
            if rows is not None:  # rows outside the range keep what they held: compare only the rendered ones
                got = tuple(a.reshape(layers, h, w)[:, rows[0]:rows[1]].copy() for a in got)
                keep = tuple(a.reshape(layers, h, w)[:, :rows[0]] for a in (c.raw(), d.raw()))
                assert np.array_equal(keep[0], gc.reshape(layers, h, w)[:, :rows[0]]) and np.array_equal(keep[1], gz.reshape(layers, h, w)[:, :rows[0]])
        else:
            r0, r1 = rows if rows is not None else (0, h)
            if clear_px:
                c.clear_rows(px, r0, r1)
            if clear_z:
                d.clear_rows(z, r0, r1)
            render(c, d, None)
            got = (c.raw().copy(), d.raw().copy())
            if rows is not None:
                got = tuple(a.reshape(layers, h, w)[:, rows[0]:rows[1]].copy() for a in got)
        out.append(got)
    return out


def _assert_same(a, b, what):
    assert np.array_equal(a[0], b[0]), f"{what}: colour differs"
    assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), f"{what}: depth differs"


@pytest.mark.parametrize("w,h", [(640, 480), (333, 217), (1920, 64)])
@pytest.mark.parametrize("depth", [e.DepthMode.LESS_WRITE, e.DepthMode.LESS_PASS, e.DepthMode.GREATER_WRITE, e.DepthMode.NONE])
def test_immediate_pipeline(w, h, depth):
    verts = _random_tris(400, 0xC1EA + w, size=0.08)  # small triangles: many tiles stay empty
    zc = 0.0 if depth is e.DepthMode.GREATER_WRITE else Z
    a, b = _both(lambda c, d, clr: e.BlendTris(depth=depth).render(verts, c, d, clear=clr), w, h, z=zc)
    _assert_same(a, b, f"BlendTris {w}x{h} {depth}")
    assert (a[0] == PX).any() and (a[0] != PX).any()


@pytest.mark.parametrize("aa", [None, 1, 2])
@pytest.mark.parametrize("w,h", [(640, 480), (333, 217)])
def test_deferred_pipeline(aa, w, h):
    verts = _random_tris(300, 0xDEF + w, size=0.15)
    kw = dict(aa=e.AaMode.Msaa(aa)) if aa else {}
    for depth in (e.DepthMode.LESS_WRITE, e.DepthMode.NONE):
        a, b = _both(lambda c, d, clr: e.VertexColor(depth=depth, **kw).render(verts, c, d, clear=clr), w, h, px=0x00000000)
        _assert_same(a, b, f"VertexColor aa={aa} {w}x{h} {depth}")


def test_one_target_only_and_empty_geometry():
    verts = _random_tris(300, 0x0E, size=0.1)
    w, h = 512, 256
    for cp, cz in ((True, False), (False, True)):
        a, b = _both(lambda c, d, clr: e.BlendTris().render(verts, c, d, clear=clr), w, h, clear_px=cp, clear_z=cz)
        _assert_same(a, b, f"clear colour={cp} depth={cz}")
    none = np.zeros(0, dtype=e.VERTEX_P4C4)
    a, b = _both(lambda c, d, clr: e.BlendTris().render(none, c, d, clear=clr), w, h)
    _assert_same(a, b, "no primitives")
    assert (b[0] == PX).all() and (b[1] == Z).all()
    # a request is consumed by the next render only
    c = e.Buffer2d.fill([w, h], 7, dtype=np.uint32)
    d = e.Buffer2d.fill([w, h], 0.5)
    e.BlendTris().render(none, c, d, clear=(PX, Z))
    c.clear(9)
    e.BlendTris().render(none, c, d)
    assert (c.raw() == 9).all()


def test_rows_and_layers():
    w, h = 640, 480
    verts, idx = scenes.blend_tris(1 << 10, w, h, seed=5)
    geom = e.Geometry(verts, idx)
    rows = (160, 320)
    a, b = _both(lambda c, d, clr: e.BlendTris().render(geom, c, d, rows=rows, clear=clr), w, h, rows=rows)
    _assert_same(a, b, "row band")
    # batch: three layers, draws only into layers 0 and 2 -- layer 1 must still be cleared
    n = 3
    verts, idx, draws, ubs = scenes.voxel_icon_batch(n)
    draws = np.asarray([draws[0], draws[2]], dtype=np.int64).copy()
    draws[1][3] = 2
    ub2 = ubs[:80] + ubs[160:240]
    g2 = e.Geometry(verts, idx)
    pipe = e.VoxelIcon(np.eye(4), scenes.VOXEL_LIGHT_DIR)
    a, b = _both(lambda c, d, clr: pipe.render_batch(g2, draws, ub2, c, d, clear=clr), 256, 256, px=0, layers=3)
    _assert_same(a, b, "batch layers")
    assert (b[0].reshape(3, 256, 256)[1] == 0).all() and (b[1].reshape(3, 256, 256)[1] == Z).all()


def test_two_pass_teapot_with_shadow_map():
    w, h, s = 640, 480, 512
    stream, u = scenes.teapot_stream(), scenes.teapot_uniforms(w, h, s)
    geom = e.Geometry(stream)
    res = []
    for fused in (False, True):
        gs = np.random.default_rng(3).random((s, s), dtype=np.float32)
        gc, gz = _garbage(w, h, 5)
        shadow = e.Buffer2d([s, s], np.float32); shadow.upload(gs)
        color = e.Buffer2d([w, h], np.uint32); color.upload(gc)
        depth = e.Buffer2d([w, h], np.float32); depth.upload(gz)
        p1 = e.TeapotShadow(u["shadow_mvp"])
        p2 = e.Teapot(u["m"], u["v"], u["p"], u["light_pos"], shadow.linear().clamped(), u["light_vp"], u["cam_pos"], aa=e.AaMode.Msaa(1))
        if fused:
            p1.render(geom, e.Empty(), shadow, clear=(None, 1.0))
            p2.render(geom, color, depth, clear=(0, 1.0))
        else:
            shadow.clear(1.0); color.clear(0); depth.clear(1.0)
            p1.render(geom, e.Empty(), shadow)
            p2.render(geom, color, depth)
        res.append((shadow.raw().copy(), color.raw().copy(), depth.raw().copy()))
    for k, name in enumerate(("shadow map", "colour", "depth")):
        assert np.array_equal(res[0][k].view(np.uint32), res[1][k].view(np.uint32)), name


def test_row_range_clear_of_an_odd_width_target():
    """Round 1 returned EUC_E_UNSUPPORTED for a row range whose first texel is not 16-byte aligned (width not a multiple of 4:
    the plain fill behind a clear the tile kernel cannot fuse, e.g. the colour target of a depth-only pass).  The fill now has
    a scalar head and tail."""
    w, h = 333, 160
    verts = _random_tris(300, 5)
    geom = e.Geometry(verts)
    gc, gz = _garbage(w, h, 9)
    c = e.Buffer2d([w, h], np.uint32); c.upload(gc)
    d = e.Buffer2d([w, h], np.float32); d.upload(gz)
    # depth-only use of the blending pipeline: pixel writes off, so the colour clear is a plain fill of rows [48, 112)
    pipe = e.BlendTris(pixel=e.PixelMode.PASS, depth=e.DepthMode.LESS_WRITE)
    pipe.render(geom, c, d, rows=(48, 112), clear=(PX, Z))
    got_c, got_z = c.raw(), d.raw()
    assert (got_c[48:112] == PX).all() and np.array_equal(got_c[:48], gc[:48]) and np.array_equal(got_c[112:], gc[112:])
    ref_c = e.Buffer2d([w, h], np.uint32); ref_c.upload(gc)
    ref_d = e.Buffer2d([w, h], np.float32); ref_d.upload(gz)
    ref_d.clear_rows(Z, 48, 112)
    pipe.render(geom, ref_c, ref_d, rows=(48, 112))
    assert np.array_equal(got_z.view(np.uint32), ref_d.raw().view(np.uint32))
    # and the stand-alone row clear on unaligned rows
    c.clear_rows(7, 1, 2)
    row = c.raw()
    assert (row[1] == 7).all() and np.array_equal(row[0], got_c[0]) and np.array_equal(row[2], got_c[2])
