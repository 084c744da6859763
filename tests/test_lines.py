"""Row N1: the Lines rasteriser (src/rasterizer/lines.rs:12-120) and LineList / LineTriangleList assembly
(src/primitives.rs:49-104).  clipline 0.2 is not in the reference tree: the walk is a restatement (closed-form
Bresenham, ties stay) — PARITY UNPINNED; these tests pin the oracle to a brute-force walk and the device to the oracle."""
import numpy as np
import pytest

import euc_b200 as e
from euc_b200 import scenes, vek
from oracle import oracle


def naive_bresenham(x1, y1, x2, y2):
    dx, dy = abs(x2 - x1), abs(y2 - y1)
    sx, sy = (1 if x1 < x2 else -1), (1 if y1 < y2 else -1)
    pts, x, y = [], x1, y1
    if dx >= dy:
        err = 2 * dy - dx
        for i in range(dx + 1):
            pts.append((x, y))
            if i == dx:
                break
            if err > 0:
                y += sy
                err -= 2 * dx
            err += 2 * dy
            x += sx
    else:
        err = 2 * dx - dy
        for i in range(dy + 1):
            pts.append((x, y))
            if i == dy:
                break
            if err > 0:
                x += sx
                err -= 2 * dy
            err += 2 * dx
            y += sy
    return pts


def line_verts(pts, w=1.0):
    v = np.zeros(len(pts), dtype=e.VERTEX_P4C4)
    for i, (x, y, z) in enumerate(pts):
        v["pos"][i] = (x * w, y * w, z * w, w)
        v["rgba"][i] = ((i * 37 % 256) / 255.0, (i * 91 % 256) / 255.0, (i * 53 % 256) / 255.0, 1.0)
    return v


def ndc(px, py, w, h):
    """NDC of the pixel centre-ish point that lands on screen coordinate (px + 0.25, py + 0.25)."""
    return ((px + 0.25) / w * 2 - 1, 1 - (py + 0.25) / h * 2)


@pytest.mark.parametrize("a,b", [((10, 20), (200, 90)), ((300, 10), (20, 70)), ((50, 5), (60, 99)), ((400, 90), (390, 3)),
                                 ((0, 0), (639, 99)), ((100, 50), (300, 50)), ((77, 10), (77, 90))])
def test_oracle_walk_equals_naive_bresenham(a, b):
    w, h = 640, 100
    v = line_verts([ndc(*a, w, h) + (0.5,), ndc(*b, w, h) + (0.5,)])
    px = np.zeros((h, w), np.uint32)
    st = oracle.render(e.VertexColor(primitives=e.LineList), v, px, None)
    expect = set(naive_bresenham(a[0], a[1], b[0], b[1]))
    got = set((int(x), int(y)) for y, x in np.argwhere(px != 0))
    assert got == expect
    assert st["fragments"] == len(expect)


def test_zero_length_line_emits_nothing():
    # norm = 1 / 0 = inf, frac = -inf (or NaN), z = NaN -> fails passes_z_clip (lines.rs:77-82, :97-99)
    v = line_verts([ndc(5, 5, 640, 100) + (0.5,)] * 2)
    px = np.zeros((100, 640), np.uint32)
    assert oracle.render(e.VertexColor(primitives=e.LineList), v, px, None)["fragments"] == 0
    # without a z clip range the NaN depth passes and the single pixel is emitted
    st = oracle.render(e.VertexColor(primitives=e.LineList, coords=e.CoordinateMode.VULKAN.without_z_clip()), v, px, None)
    assert st["fragments"] == 1  # its colour is NaN -> `as u8` -> 0x00000000


def test_line_triangle_list_and_partial_primitives():
    w, h = 640, 100
    tri = [ndc(20, 10, w, h) + (0.5,), ndc(300, 20, w, h) + (0.5,), ndc(100, 90, w, h) + (0.5,)]
    v = line_verts(tri + tri[:2])  # 5 vertices: one triangle, trailing partial primitive dropped (pipeline.rs:283)
    px = np.zeros((h, w), np.uint32)
    st = oracle.render(e.VertexColor(primitives=e.LineTriangleList), v, px, None)
    assert st["primitives"] == 3  # a-b, b-c, c-a (primitives.rs:56-76)
    expect = set(naive_bresenham(20, 10, 300, 20)) | set(naive_bresenham(300, 20, 100, 90)) | set(naive_bresenham(100, 90, 20, 10))
    assert set((int(x), int(y)) for y, x in np.argwhere(px != 0)) == expect
    # LineList with an odd vertex count drops the last vertex
    px2 = np.zeros((h, w), np.uint32)
    assert oracle.render(e.VertexColor(primitives=e.LineList), v[:3], px2, None)["primitives"] == 1


def test_far_away_end_points_are_cheap_and_clipped():
    w, h = 640, 100
    v = line_verts([(-1e6, -0.3, 0.5), (1e6, 0.4, 0.5)])
    px = np.zeros((h, w), np.uint32)
    st = oracle.render(e.VertexColor(primitives=e.LineList), v, px, None)
    assert st["fragments"] == 640 and (px != 0).sum(axis=0).tolist() == [1] * 640  # one pixel per column, x-major


def _random_lines(n, seed, nasty):
    r = scenes.u01(seed, n * 2 * 8).reshape(n, 2, 8)
    v = np.zeros((n, 2), dtype=e.VERTEX_P4C4)
    wv = 0.5 + 1.5 * r[:, :, 4]
    v["pos"][:, :, 0] = (r[:, :, 0] * 2.6 - 1.3) * wv
    v["pos"][:, :, 1] = (r[:, :, 1] * 2.6 - 1.3) * wv
    v["pos"][:, :, 2] = (r[:, :, 2] * 1.2 - 0.1) * wv
    v["pos"][:, :, 3] = wv
    v["rgba"][:, :, :3] = r[:, :, 5:8]
    v["rgba"][:, :, 3] = 1.0
    if nasty:
        v["pos"][0::7, 1, 3] = -0.5      # w <= 0 -> clamped to 0.0001: huge screen coordinates
        v["pos"][1::11, 0, 3] = 0.0
        v["pos"][2::13, 1, 0] = np.nan
        v["pos"][3::17, 0, 1] = np.inf
        v["pos"][4::19, :, :2] *= 1e6
        v["pos"][5::23, 1] = v["pos"][5::23, 0]  # zero-length
    return v.reshape(-1)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(640, 480), (1920, 64), (333, 217)])
@pytest.mark.parametrize("nasty", [False, True])
@pytest.mark.parametrize("depth,coords", [(e.DepthMode.NONE, "VULKAN"), (e.DepthMode.LESS_WRITE, "VULKAN"), (e.DepthMode.GREATER_WRITE, "OPENGL")])
def test_gpu_random_lines(w, h, nasty, depth, coords):
    from conftest import assert_colour_within_1lsb, assert_depth_bit_exact
    cm = e.CoordinateMode.VULKAN if coords == "VULKAN" else e.CoordinateMode.OPENGL
    v = _random_lines(400, 7 + w, nasty)
    ctx = e.default_context()
    ctx.set_stats(True)
    pipe = e.VertexColor(primitives=e.LineList, depth=depth, coords=cm)
    clear_z = 0.0 if depth.test == "Greater" else 1.0
    px = e.Buffer2d.fill([w, h], 0, dtype=np.uint32)
    z = e.Buffer2d.fill([w, h], clear_z) if depth.uses_depth() else e.Empty()
    pipe.render(v, px, z)
    gs = ctx.get_stats()
    rpx = np.zeros((h, w), np.uint32)
    rz = np.full((h, w), clear_z, np.float32) if depth.uses_depth() else None
    rs = oracle.render(pipe, v, rpx, rz, n_threads=0)
    assert gs["primitives"] == rs["primitives"] == 400
    assert gs["fragments"] == rs["fragments"] and rs["fragments"] > 1000
    if rz is not None:
        assert_depth_bit_exact(z.raw(), rz, "lines depth")
    assert np.array_equal(px.raw() != 0, rpx != 0), "coverage differs"
    assert_colour_within_1lsb(px.raw(), rpx, "lines colour")


@pytest.mark.gpu
def test_gpu_wireframe_teapot():
    """examples/wireframes.rs at 1280x960: LineTriangleList, constant red, no depth."""
    w, h = 1280, 960
    stream = scenes.teapot_stream()
    p = vek.perspective_fov_lh_zo(1.3, w, h, 0.01, 100.0)
    v = vek.mul(vek.identity(), vek.translation_3d((0, 0, 6.0)), vek.rotation_x(0.3), vek.rotation_y(0.5))
    m = vek.mul(vek.translation_3d((0, 0, 0)), vek.rotation_x(np.pi))
    pipe = e.Wireframe(m, v, p)
    ctx = e.default_context()
    ctx.set_stats(True)
    px = e.Buffer2d.fill([w, h], 0, dtype=np.uint32)
    pipe.render(e.Geometry(stream), px, e.Empty())
    gs = ctx.get_stats()
    rpx = np.zeros((h, w), np.uint32)
    rs = oracle.render(pipe, stream, rpx, None, n_threads=0)
    assert gs["primitives"] == rs["primitives"] == 6768 and gs["fragments"] == rs["fragments"] > 100000
    assert np.array_equal(px.raw(), rpx)  # constant colour: bit-exact
    assert int(rpx[rpx != 0][0]) == 0xFFFF0000  # BGRA red


@pytest.mark.gpu
def test_gpu_lines_msaa_and_line_triangle_list():
    from conftest import assert_colour_within_1lsb
    w, h = 800, 600
    v = _random_lines(300, 99, False)
    pipe = e.VertexColor(primitives=e.LineTriangleList, aa=e.AaMode.Msaa(1))
    px = e.Buffer2d.fill([w, h], 0, dtype=np.uint32)
    pipe.render(v, px, e.Empty())
    rpx = np.zeros((h, w), np.uint32)
    rs = oracle.render(pipe, v, rpx, None, n_threads=0)
    assert rs["primitives"] == 600  # 200 collected triangles x 3 lines
    assert np.array_equal(px.raw() != 0, rpx != 0)
    assert_colour_within_1lsb(px.raw(), rpx, "msaa lines")


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [e.LineList, e.LineTriangleList])
def test_gpu_lines_for_immediate_mode_pipeline(kind):
    """primitives.rs:49-104 is generic over the pipeline: a blending (immediate-mode) pipeline draws lines in submission
    order too.  BLEND_TRIS has no transcendental: colour bit-exact."""
    from conftest import assert_depth_bit_exact
    w, h = 800, 300
    v = _random_lines(450, 21, False)
    pipe = lambda: e.BlendTris(primitives=kind, depth=e.DepthMode.LESS_WRITE)
    px, z = e.Buffer2d.fill([w, h], 0xFF000000, dtype=np.uint32), e.Buffer2d.fill([w, h], 1.0)
    pipe().render(v, px, z)
    rpx, rz = np.full((h, w), 0xFF000000, np.uint32), np.full((h, w), 1.0, np.float32)
    rs = oracle.render(pipe(), v, rpx, rz, n_threads=0)
    assert rs["fragments"] > 1000
    assert_depth_bit_exact(z.raw(), rz, "blend lines depth")
    assert np.array_equal(px.raw(), rpx), "blend lines colour"


@pytest.mark.gpu
def test_gpu_phong_wireframe_teapot():
    """The teapot drawn as LineTriangleList through the Phong pipeline (nine varyings, shadow-map sampler): every pipeline
    with a fragment stage renders lines."""
    from conftest import assert_colour_within_1lsb, assert_depth_bit_exact
    w, h, s = 640, 480, 512
    stream, u = scenes.teapot_stream(), scenes.teapot_uniforms(w, h, s)
    shadow = e.Buffer2d.fill([s, s], 1.0)
    e.TeapotShadow(u["shadow_mvp"]).render(stream, e.Empty(), shadow)
    mk = lambda smp: e.Teapot(u["m"], u["v"], u["p"], u["light_pos"], smp, u["light_vp"], u["cam_pos"], primitives=e.LineTriangleList)
    px, z = e.Buffer2d.fill([w, h], 0, dtype=np.uint32), e.Buffer2d.fill([w, h], 1.0)
    mk(shadow.linear().clamped()).render(stream, px, z)
    r_shadow = shadow.raw()
    rpx, rz = np.zeros((h, w), np.uint32), np.full((h, w), 1.0, np.float32)
    rs = oracle.render(mk(e.Sampler(r_shadow, e.abi.TEXEL_F32, e.abi.FILTER_LINEAR).clamped()), stream, rpx, rz, n_threads=0)
    assert rs["fragments"] > 20000
    assert_depth_bit_exact(z.raw(), rz, "phong wireframe depth")
    assert_colour_within_1lsb(px.raw(), rpx, "phong wireframe colour")
