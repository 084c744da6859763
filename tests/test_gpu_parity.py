"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.
Bar (BASELINE.json north_star): coverage and depth bit-exact; colour within 1 LSB per 8-bit channel."""
import numpy as np
import pytest

import euc_b200 as e
from euc_b200 import scenes
from oracle import oracle
from conftest import assert_colour_within_1lsb, assert_depth_bit_exact

pytestmark = pytest.mark.gpu


def run_both(make_pipe, verts, w, h, clear_px=0, clear_z=1.0, indices=None, want_px=True, want_z=True, tex=None, resident=False):
    """make_pipe(sampler_texture) -> pipeline.  Returns (gpu_px, gpu_z, ref_px, ref_z, gpu_stats, ref_stats)."""
    ctx = e.default_context()
    ctx.set_stats(True)
    px = e.Buffer2d.fill([w, h], clear_px, dtype=np.uint32) if want_px else e.Empty()
    z = e.Buffer2d.fill([w, h], float(clear_z), dtype=np.float32) if want_z else e.Empty()
    dev_tex = e.Buffer2d.from_array(tex) if tex is not None else None
    pipe = make_pipe(dev_tex)
    if resident:
        v = e.Geometry(verts, indices)
    else:
        v = e.IndexedVertices(indices, verts) if indices is not None else verts
    pipe.render(v, px, z)
    gstats = ctx.get_stats()
    gpx = px.raw() if want_px else None
    gz = z.raw() if want_z else None
    rpx = np.full((h, w), clear_px, dtype=np.uint32) if want_px else None
    rz = np.full((h, w), clear_z, dtype=np.float32) if want_z else None
    rpipe = make_pipe(tex)
    rv = e.IndexedVertices(indices, verts) if indices is not None else verts
    rstats = oracle.render(rpipe, rv, rpx, rz, n_threads=0)
    return gpx, gz, rpx, rz, gstats, rstats


def test_readme_triangle():
    verts = np.zeros(3, dtype=e.VERTEX_P4C4)
    verts["pos"] = [(-1, -1, 0, 1), (1, -1, 0, 1), (0, 1, 0, 1)]
    verts["rgba"] = [(1, 0, 0, 1), (0, 1, 0, 1), (0, 0, 1, 1)]
    gpx, _, rpx, _, gs, rs = run_both(lambda t: e.VertexColor(), verts, 640, 480, want_z=False)
    assert np.array_equal(gpx != 0, rpx != 0), "coverage mask differs"
    assert_colour_within_1lsb(gpx, rpx, "README triangle")
    assert gs["fragments"] == rs["fragments"] == 153280


def _random_tris(n, seed, w_lo=0.5, w_hi=2.0, size=0.3, nasty=False):
    r = scenes.u01(seed, n * 3 * 8).reshape(n, 3, 8)
    v = np.zeros((n, 3), dtype=e.VERTEX_P4C4)
    cx = r[:, :1, 0] * 2.4 - 1.2
    cy = r[:, :1, 1] * 2.4 - 1.2
    wv = w_lo + r[:, :, 4] * (w_hi - w_lo)
    x = cx + (r[:, :, 2] - 0.5) * size * 2
    y = cy + (r[:, :, 3] - 0.5) * size * 2
    z = r[:, :, 5] * 1.2 - 0.1
    v["pos"][:, :, 0] = x * wv
    v["pos"][:, :, 1] = y * wv
    v["pos"][:, :, 2] = z * wv
    v["pos"][:, :, 3] = wv
    v["rgba"][:, :, :3] = r[:, :, 5:8]
    v["rgba"][:, :, 3] = 0.25 + 0.5 * r[:, :, 6]
    if nasty:
        # degenerate, w <= 0, NaN / Inf vertices, huge coordinates
        v["pos"][0::17, 1] = v["pos"][0::17, 0]
        v["pos"][1::19, 2, 3] = -0.5
        v["pos"][2::23, 0, 3] = 0.0
        v["pos"][3::29, 1, 0] = np.nan
        v["pos"][4::31, 2, 1] = np.inf
        v["pos"][5::37, 0, 0] = 1e30
        v["pos"][6::41, :, :2] *= 50.0
    return v.reshape(-1)


@pytest.mark.parametrize("w,h", [(640, 480), (1920, 64), (333, 217), (2048, 40)])
@pytest.mark.parametrize("size,n", [(0.05, 3000), (0.6, 200)])
def test_random_triangles_blend(w, h, size, n):
    verts = _random_tris(n, 0xABC0 + w + n, size=size)
    gpx, gz, rpx, rz, gs, rs = run_both(lambda t: e.BlendTris(), verts, w, h, clear_px=0xFF000000)
    assert_depth_bit_exact(gz, rz, f"random tris {w}x{h}")
    assert_colour_within_1lsb(gpx, rpx, f"random tris {w}x{h}")
    assert gs["fragments"] == rs["fragments"]
    assert gs["primitives"] == rs["primitives"] == n


def _axis_aligned_tris(n, seed, w, h):
    """Right triangles with exactly horizontal and vertical edges on whole / half pixel positions: two of the three
    barycentric weights have a zero, or a cancellation-sized (down to denormal), slope along x or y."""
    r = scenes.u01(seed, n * 8).reshape(n, 8)
    v = np.zeros((n, 3), dtype=e.VERTEX_P4C4)
    px0 = np.floor(r[:, 0] * w) + (r[:, 6] < 0.5) * 0.5
    py0 = np.floor(r[:, 1] * h) + (r[:, 7] < 0.5) * 0.5
    ex = np.floor(1 + r[:, 2] * 40) * np.where(r[:, 3] < 0.5, -1, 1)
    ey = np.floor(1 + r[:, 4] * 40) * np.where(r[:, 5] < 0.5, -1, 1)
    sx = np.stack([px0, px0 + ex, px0], axis=1)
    sy = np.stack([py0, py0, py0 + ey], axis=1)
    wv = np.where(r[:, 5:6] < 0.25, 1.0, 0.5 + r[:, 2:5])  # a quarter of them without perspective
    v["pos"][:, :, 0] = (sx / w * 2 - 1) * wv
    v["pos"][:, :, 1] = (sy / h * 2 - 1) * wv
    v["pos"][:, :, 2] = (0.1 + 0.8 * r[:, 3:4]) * wv
    v["pos"][:, :, 3] = wv
    v["rgba"][:, :, :3] = r[:, None, 5:8]
    v["rgba"][:, :, 3] = 0.5
    return v.astype(e.VERTEX_P4C4).reshape(-1)


@pytest.mark.parametrize("w,h,seed", [(640, 480, 1), (1024, 96, 2), (4096, 48, 3)])
def test_axis_aligned_edges(w, h, seed):
    verts = _axis_aligned_tris(4000, 0xA715 + seed, w, h)
    gpx, gz, rpx, rz, gs, rs = run_both(lambda t: e.BlendTris(), verts, w, h, clear_px=0xFF000000)
    assert_depth_bit_exact(gz, rz, f"axis aligned {w}x{h}")
    assert_colour_within_1lsb(gpx, rpx, f"axis aligned {w}x{h}")
    assert gs["fragments"] == rs["fragments"]


@pytest.mark.parametrize("cull", [e.CullMode.NONE, e.CullMode.Back, e.CullMode.Front])
@pytest.mark.parametrize("coords", ["VULKAN", "OPENGL", "noclip"])
def test_nasty_triangles_modes(cull, coords):
    cm = {"VULKAN": e.CoordinateMode.VULKAN, "OPENGL": e.CoordinateMode.OPENGL, "noclip": e.CoordinateMode.VULKAN.without_z_clip()}[coords]
    verts = _random_tris(1500, 0x5EED, w_lo=-0.3, w_hi=1.5, size=0.2, nasty=True)
    gpx, gz, rpx, rz, gs, rs = run_both(lambda t: e.BlendTris(cull=cull, coords=cm), verts, 800, 600, clear_px=0xFF102030)
    assert_depth_bit_exact(gz, rz, "nasty")
    assert_colour_within_1lsb(gpx, rpx, "nasty")
    assert gs["fragments"] == rs["fragments"]


@pytest.mark.parametrize("depth", [e.DepthMode.NONE, e.DepthMode.LESS_PASS, e.DepthMode.GREATER_WRITE, e.DepthMode("Equal", True)])
def test_depth_modes(depth):
    verts = _random_tris(800, 0xD0, size=0.25)
    gpx, gz, rpx, rz, gs, rs = run_both(lambda t: e.BlendTris(depth=depth), verts, 640, 480, clear_px=0xFF000000, clear_z=0.5,
                                        want_z=depth.uses_depth())
    if depth.uses_depth():
        assert_depth_bit_exact(gz, rz, str(depth))
    assert_colour_within_1lsb(gpx, rpx, str(depth))
    assert gs["fragments"] == rs["fragments"]


def _teapot_both(w, h, shadow_size, aa=None):
    ctx = e.default_context()
    ctx.set_stats(True)
    stream = scenes.teapot_stream()
    u = scenes.teapot_uniforms(w, h, shadow_size)
    geom = e.Geometry(stream)
    shadow = e.Buffer2d.fill([shadow_size, shadow_size], 1.0)
    color = e.Buffer2d.fill([w, h], 0, dtype=np.uint32)
    depth = e.Buffer2d.fill([w, h], 1.0)
    e.TeapotShadow(u["shadow_mvp"]).render(geom, e.Empty(), shadow)
    fr_shadow = ctx.get_stats()["fragments"]
    e.Teapot(u["m"], u["v"], u["p"], u["light_pos"], shadow.linear().clamped(), u["light_vp"], u["cam_pos"], aa=aa).render(geom, color, depth)
    fr_color = ctx.get_stats()["fragments"]
    g = (shadow.raw(), color.raw(), depth.raw(), fr_shadow, fr_color)
    rshadow = np.full((shadow_size, shadow_size), 1.0, dtype=np.float32)
    rcolor = np.zeros((h, w), dtype=np.uint32)
    rdepth = np.full((h, w), 1.0, dtype=np.float32)
    s1 = oracle.render(e.TeapotShadow(u["shadow_mvp"]), stream, None, rshadow, n_threads=0)
    s2 = oracle.render(e.Teapot(u["m"], u["v"], u["p"], u["light_pos"], e.Sampler(rshadow, e.abi.TEXEL_F32, e.abi.FILTER_LINEAR).clamped(),
                                u["light_vp"], u["cam_pos"], aa=aa), stream, rcolor, rdepth, n_threads=0)
    return g, (rshadow, rcolor, rdepth, s1["fragments"], s2["fragments"])


def test_teapot_c1():
    g, r = _teapot_both(640, 480, 512)
    assert_depth_bit_exact(g[0], r[0], "shadow map")
    assert_depth_bit_exact(g[2], r[2], "depth")
    assert_colour_within_1lsb(g[1], r[1], "teapot colour")
    assert np.array_equal(g[1] != 0, r[1] != 0)
    assert g[3] == r[3] and g[4] == r[4]
    assert r[4] > 20000  # the teapot is actually on screen


@pytest.mark.parametrize("level", [1, 2])
def test_teapot_msaa(level):
    g, r = _teapot_both(1280, 720, 1024, aa=e.AaMode.Msaa(level))
    assert_depth_bit_exact(g[0], r[0], "shadow map")
    assert_depth_bit_exact(g[2], r[2], "depth")
    assert_colour_within_1lsb(g[1], r[1], f"teapot colour msaa {level}")
    assert g[4] == r[4]


@pytest.mark.parametrize("frame,wrap", [(0, "tiled"), (250, "tiled"), (500, "mirrored"), (100, "clamped"), (100, "none")])
@pytest.mark.parametrize("filt", ["linear", "nearest"])
def test_textured_cube(frame, wrap, filt):
    w, h = 1920, 1080
    verts, idx = scenes.cube_geometry(uv_scale=3.0 if wrap != "none" else 1.0)
    tex = scenes.rust_texture()
    mvp = scenes.cube_mvp(frame, w, h)

    def make(t):
        if isinstance(t, e.Buffer2d):
            s = t.linear() if filt == "linear" else t.nearest()
        else:
            s = e.Sampler(t.view(np.uint32).reshape(t.shape[0], t.shape[1]), e.abi.TEXEL_RGBA8_TO_F32,
                          e.abi.FILTER_LINEAR if filt == "linear" else e.abi.FILTER_NEAREST)
        s = {"tiled": s.tiled, "mirrored": s.mirrored, "clamped": s.clamped, "none": lambda: s}[wrap]()
        return e.Cube(mvp, s)

    gpx, _, rpx, _, gs, rs = run_both(make, verts, w, h, clear_px=180, indices=idx, want_z=False, tex=tex)
    assert np.array_equal(gpx != 180, rpx != 180), "coverage (texels are opaque: a covered pixel is never the clear value 180)"
    assert_colour_within_1lsb(gpx, rpx, "cube")
    assert gs["fragments"] == rs["fragments"] and rs["fragments"] > 100000


def test_blend_scene_c4_small():
    w, h = 1280, 720
    verts, idx = scenes.blend_tris(1 << 13, w, h)
    gpx, gz, rpx, rz, gs, rs = run_both(lambda t: e.BlendTris(), verts, w, h, clear_px=0xFF000000, indices=idx, resident=True)
    assert_depth_bit_exact(gz, rz, "C4 small")
    assert_colour_within_1lsb(gpx, rpx, "C4 small")
    assert gs["fragments"] == rs["fragments"]


def test_voxel_icons_batch():
    n = 6
    verts, idx, draws, ubs = scenes.voxel_icon_batch(n)
    geom = e.Geometry(verts, idx)
    color = e.Buffer2d.fill([256, 256], 0, dtype=np.uint32, layers=n)
    depth = e.Buffer2d.fill([256, 256], 1.0, layers=n)
    pipe = e.VoxelIcon(np.eye(4), scenes.VOXEL_LIGHT_DIR)
    pipe.render_batch(geom, draws, ubs, color, depth)
    gpx, gz = color.raw(), depth.raw()
    for k in range(n):
        rpx = np.zeros((256, 256), dtype=np.uint32)
        rz = np.full((256, 256), 1.0, dtype=np.float32)
        first, count, base, _ = draws[k]
        rs = oracle.render(e.VoxelIcon(scenes.voxel_icon_mvp(k), scenes.VOXEL_LIGHT_DIR), e.IndexedVertices(idx, verts), rpx, rz,
                           draw=(first, count, base))
        assert rs["fragments"] > 1000, "icon is on screen"
        assert_depth_bit_exact(gz[k], rz, f"icon {k}")
        assert_colour_within_1lsb(gpx[k], rpx, f"icon {k}")


def test_many_layers_large_batch():
    """A batch the size of the benchmark's icon batches: 1200 layers = 307200 tiles, two thirds of them empty, about 75 tiles per
    warp of the persistent tile kernel.  Eight distinct icons repeated: the small batch (2048 tiles) is checked against the
    oracle, and every layer of the large batch must equal the layer of its icon in the small one bit for bit - with the clears
    fused into the render and without."""
    nd, reps = 8, 150
    verts, idx, draws, ubs = scenes.voxel_icon_batch(nd)
    geom = e.Geometry(verts, idx)
    pipe = e.VoxelIcon(np.eye(4), scenes.VOXEL_LIGHT_DIR)
    small_c = e.Buffer2d.fill([256, 256], 0, dtype=np.uint32, layers=nd)
    small_z = e.Buffer2d.fill([256, 256], 1.0, layers=nd)
    pipe.render_batch(geom, draws, ubs, small_c, small_z)
    sc, sz = small_c.raw(), small_z.raw()
    for k in (0, 5):
        rpx, rz = np.zeros((256, 256), dtype=np.uint32), np.full((256, 256), 1.0, dtype=np.float32)
        first, count, base, _ = draws[k]
        oracle.render(e.VoxelIcon(scenes.voxel_icon_mvp(k), scenes.VOXEL_LIGHT_DIR), e.IndexedVertices(idx, verts), rpx, rz, draw=(first, count, base))
        assert_depth_bit_exact(sz[k], rz, f"icon {k}")
        assert_colour_within_1lsb(sc[k], rpx, f"icon {k}")
    n = nd * reps
    ub = len(ubs) // nd
    big_draws = [(int(draws[k % nd][0]), int(draws[k % nd][1]), int(draws[k % nd][2]), k) for k in range(n)]
    big_ubs = b"".join(ubs[(k % nd) * ub:(k % nd + 1) * ub] for k in range(n))
    for fused in (False, True):
        # the fused form starts from garbage: every tile of every layer, empty or not, must be written by the render
        big_c = e.Buffer2d.fill([256, 256], 0x12345678 if fused else 0, dtype=np.uint32, layers=n)
        big_z = e.Buffer2d.fill([256, 256], 0.25 if fused else 1.0, layers=n)
        pipe.render_batch(geom, big_draws, big_ubs, big_c, big_z, clear=(0, 1.0) if fused else None)
        bc, bz = big_c.raw(), big_z.raw()
        for k in range(n):
            assert np.array_equal(bc[k], sc[k % nd]), f"layer {k} colour (fused clear: {fused})"
            assert np.array_equal(bz[k].view(np.uint32), sz[k % nd].view(np.uint32)), f"layer {k} depth (fused clear: {fused})"


def test_row_bands_equal_full_render():
    """Multi-GPU partitioning property: rendering row bands separately gives the rows of the full render."""
    w, h = 1280, 720
    verts, idx = scenes.blend_tris(1 << 12, w, h, seed=77)
    geom = e.Geometry(verts, idx)
    full_c = e.Buffer2d.fill([w, h], 0xFF000000, dtype=np.uint32)
    full_z = e.Buffer2d.fill([w, h], 1.0)
    e.BlendTris().render(geom, full_c, full_z)
    part_c = e.Buffer2d.fill([w, h], 0xFF000000, dtype=np.uint32)
    part_z = e.Buffer2d.fill([w, h], 1.0)
    for r0, r1 in [(0, 240), (240, 400), (400, 720)]:
        e.BlendTris().render(geom, part_c, part_z, rows=(r0, r1))
    assert np.array_equal(full_c.raw(), part_c.raw())
    assert np.array_equal(full_z.raw().view(np.uint32), part_z.raw().view(np.uint32))


def test_quirks_and_errors():
    verts = _random_tris(10, 1)
    # h < group_rows: the reference spawns zero threads and renders nothing (pipeline.rs:330,337)
    px = e.Buffer2d.fill([32, 32], 7, dtype=np.uint32)
    z = e.Buffer2d.fill([32, 32], 1.0)
    e.BlendTris().render(verts, px, z)
    assert (px.raw() == 7).all() and (z.raw() == 1.0).all()
    # size mismatch (pipeline.rs:262-266)
    with pytest.raises(e.EucError) as ei:
        e.BlendTris().render(verts, e.Buffer2d.fill([64, 640], 0, dtype=np.uint32), e.Buffer2d.fill([64, 641], 1.0))
    assert ei.value.code == e.abi.E_SIZE_MISMATCH
    # index out of range (index.rs:53 slice panic)
    with pytest.raises(e.EucError) as ei:
        e.BlendTris().render(e.IndexedVertices([0, 1, 999], verts), e.Buffer2d.fill([64, 640], 0, dtype=np.uint32), e.Buffer2d.fill([64, 640], 1.0))
    assert ei.value.code == e.abi.E_OUT_OF_BOUNDS
    # a render after an error still works (tile counters were restored)
    px = e.Buffer2d.fill([64, 640], 0, dtype=np.uint32)
    z = e.Buffer2d.fill([64, 640], 1.0)
    e.BlendTris().render(verts, px, z)
    rpx, rz = np.zeros((640, 64), dtype=np.uint32), np.full((640, 64), 1.0, dtype=np.float32)
    oracle.render(e.BlendTris(), verts, rpx, rz)
    assert_depth_bit_exact(z.raw(), rz, "after error")
    # width > 20000: the reference divides by zero
    with pytest.raises(e.EucError) as ei:
        e.BlendTris().render(verts, e.Buffer2d.fill([20001, 4], 0, dtype=np.uint32), e.Buffer2d.fill([20001, 4], 1.0))
    assert ei.value.code == e.abi.E_UNSUPPORTED
    # trailing partial primitive is dropped (pipeline.rs:283)
    px = e.Buffer2d.fill([64, 640], 0, dtype=np.uint32)
    z = e.Buffer2d.fill([64, 640], 1.0)
    e.BlendTris().render(verts[:29], px, z)
    rpx, rz = np.zeros((640, 64), dtype=np.uint32), np.full((640, 64), 1.0, dtype=np.float32)
    oracle.render(e.BlendTris(), verts[:29], rpx, rz)
    assert_depth_bit_exact(z.raw(), rz, "partial primitive")


def _slivers_and_giants(n, seed, w, h):
    """Stress for the per-row interval test of the tile kernel: near-horizontal and near-vertical slivers that cross the
    whole target (tiny weight slopes, long accumulation chains), triangles far larger than the screen (weights that
    nearly cancel at the visible rows), and sub-pixel triangles; perspective on half of them."""
    r = scenes.u01(seed, n * 12).reshape(n, 12)
    v = np.zeros((n, 3), dtype=e.VERTEX_P4C4)
    kind = (r[:, 0] * 4).astype(int)
    x0, y0 = r[:, 1] * 2.4 - 1.2, r[:, 2] * 2.4 - 1.2
    px, py = 2.0 / w, 2.0 / h
    ndc = np.zeros((n, 3, 2))
    # 0: horizontal sliver   1: vertical sliver   2: giant   3: sub-pixel
    ndc[:, 0] = np.stack([x0, y0], 1)
    hs = kind == 0
    ndc[hs, 1] = np.stack([x0 + 2.5 * (r[:, 3] - 0.5) * 2, y0 + (r[:, 4] - 0.5) * 6 * py], 1)[hs]
    ndc[hs, 2] = np.stack([x0 + 2.5 * (r[:, 5] - 0.5) * 2, y0 + (r[:, 6] - 0.5) * 6 * py], 1)[hs]
    vs = kind == 1
    ndc[vs, 1] = np.stack([x0 + (r[:, 3] - 0.5) * 6 * px, y0 + 2.5 * (r[:, 4] - 0.5) * 2], 1)[vs]
    ndc[vs, 2] = np.stack([x0 + (r[:, 5] - 0.5) * 6 * px, y0 + 2.5 * (r[:, 6] - 0.5) * 2], 1)[vs]
    gs = kind == 2
    ndc[gs, 1] = np.stack([x0 + (r[:, 3] - 0.5) * 80, y0 + (r[:, 4] - 0.5) * 80], 1)[gs]
    ndc[gs, 2] = np.stack([x0 + (r[:, 5] - 0.5) * 80, y0 + (r[:, 6] - 0.5) * 80], 1)[gs]
    ss = kind == 3
    ndc[ss, 1] = np.stack([x0 + (r[:, 3] - 0.5) * 3 * px, y0 + (r[:, 4] - 0.5) * 3 * py], 1)[ss]
    ndc[ss, 2] = np.stack([x0 + (r[:, 5] - 0.5) * 3 * px, y0 + (r[:, 6] - 0.5) * 3 * py], 1)[ss]
    wv = np.where(r[:, 7:8] < 0.5, 1.0, 0.4 + 2.0 * r[:, 8:11])
    v["pos"][:, :, 0] = ndc[:, :, 0] * wv
    v["pos"][:, :, 1] = ndc[:, :, 1] * wv
    v["pos"][:, :, 2] = (0.05 + 0.9 * r[:, 11:12]) * wv
    v["pos"][:, :, 3] = wv
    v["rgba"][:, :, :3] = r[:, None, 8:11]
    v["rgba"][:, :, 3] = 0.4
    return v.reshape(-1)


@pytest.mark.parametrize("w,h,seed", [(4096, 64, 11), (640, 480, 12), (1920, 1080, 13)])
def test_slivers_giants_and_subpixel_triangles(w, h, seed):
    verts = _slivers_and_giants(1500, 0x511 + seed, w, h)
    gpx, gz, rpx, rz, gs, rs = run_both(lambda t: e.BlendTris(), verts, w, h, clear_px=0xFF000000)
    assert_depth_bit_exact(gz, rz, f"slivers/giants {w}x{h}")
    assert_colour_within_1lsb(gpx, rpx, f"slivers/giants {w}x{h}")
    assert gs["fragments"] == rs["fragments"] > 1000


def test_many_small_and_a_few_screen_sized_primitives():
    """More than 65536 primitives switch the CTA-cooperative bin walk off; screen-sized primitives among them then go through
    the flattened per-warp walk (thousands of tiles per primitive), medium ones through the chunked per-thread walk."""
    w, h = 1024, 512
    small = _random_tris(66000, 0xB16, size=0.01).reshape(-1, 3)
    big = np.zeros((4, 3), dtype=e.VERTEX_P4C4)
    big["pos"][0] = [(-1, -1, 0.9, 1), (1, -1, 0.9, 1), (0, 1, 0.9, 1)]
    big["pos"][1] = [(-1.5, -1.2, 0.5, 1), (1.5, -0.9, 0.5, 1), (-1.4, 1.3, 0.5, 1)]
    big["pos"][2] = [(-0.3, -0.3, 0.2, 1), (0.4, -0.2, 0.2, 1), (0.0, 0.5, 0.2, 1)]   # ~150 tiles
    big["pos"][3] = [(-0.1, -0.1, 0.1, 1), (0.1, -0.1, 0.1, 1), (0.0, 0.1, 0.1, 1)]   # ~20 tiles
    big["rgba"][:, :, :3] = [[(1, 0, 0)], [(0, 1, 0)], [(0, 0, 1)], [(1, 1, 0)]]
    big["rgba"][:, :, 3] = 0.5
    verts = np.concatenate([small[:20000], big[:2], small[20000:40000], big[2:], small[40000:]]).reshape(-1)
    gpx, gz, rpx, rz, gs, rs = run_both(lambda t: e.BlendTris(), verts, w, h, clear_px=0xFF000000)
    assert_depth_bit_exact(gz, rz, "small + screen-sized")
    assert_colour_within_1lsb(gpx, rpx, "small + screen-sized")
    assert gs["fragments"] == rs["fragments"]
    assert gs["primitives"] == 66004


def test_pair_list_overflow_relaunch():
    """A fresh context sizes the (tile, primitive) list optimistically; full-screen triangles overflow that guess, the
    device flags it, and the host re-launches fill + raster after growing the list.  Results must be unaffected."""
    ctx = e.Context(0)
    w, h = 1024, 768
    n = 300
    r = scenes.u01(99, n * 3 * 8).reshape(n, 3, 8)
    v = np.zeros((n, 3), dtype=e.VERTEX_P4C4)
    v["pos"][:, :, 0] = np.array([-1.5, 1.5, 0.0])[None, :] + (r[:, :, 0] - 0.5) * 0.2
    v["pos"][:, :, 1] = np.array([-1.5, -1.5, 1.5])[None, :] + (r[:, :, 1] - 0.5) * 0.2
    v["pos"][:, :, 2] = r[:, :, 2]
    v["pos"][:, :, 3] = 1.0
    v["rgba"][:, :, :3] = r[:, :, 3:6]
    v["rgba"][:, :, 3] = 0.5
    v = v.reshape(-1)
    px = e.Buffer2d.fill([w, h], 0xFF000000, dtype=np.uint32, ctx=ctx)
    z = e.Buffer2d.fill([w, h], 1.0, ctx=ctx)
    ctx.set_stats(True)
    e.BlendTris().render(v, px, z)
    st = ctx.get_stats()
    assert st["binned_pairs"] > (1 << 16), "scene must overflow the initial list guess"
    rpx, rz = np.full((h, w), 0xFF000000, np.uint32), np.full((h, w), 1.0, np.float32)
    rs = oracle.render(e.BlendTris(), v, rpx, rz, n_threads=0)
    assert st["fragments"] == rs["fragments"]
    assert_depth_bit_exact(z.raw(), rz, "overflow relaunch")
    assert_colour_within_1lsb(px.raw(), rpx, "overflow relaunch")
    del px, z
    ctx.close()


def test_long_tile_lists_sorted():
    """Many small triangles stacked on the same tile: list lengths beyond the register sort (128) and the
    shared-memory sort (2048) paths; blending makes the result order-sensitive."""
    w, h = 640, 64
    for n in (100, 700, 3000):
        r = scenes.u01(1234 + n, n * 3 * 8).reshape(n, 3, 8)
        v = np.zeros((n, 3), dtype=e.VERTEX_P4C4)
        v["pos"][:, :, 0] = -0.9 + r[:, :, 0] * 0.08
        v["pos"][:, :, 1] = 0.2 + r[:, :, 1] * 0.6
        v["pos"][:, :, 2] = 0.9 - 0.8 * (np.arange(n)[:, None] / n) + r[:, :, 2] * 1e-4   # later triangles are nearer: all pass
        v["pos"][:, :, 3] = 1.0
        v["rgba"][:, :, :3] = r[:, :, 3:6]
        v["rgba"][:, :, 3] = 0.3
        gpx, gz, rpx, rz, gs, rs = run_both(lambda t: e.BlendTris(), v.reshape(-1), w, h, clear_px=0xFF000000)
        assert_depth_bit_exact(gz, rz, f"long list {n}")
        assert_colour_within_1lsb(gpx, rpx, f"long list {n}")
        assert gs["fragments"] == rs["fragments"] > n


@pytest.mark.parametrize("kind", ["blend", "phong"])
def test_mirrored_row_bands_fill_every_mirror(kind):
    """Fused gather (euc_render_geom_rows_mirrored): two "ranks" (here: two buffers on one GPU) each render one row band
    and mirror it into the other's framebuffer; afterwards both hold the complete frame, equal to a full render.
    Covers the raster write-back (immediate pipelines), the resolve kernel (deferred pipelines) and empty tiles."""
    w, h = 1280, 720
    bands = [(0, 368), (368, 720)]
    if kind == "blend":
        verts, idx = scenes.blend_tris(1 << 11, w, h, seed=5)   # sparse: plenty of tiles without primitives
        geom = e.Geometry(verts, idx)
        make = lambda: e.BlendTris()
        clear = 0xFF000000
        extra = {}
    else:
        stream, u = scenes.teapot_stream(), scenes.teapot_uniforms(w, h, 512)
        geom = e.Geometry(stream)
        shadow = e.Buffer2d.fill([512, 512], 1.0)
        e.TeapotShadow(u["shadow_mvp"]).render(geom, e.Empty(), shadow)
        make = lambda: e.Teapot(u["m"], u["v"], u["p"], u["light_pos"], shadow.linear().clamped(), u["light_vp"], u["cam_pos"])
        clear = 0x00102030
    full_c = e.Buffer2d.fill([w, h], clear, dtype=np.uint32)
    full_z = e.Buffer2d.fill([w, h], 1.0)
    make().render(geom, full_c, full_z)
    cols = [e.Buffer2d.fill([w, h], 0xDEADBEEF, dtype=np.uint32) for _ in bands]   # poison: every pixel must be written
    zs = [e.Buffer2d.fill([w, h], 1.0) for _ in bands]
    for k, (r0, r1) in enumerate(bands):
        cols[k].clear_rows(clear, r0, r1)
        make().render(geom, cols[k], zs[k], rows=(r0, r1), mirrors=[cols[1 - k]])
    want = full_c.raw()
    for k in range(2):
        assert np.array_equal(cols[k].raw(), want), f"buffer {k} is not the complete frame"


@pytest.mark.parametrize("level", [1, 3, 6])
@pytest.mark.parametrize("w,h", [(640, 480), (333, 217), (2500, 600)])
def test_msaa_immediate_mode_levels(level, w, h):
    """euc's coarse-shading MSAA in the immediate (blending) path, including the maximum level and odd target sizes
    (band height 20000*2^level/w changes with the level: pipeline.rs:329)."""
    verts = _random_tris(300, 0x77 + level, size=0.35)
    gpx, gz, rpx, rz, gs, rs = run_both(lambda t: e.BlendTris(aa=e.AaMode.Msaa(level)), verts, w, h, clear_px=0xFF000000)
    assert_depth_bit_exact(gz, rz, f"msaa {level}")
    assert_colour_within_1lsb(gpx, rpx, f"msaa {level}")
    assert gs["fragments"] == rs["fragments"]
    if h >= 20000 * (1 << level) // w:   # otherwise h / group_rows == 0: the reference renders nothing (pipeline.rs:330,337)
        assert rs["fragments"] > 1000
    else:
        assert rs["fragments"] == 0 and (gpx == 0xFF000000).all()


def test_msaa_vertex_color_deferred_odd_size():
    verts = _random_tris(200, 0x99, size=0.4)
    gpx, gz, rpx, rz, gs, rs = run_both(lambda t: e.VertexColor(aa=e.AaMode.Msaa(2), depth=e.DepthMode.LESS_WRITE, cull=e.CullMode.NONE),
                                        verts, 1001, 333, clear_px=0)
    assert_depth_bit_exact(gz, rz, "deferred msaa")
    assert_colour_within_1lsb(gpx, rpx, "deferred msaa")


def test_batch_draws_share_and_split_layers():
    """euc_render_batch: draws 0 and 1 blend into layer 0 in submission order, draw 2 goes to layer 1, each with its own
    uniform block and vertex range."""
    n = 3
    verts, idx, draws, ubs = scenes.voxel_icon_batch(n)
    draws = [(draws[0][0], draws[0][1], draws[0][2], 0), (draws[1][0], draws[1][1], draws[1][2], 0), (draws[2][0], draws[2][1], draws[2][2], 1)]
    geom = e.Geometry(verts, idx)
    color = e.Buffer2d.fill([256, 256], 0, dtype=np.uint32, layers=2)
    depth = e.Buffer2d.fill([256, 256], 1.0, layers=2)
    # alpha-blended voxels make the order of draws 0 and 1 visible
    e.VoxelIcon(np.eye(4), scenes.VOXEL_LIGHT_DIR, depth=e.DepthMode.NONE).render_batch(geom, draws, ubs, color, depth)
    gpx = color.raw()
    iv = e.IndexedVertices(idx, verts)
    ref = np.zeros((2, 256, 256), np.uint32)
    for k, (first, count, base, layer) in enumerate(draws):
        oracle.render(e.VoxelIcon(scenes.voxel_icon_mvp(k), scenes.VOXEL_LIGHT_DIR, depth=e.DepthMode.NONE), iv, ref[layer], None, draw=(first, count, base))
    for layer in range(2):
        assert_colour_within_1lsb(gpx[layer], ref[layer], f"layer {layer}")
    assert (ref[0] != 0).sum() > 5000


def test_row_bands_with_msaa_equal_full_render():
    w, h = 1280, 720
    stream, u = scenes.teapot_stream(), scenes.teapot_uniforms(w, h, 512)
    geom = e.Geometry(stream)
    shadow = e.Buffer2d.fill([512, 512], 1.0)
    e.TeapotShadow(u["shadow_mvp"]).render(geom, e.Empty(), shadow)
    mk = lambda: e.Teapot(u["m"], u["v"], u["p"], u["light_pos"], shadow.linear().clamped(), u["light_vp"], u["cam_pos"], aa=e.AaMode.Msaa(1))
    full_c, full_z = e.Buffer2d.fill([w, h], 0, dtype=np.uint32), e.Buffer2d.fill([w, h], 1.0)
    mk().render(geom, full_c, full_z)
    part_c, part_z = e.Buffer2d.fill([w, h], 0, dtype=np.uint32), e.Buffer2d.fill([w, h], 1.0)
    for r0, r1 in [(0, 304), (304, 512), (512, 720)]:   # cuts through euc's 31-row bands
        mk().render(geom, part_c, part_z, rows=(r0, r1))
    assert np.array_equal(full_c.raw(), part_c.raw())
    assert np.array_equal(full_z.raw().view(np.uint32), part_z.raw().view(np.uint32))


def test_depth_pass_without_write_and_pixel_only():
    verts = _random_tris(500, 0x31, size=0.3)
    # LESS_PASS: depth tested against a constant buffer, never written; blending sees every passing fragment in order
    gpx, gz, rpx, rz, gs, rs = run_both(lambda t: e.BlendTris(depth=e.DepthMode.LESS_PASS), verts, 640, 480, clear_px=0xFF000000, clear_z=0.6)
    assert_depth_bit_exact(gz, rz, "LESS_PASS leaves depth alone")
    assert (gz == np.float32(0.6)).all()
    assert_colour_within_1lsb(gpx, rpx, "LESS_PASS")
    assert gs["fragments"] == rs["fragments"]
