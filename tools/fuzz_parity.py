"""Randomised differential soak: random scenes (sizes, depth / cull / coordinate modes, MSAA levels, pipelines, hostile
vertices, fused clears) rendered by the CUDA path and by the CPU oracle; depth and fragment counts must be bit-exact,
colour within 1 LSB.  Open-ended run: python tools/fuzz_parity.py [seconds] [seed].  tests/test_fuzz_slice.py runs a fixed
200-scene slice of the same generators (triangle_scene / sampler_or_line_scene) in the `-m gpu` suite."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import euc_b200 as e
from euc_b200 import scenes
from oracle import oracle



def triangle_scene(k, seed0):
    """One random triangle scene through both paths; returns the fragment count.  Raises AssertionError on a mismatch."""
    ctx = e.default_context(); ctx.set_stats(True)
    rng = np.random.default_rng(seed0 * 100003 + k)
    w = int(rng.choice([64, 100, 333, 640, 1000, 1920, 2500, 4096])); h = int(rng.choice([48, 64, 217, 480, 720]))
    n = int(rng.choice([1, 7, 60, 400, 3000]))
    size = float(rng.choice([0.004, 0.03, 0.15, 0.7, 3.0]))
    r = rng.random((n, 3, 8), dtype=np.float32)
    v = np.zeros((n, 3), dtype=e.VERTEX_P4C4)
    wv = np.where(rng.random((n, 1)) < 0.3, 1.0, 0.3 + 2.0 * r[:, :, 4]).astype(np.float32)
    if rng.random() < 0.3:
        wv = wv - 0.5  # some w <= 0
    v["pos"][:, :, 0] = ((r[:, :1, 0] * 2.6 - 1.3) + (r[:, :, 2] - 0.5) * size * 2) * wv
    v["pos"][:, :, 1] = ((r[:, :1, 1] * 2.6 - 1.3) + (r[:, :, 3] - 0.5) * size * 2) * wv
    v["pos"][:, :, 2] = (r[:, :, 5] * 1.3 - 0.15) * wv
    v["pos"][:, :, 3] = wv
    v["rgba"][:, :, :3] = r[:, :, 5:8]; v["rgba"][:, :, 3] = 0.2 + 0.6 * r[:, :, 6]
    if rng.random() < 0.25:  # snapped to pixel / half-pixel positions: edges through pixel centres
        v["pos"][:, :, 0] = (np.round((v["pos"][:, :, 0] / wv * 0.5 + 0.5) * w * 2) / 2 / w * 2 - 1) * wv
        v["pos"][:, :, 1] = (np.round((v["pos"][:, :, 1] / wv * 0.5 + 0.5) * h * 2) / 2 / h * 2 - 1) * wv
    if rng.random() < 0.2:
        bad = rng.integers(0, n, size=max(1, n // 20))
        v["pos"][bad, rng.integers(0, 3), rng.integers(0, 4)] = rng.choice([np.nan, np.inf, -np.inf, 1e30, -1e30, 0.0])
    verts = v.reshape(-1)
    depth = [e.DepthMode.LESS_WRITE, e.DepthMode.LESS_PASS, e.DepthMode.GREATER_WRITE, e.DepthMode.NONE, e.DepthMode("Equal", True)][int(rng.integers(0, 5))]
    cull = [e.CullMode.NONE, e.CullMode.Back, e.CullMode.Front][int(rng.integers(0, 3))]
    cm = [e.CoordinateMode.VULKAN, e.CoordinateMode.OPENGL, e.CoordinateMode.VULKAN.without_z_clip()][int(rng.integers(0, 3))]
    aa_level = int(rng.choice([0, 0, 1, 2, 3]))
    kind = int(rng.integers(0, 2))
    kw = dict(depth=depth, cull=cull, coords=cm)
    if aa_level:
        kw["aa"] = e.AaMode.Msaa(aa_level)
    if w > 20000 * (1 << aa_level) or h < (20000 << aa_level) // w:
        pass  # the "renders nothing" quirk is a valid scene too
    mk = (lambda: e.BlendTris(**kw)) if kind == 0 else (lambda: e.VertexColor(**kw))
    cpx, cz = int(rng.integers(0, 2**32)), float(rng.random())
    fused = rng.random() < 0.5
    px = e.Buffer2d.fill([w, h], 0 if fused else cpx, dtype=np.uint32)
    use_z = depth.uses_depth()
    z = e.Buffer2d.fill([w, h], 0.25 if fused else cz) if use_z else e.Empty()
    mk().render(verts, px, z, clear=(cpx, cz if use_z else None) if fused else None)
    gs = ctx.get_stats()
    gpx, gz = px.raw(), (z.raw() if use_z else None)
    rpx = np.full((h, w), cpx, np.uint32); rz = np.full((h, w), cz, np.float32) if use_z else None
    rs = oracle.render(mk(), verts, rpx, rz, n_threads=0)
    what = f"scene {k} seed {seed0}: {w}x{h} n={n} size={size} depth={depth} cull={cull} aa={aa_level} pipe={kind} fused={fused}"
    assert gs["fragments"] == rs["fragments"], what + f" fragments {gs['fragments']} != {rs['fragments']}"
    if use_z:
        assert np.array_equal(gz.view(np.uint32), rz.view(np.uint32)), what + " depth differs"
    dmax = np.abs(gpx.view(np.uint8).astype(np.int16) - rpx.view(np.uint8).astype(np.int16)).max()
    assert dmax <= 1, what + f" colour differs by {dmax}"
    return rs["fragments"]


# ---- samplers (textured cube, random texture / filter / wrap / uv scale / pose) and lines -------------------------------
def sampler_or_line_scene(k, seed0):
    """One random textured-cube or line-list scene through both paths; returns "cube" or "lines"."""
    ctx = e.default_context(); ctx.set_stats(True)
    rng = np.random.default_rng(seed0 * 100003 + k)
    w = int(rng.choice([320, 640, 1000, 1920])); h = int(rng.choice([200, 480, 1080]))
    if rng.random() < 0.6:
        th, tw = int(rng.integers(1, 200)), int(rng.integers(1, 200))
        tex = rng.integers(0, 256, size=(th, tw, 4), dtype=np.uint8)
        verts, idx = scenes.cube_geometry(uv_scale=float(rng.choice([1.0, 3.0, -2.5, 0.37, 17.0])))
        mvp = scenes.cube_mvp(int(rng.integers(0, 2000)), w, h)
        filt, wrap = str(rng.choice(["linear", "nearest"])), str(rng.choice(["tiled", "mirrored", "clamped", "none"]))

        def make(t):
            if isinstance(t, e.Buffer2d):
                sm = t.linear() if filt == "linear" else t.nearest()
            else:
                sm = e.Sampler(t.view(np.uint32).reshape(t.shape[0], t.shape[1]), e.abi.TEXEL_RGBA8_TO_F32,
                               e.abi.FILTER_LINEAR if filt == "linear" else e.abi.FILTER_NEAREST)
            return e.Cube(mvp, {"tiled": sm.tiled, "mirrored": sm.mirrored, "clamped": sm.clamped, "none": lambda: sm}[wrap]())
        px = e.Buffer2d.fill([w, h], 0, dtype=np.uint32)
        make(e.Buffer2d.from_array(tex)).render(e.IndexedVertices(idx, verts), px, e.Empty(), clear=(180, None))
        gfr = ctx.get_stats()["fragments"]
        rpx = np.full((h, w), 180, np.uint32)
        rs = oracle.render(make(tex), e.IndexedVertices(idx, verts), rpx, None, n_threads=0)
        what = f"cube scene {k} seed {seed0}: {w}x{h} tex {tw}x{th} {filt} {wrap}"
        which = "cube"
    else:
        n = int(rng.choice([2, 20, 400]))
        v = np.zeros(2 * n, dtype=e.VERTEX_P4C4)
        wv = np.where(rng.random(2 * n) < 0.5, 1.0, 0.2 + 2.0 * rng.random(2 * n)).astype(np.float32)
        v["pos"][:, 0] = (rng.random(2 * n) * 3 - 1.5) * wv; v["pos"][:, 1] = (rng.random(2 * n) * 3 - 1.5) * wv
        v["pos"][:, 2] = (rng.random(2 * n) * 1.2 - 0.1) * wv; v["pos"][:, 3] = wv
        v["rgba"] = rng.random((2 * n, 4), dtype=np.float32)
        depth = [e.DepthMode.LESS_WRITE, e.DepthMode.NONE][int(rng.integers(0, 2))]
        pipe = lambda: e.VertexColor(primitives=e.LineList, depth=depth)
        px = e.Buffer2d.fill([w, h], 0, dtype=np.uint32)
        z = e.Buffer2d.fill([w, h], 1.0) if depth.uses_depth() else e.Empty()
        pipe().render(v, px, z)
        gfr = ctx.get_stats()["fragments"]
        rpx = np.zeros((h, w), np.uint32); rz = np.full((h, w), 1.0, np.float32) if depth.uses_depth() else None
        rs = oracle.render(pipe(), v, rpx, rz, n_threads=0)
        what = f"line scene {k} seed {seed0}: {w}x{h} n={n} {depth}"
        if rz is not None:
            assert np.array_equal(z.raw().view(np.uint32), rz.view(np.uint32)), what + " depth differs"
        which = "lines"
    assert gfr == rs["fragments"], what + f" fragments {gfr} != {rs['fragments']}"
    dmax = np.abs(px.raw().view(np.uint8).astype(np.int16) - rpx.view(np.uint8).astype(np.int16)).max()
    assert dmax <= 1, what + f" colour differs by {dmax}"
    return which


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    t_end = time.time() + budget
    n_scenes = n_frag = k = 0
    while time.time() < t_end:
        k += 1
        n_frag += triangle_scene(k, seed0); n_scenes += 1
    print(f"fuzz ok: {n_scenes} triangle scenes, {n_frag} fragments, seed {seed0}, {budget:.0f} s")
    t_end = time.time() + budget / 3
    counts = {"cube": 0, "lines": 0}
    while time.time() < t_end:
        k += 1
        counts[sampler_or_line_scene(k, seed0)] += 1
    print(f"fuzz ok: {counts['cube']} textured-cube scenes, {counts['lines']} line scenes")
