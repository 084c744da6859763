"""The tile kernel only visits the (primitive, 8-pixel row segment) pairs that its lane-mask stage lets through
(euc_b200/csrc/kernels.cuh, `lane_mask`): per tile row the x-interval on which all three weights can be non-negative,
each bound linear in y with an error budget folded in.  A segment that is wrongly rejected would lose pixels, so the
mask must be a SUPERSET of the pixels euc's accumulated chain accepts.  This test restates the device arithmetic in
numpy float32 (same operations, same order; fma through float64) from the oracle's own setup values and checks the
superset property against the oracle's exact coverage, triangle by triangle, on ordinary, hostile and degenerate
triangles.  CPU only: it pins the mathematics of the bound, the GPU parity tests pin the CUDA code."""
import numpy as np
import pytest

import euc_b200 as e
from euc_b200 import scenes
from oracle import oracle

F = np.float32
TILE = 16
BIG = F(3.0e38)


def fma(a, b, c):
    return F(np.float64(a) * np.float64(b) + np.float64(c))


def fmaxf(a, b):  # CUDA fmaxf: the non-NaN operand
    if np.isnan(a):
        return b
    if np.isnan(b):
        return a
    return a if a > b else b


def fminf(a, b):
    if np.isnan(a):
        return b
    if np.isnan(b):
        return a
    return a if a < b else b


def lane_mask(o, dx, dy, x0, x1, y0, y1, tile_x0, tile_y0):
    """Restatement of the triangle branch of `lane_mask`: returns {(row, half)} the kernel would visit in this tile."""
    ra, rb = max(y0, tile_y0) - tile_y0, min(y1, tile_y0 + TILE) - tile_y0
    out = set()
    if not (y1 > tile_y0 and y0 < tile_y0 + TILE and rb > ra and x1 > x0):
        return out
    seg0 = x0 < tile_x0 + 8 and x1 > tile_x0
    seg1 = x0 < tile_x0 + 16 and x1 > tile_x0 + 8
    a0, a1, a2 = o
    d0, d1, d2 = dx
    b0, b1, b2 = dy
    with np.errstate(all="ignore"):
        au, bu, du = F(F(a2 - a0) - a1), F(F(b2 - b0) - b1), F(F(d2 - d0) - d1)
        x1f, ymaxf = F(x1), F(tile_y0 + TILE)
        kerr = F(F(x1 - x0 + 12) * F(1.1920929e-07))
        s = [F(F(abs(a) + F(abs(b) * ymaxf)) + F(abs(d) * x1f)) for a, b, d in ((a0, b0, d0), (a1, b1, d1), (a2, b2, d2))]
        m0, m1 = F(kerr * s[0]), F(kerr * s[1])
        mu = F(F(2.0) * F(kerr * F(F(s[0] + s[1]) + s[2])))
        slack = F(F(0.01) + F(x1f * F(1e-5)))

        def edge(a, b, d, m):
            i = F(F(1.0) / d) if d != 0 else F(np.copysign(np.inf, d))
            if not (abs(i) < F(1.0e30)):
                i = F(np.copysign(1.0e30, d))
            tp, tq = F(-b * i), F(F(-m - a) * i)
            lower = i > 0
            return (tp if lower else F(0), F(tq - slack) if lower else -BIG, F(0) if lower else tp, BIG if lower else F(tq + slack))
        e0, e1, eu = edge(a0, b0, d0, m0), edge(a1, b1, d1, m1), edge(au, bu, du, mu)
        xa0, xb0 = F(max(tile_x0, x0)), F(min(tile_x0 + 7, x1 - 1))
        xa1, xb1 = F(max(tile_x0 + 8, x0)), F(min(tile_x0 + 15, x1 - 1))
        if not (fminf(fminf(s[0], s[1]), s[2]) > F(1.0e-18)):
            for r in range(ra, rb):
                if seg0:
                    out.add((r, 0))
                if seg1:
                    out.add((r, 1))
            return out
        yr = F(tile_y0 + ra)
        for r in range(ra, rb):
            lo = fmaxf(fmaxf(fma(yr, e0[0], e0[1]), fma(yr, e1[0], e1[1])), fma(yr, eu[0], eu[1]))
            hi = fminf(fminf(fma(yr, e0[2], e0[3]), fma(yr, e1[2], e1[3])), fma(yr, eu[2], eu[3]))
            ilo, ihi = np.ceil(lo), np.floor(hi)
            if seg0 and not (fmaxf(xa0, ilo) > fminf(xb0, ihi)):
                out.add((r, 0))
            if seg1 and not (fmaxf(xa1, ilo) > fminf(xb1, ihi)):
                out.add((r, 1))
            yr = F(yr + F(1.0))
    return out


def _triangles(kind, n, seed, w, h):
    if kind == "slivers":
        import test_gpu_parity as t
        return t._slivers_and_giants(n, seed, w, h).reshape(-1, 3)
    if kind == "axis":
        import test_gpu_parity as t
        return t._axis_aligned_tris(n, seed, w, h).reshape(-1, 3)
    import test_gpu_parity as t
    return t._random_tris(n, seed, w_lo=-0.2 if kind == "nasty" else 0.5, w_hi=2.0, size=0.25 if kind != "small" else 0.03, nasty=kind == "nasty").reshape(-1, 3)


@pytest.mark.parametrize("kind,n,w,h", [("random", 250, 640, 96), ("small", 400, 4096, 48), ("nasty", 400, 800, 96), ("slivers", 300, 4096, 64),
                                        ("axis", 300, 640, 96)])
def test_lane_mask_is_a_superset_of_the_exact_coverage(kind, n, w, h):
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    tris = _triangles(kind, n, 0x1A5E + n, w, h)
    tris["rgba"] = 1.0
    checked = visited = covered_segments = 0
    for tri in tris:
        px = np.zeros((h, w), np.uint32)
        st = oracle.render(e.VertexColor(), tri.copy(), px, None, dump_setup=True)
        d = st["setup"][0]
        if d.culled or st["fragments"] == 0:
            continue
        x0, y0, x1, y1 = d.bounds_min[0], d.bounds_min[1], d.bounds_max[0], d.bounds_max[1]
        o, dx, dy = [F(v) for v in d.w_hom_origin], [F(v) for v in d.w_hom_dx], [F(v) for v in d.w_hom_dy]
        ys, xs = np.nonzero(px)
        segs = set(zip((ys // TILE).tolist(), (xs // TILE).tolist(), (ys % TILE).tolist(), ((xs % TILE) // 8).tolist()))
        covered_segments += len(segs)
        masks = {}
        for ty, tx, r, half in segs:
            if (ty, tx) not in masks:
                masks[(ty, tx)] = lane_mask(o, dx, dy, x0, x1, y0, y1, tx * TILE, ty * TILE)
                visited += len(masks[(ty, tx)])
            assert (r, half) in masks[(ty, tx)], f"{kind}: a covered segment is rejected: tile ({tx},{ty}) row {r} half {half}, bounds {x0,x1,y0,y1}"
        checked += 1
    assert checked > n // 10 and covered_segments > 0
    # the mask is not vacuous either: it visits few segments beyond the covered ones
    assert visited <= 3 * covered_segments + 64 * checked
