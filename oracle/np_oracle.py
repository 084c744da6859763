"""Independent numpy-float32 restatement of euc's triangle path, written from the Rust source (NOT from the C++
oracle) to catch transcription errors in either.  Scalar np.float32 arithmetic in Python loops: small scenes only.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (no rustc here): this and oracle/euc_oracle.hpp are two independent
readings of the same source text.

Follows src/pipeline.rs:304-366 (bands), :519-578 (depth test / write, fragment, blend) and
src/rasterizer/triangles.rs:27-304.  Handles TriangleList with a 4-float varying (rgba) — enough to pin coverage,
depth and interpolation; shader-stage parity of the other pipelines is checked against the C++ oracle only.
"""
import math

import numpy as np

F = np.float32
EPS = F(1.1920929e-07)


def as_usize(x):
    x = float(x)
    if math.isnan(x) or x <= 0.0:
        return 0
    if x >= 18446744073709551616.0:
        return (1 << 64) - 1
    return int(x)


def fmin(a, b):  # f32::min: the non-NaN operand wins
    if math.isnan(a):
        return b
    if math.isnan(b):
        return a
    return a if a < b else b


def fmax(a, b):
    if math.isnan(a):
        return b
    if math.isnan(b):
        return a
    return a if a > b else b


def clampu(v, lo, hi):
    return max(lo, min(v, hi))


def cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def sub(a, b):
    return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]


def add(a, b):
    return [a[0] + b[0], a[1] + b[1], a[2] + b[2]]


def dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def scale(a, s):
    return [a[0] * s, a[1] * s, a[2] * s]


def matmul(a, b):
    return [[a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j] for j in range(3)] for i in range(3)]


def matvec(m, v):
    return [m[i][0] * v[0] + m[i][1] * v[1] + m[i][2] * v[2] for i in range(3)]


def as_u8(x):
    x = float(x)
    if math.isnan(x) or x <= 0.0:
        return 0
    return 255 if x >= 255.0 else int(x)


def blend_vertex_color(old, c):  # examples/triangle.rs:23-25
    b = [as_u8(e * F(255.0)) for e in c]
    return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24)


def blend_src_over(old, n):  # BASELINE config 4 (see euc_oracle.cpp BlendTris)
    a = n[3]
    ia = F(1.0) - a
    out = []
    for i in range(3):
        o = F((int(old) >> (8 * i)) & 0xFF)
        c = (n[i] * F(255.0)) * a + o * ia
        out.append(as_u8(fmin(fmax(c, F(0.0)), F(255.0))))
    return out[0] | (out[1] << 8) | (out[2] << 16) | (255 << 24)


def render(clip, rgba, pixel, depth, blend, cull="Back", depth_test=None, depth_write=False, y_up=False,
           z_clip=(0.0, 1.0), record=None):
    """clip: (n,4) f32 vertex-stage outputs, rgba: (n,4) f32 varyings (n multiple of 3).  pixel: (h,w) uint32 or None;
    depth: (h,w) f32 or None.  `record`, if a list, receives (x, y, tri, z_bits) per emitted fragment."""
    np.seterr(all="ignore")
    tgt = pixel if pixel is not None else depth
    H, W = tgt.shape
    group_rows = 20000 // max(W, 1)  # pipeline.rs:329 (msaa level 0)
    if H // group_rows == 0:         # needed_threads == 0
        return
    for row_start in range(0, H, group_rows):  # pipeline.rs:340-349
        tmin = (0, row_start)
        tmax = (W, min(row_start + group_rows, H))
        _band(clip, rgba, pixel, depth, blend, cull, depth_test, depth_write, y_up, z_clip, (W, H), tmin, tmax, record)


def _band(clip, rgba, pixel, depth, blend, cull, depth_test, depth_write, y_up, z_clip, size, tmin, tmax, record):
    flip = (F(1.0), F(-1.0) if y_up else F(1.0))
    sx, sy = F(size[0]), F(size[1])
    to_ndc = [[F(2.0) / sx, F(0.0), F(-1.0)], [F(0.0), F(-2.0) / sy, F(1.0)], [F(0.0), F(0.0), F(1.0)]]

    def zok(z):
        return z_clip is None or (F(z_clip[0]) <= z <= F(z_clip[1]))

    for t in range(clip.shape[0] // 3):
        hom = [[F(clip[3 * t + i][0]) * flip[0], F(clip[3 * t + i][1]) * flip[1], F(clip[3 * t + i][2]), F(clip[3 * t + i][3])] for i in range(3)]
        out = [[F(e) for e in rgba[3 * t + i]] for i in range(3)]
        euc = [[v[0] / v[3], v[1] / v[3], v[2] / v[3]] for v in hom]
        winding = cross(sub(euc[1], euc[0]), sub(euc[2], euc[0]))[2]
        if cull != "None" and winding * (F(1.0) if cull == "Back" else F(-1.0)) < 0.0:
            continue
        if winding >= 0.0:
            hom, euc, out = hom[::-1], euc[::-1], out[::-1]
        a, b, c4 = hom
        c = [c4[0], c4[1], c4[3]]
        ca = sub([a[0], a[1], a[3]], c)
        cb = sub([b[0], b[1], b[3]], c)
        n = cross(ca, cb)
        if dot(n, n) > 0.0:
            rec_det = F(1.0) / fmin(dot(n, c), -EPS)
        else:
            rec_det = F(1.0)
        c2w = matmul([scale(cross(cb, c), rec_det), scale(cross(c, ca), rec_det), scale(n, rec_det)], to_ndc)
        scr = [[sx * (e[0] * F(0.5) + F(0.5)), sy * (e[1] * F(-0.5) + F(0.5))] for e in euc]
        bmin = [clampu(as_usize(fmin(fmin(scr[0][k], scr[1][k]), scr[2][k]) + F(0.0)), tmin[k], tmax[k]) for k in range(2)]
        bmax = [clampu(as_usize(fmax(fmax(scr[0][k], scr[1][k]), scr[2][k]) + F(1.0)), tmin[k], tmax[k]) for k in range(2)]
        origin = matvec(c2w, [F(0.0), F(0.0), F(1.0)])
        k1000 = F(1.0) / F(1000.0)
        wdx = scale(sub(matvec(c2w, [F(1000.0), F(0.0), F(1.0)]), origin), k1000)
        wdy = scale(sub(matvec(c2w, [F(0.0), F(1000.0), F(1.0)]), origin), k1000)
        min_y = fmin(fmin(scr[0][1], scr[1][1]), scr[2][1])
        if scr[0][1] == min_y:
            by = [scr[0], scr[1], scr[2]] if scr[1][1] < scr[2][1] else [scr[0], scr[2], scr[1]]
        elif scr[1][1] == min_y:
            by = [scr[1], scr[0], scr[2]] if scr[0][1] < scr[2][1] else [scr[1], scr[2], scr[0]]
        else:
            by = [scr[2], scr[0], scr[1]] if scr[0][1] < scr[1][1] else [scr[2], scr[1], scr[0]]
        nvc = all(zok(e[2]) for e in euc)
        zs = [hom[0][2], hom[1][2], hom[2][2]]
        ext = (bmax[0] - bmin[0]) * (bmax[1] - bmin[1])
        for y in range(bmin[1], bmax[1]):
            yf = F(y)
            if ext < 128:
                r0, r1 = bmin[0], bmax[0]
            else:
                A, B, Cc = by
                ac = A[0] + ((yf - A[1]) / (Cc[1] - A[1])) * (Cc[0] - A[0])
                if yf < B[1]:
                    e = A[0] + ((yf - A[1]) / (B[1] - A[1])) * (B[0] - A[0])
                else:
                    e = B[0] + ((yf - B[1]) / (Cc[1] - B[1])) * (Cc[0] - B[0])
                lo, hi = fmin(e, ac), fmax(e, ac)
                fl, ce = F(np.floor(lo)), F(np.ceil(hi))
                r0 = as_usize(fl) if (fl >= F(bmin[0]) and fl < F(bmax[0])) else bmin[0]
                r1 = as_usize(ce) if (ce >= F(bmin[0]) and ce < F(bmax[0])) else bmax[0]
            w = add(add(origin, scale(wdy, yf)), scale(wdx, F(r0)))
            for x in range(r0, r1):
                wu = [w[0], w[1], w[2] - w[0] - w[1]]
                if wu[0] >= 0.0 and wu[1] >= 0.0 and wu[2] >= 0.0:
                    z = dot(zs, wu)
                    ok = nvc or zok(z)
                    if ok and depth_test is not None:
                        old = F(depth[y, x])
                        ok = (z < old) if depth_test == "Less" else ((z == old) if depth_test == "Equal" else (z > old))
                    if ok:
                        if record is not None:
                            record.append((x, y, t, int(np.float32(z).view(np.uint32))))
                        if depth_write:
                            depth[y, x] = z
                        if pixel is not None:
                            wh = add(add(origin, scale(wdy, yf)), scale(wdx, F(x)))
                            wub = [wh[0], wh[1], wh[2] - wh[0] - wh[1]]
                            r = F(1.0) / wh[2]
                            ww = [e * r for e in wub]
                            frag = [out[0][k] * ww[0] + out[1][k] * ww[1] + out[2][k] * ww[2] for k in range(4)]
                            pixel[y, x] = blend(int(pixel[y, x]), frag)
                w = add(w, wdx)
