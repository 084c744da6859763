"""Scene generators for the five BASELINE configs.  Everything is produced on the host once, as
numpy arrays, and handed bit-identically to the device path and to the oracle.

  C1  teapot, 512² shadow pass + Phong pass, 640×480                     (benches/teapot.rs:144-209)
  C2  textured cube, bilinear + tiled sampler, 1920×1080                 (examples/texture_mapping.rs)
  C3  teapot, 2048² shadow pass + Phong pass, 3840×2160, MSAA level 1
  C4  2^20 indexed triangles, random depth + alpha blend, 3840×2160      (defined by this build)
  C5  4096 voxel icons 256×256, depth + blend                            (defined by this build)
"""
import os

import numpy as np

from . import vek
from .pipelines import VERTEX_P4C4, VERTEX_P4UV, VERTEX_PN, VERTEX_VOXEL

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
f32 = np.float32


# ---- PRNG: splitmix64, u01 = (next >> 40) * 2^-24 (exact in f32) -----------------------------------------
def splitmix64(seed, n):
    """First n outputs of splitmix64 seeded with `seed`, vectorised."""
    with np.errstate(over="ignore"):
        g = np.uint64(0x9E3779B97F4A7C15)
        z = np.uint64(seed) + g * np.arange(1, n + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def u01(seed, n):
    return ((splitmix64(seed, n) >> np.uint64(40)).astype(np.float64) * 2.0 ** -24).astype(f32)


# ---- C1 / C3: teapot --------------------------------------------------------------------------------------
def teapot_stream():
    """`model.vertices()` of wavefront 0.2: three face-vertices per `f` line, file order (6768 vertices)."""
    d = np.load(os.path.join(_DATA, "teapot.npz"))
    faces = d["faces"].reshape(-1, 2)
    out = np.zeros(faces.shape[0], dtype=VERTEX_PN)
    out["pos"] = d["positions"][faces[:, 0]]
    out["normal"] = d["normals"][faces[:, 1]]
    return out


def teapot_uniforms(w, h, shadow_size):
    """benches/teapot.rs:154-179, :188, :195-203."""
    ori = (f32(-0.55), f32(-0.25))
    dist = f32(4.5)
    teapot_pos = np.zeros(3, dtype=f32)
    light_pos = np.array([-8.0, 5.0, -5.0], dtype=f32)
    light_p = vek.perspective_fov_lh_zo(0.75, shadow_size, shadow_size, 0.1, 100.0)
    light_v = vek.look_at_lh(light_pos, -teapot_pos, np.array([0, 1, 0], dtype=f32))
    light_vp = vek.mul(light_p, light_v)
    p = vek.perspective_fov_lh_zo(1.3, w, h, 0.01, 100.0)
    v = vek.mul(vek.identity(), vek.translation_3d((0.0, 0.0, dist)))
    m = vek.mul(vek.translation_3d(-teapot_pos), vek.rotation_x(np.pi), vek.rotation_x(ori[0]), vek.rotation_y(ori[1]))
    cam_pos = vek.mul_point(vek.inverted(v), (0.0, 0.0, 0.0))
    return dict(m=m, v=v, p=p, light_pos=light_pos, light_vp=light_vp, cam_pos=cam_pos, shadow_mvp=vek.mul(light_vp, m))


# ---- C2: textured cube ------------------------------------------------------------------------------------
_CUBE_POS = [(-1, -1, 1), (-1, 1, 1), (1, 1, 1), (1, -1, 1), (-1, -1, -1), (-1, 1, -1), (1, 1, -1), (1, -1, -1),
             (-1, 1, 1), (-1, 1, -1), (1, 1, -1), (1, 1, 1), (-1, -1, 1), (-1, -1, -1), (1, -1, -1), (1, -1, 1),
             (1, -1, 1), (1, -1, -1), (1, 1, -1), (1, 1, 1), (-1, -1, 1), (-1, -1, -1), (-1, 1, -1), (-1, 1, 1)]
_CUBE_UV = [(0, 1), (0, 0), (1, 0), (1, 1), (0, 0), (0, 1), (1, 1), (1, 0), (0, 0), (0, 1), (1, 1), (1, 0),
            (0, 0), (0, 1), (1, 1), (1, 0), (1, 1), (1, 0), (0, 0), (0, 1), (0, 1), (0, 0), (1, 0), (1, 1)]
_CUBE_IDX = [0, 3, 1, 1, 3, 2, 4, 5, 7, 5, 6, 7, 8, 11, 9, 9, 11, 10, 12, 13, 15, 13, 14, 15,
             16, 17, 19, 17, 18, 19, 20, 23, 21, 21, 23, 22]


def cube_geometry(uv_scale=1.0):
    """examples/texture_mapping.rs:44-108 (positions, uvs) and :146-147 (indices).  C2 multiplies the uvs by 3.0
    on the host so that `.tiled()` wrapping is exercised (the example itself uses plain `.linear()`)."""
    v = np.zeros(24, dtype=VERTEX_P4UV)
    v["pos"][:, :3] = np.array(_CUBE_POS, dtype=f32)
    v["pos"][:, 3] = 1.0
    v["uv"] = np.array(_CUBE_UV, dtype=f32) * f32(uv_scale)
    return v, np.array(_CUBE_IDX, dtype=np.uint32)


def cube_mvp(i, w, h):
    """examples/texture_mapping.rs:127-133 at frame i."""
    p = vek.perspective_fov_rh_no(1.4, w, h, 0.01, 100.0)
    v = vek.mul(vek.translation_3d((0.0, 0.0, -2.0)), vek.scaling_3d(0.6), vek.rotation_x(0.6))
    fi = f32(i)
    m = vek.mul(vek.rotation_x(f32(np.sin(fi * f32(0.004))) * f32(0.4)), vek.rotation_y((fi * f32(0.0008)) * f32(4.0)),
                vek.rotation_z(f32(np.cos(fi * f32(0.006))) * f32(0.4)))
    return vek.mul(p, v, m)


def rust_texture():
    """rust.png as (899, 860, 4) uint8 RGBA (`image::open(..).to_rgba8()`, examples/texture_mapping.rs:111)."""
    return np.load(os.path.join(_DATA, "rust_rgba.npz"))["rgba"]


# ---- C4: random blended triangles -------------------------------------------------------------------------
def blend_tris(n_quads=1 << 19, w=3840, h=2160, seed=0xE0C40004, size_px=(2.0, 6.0)):
    """n_quads quads -> 4*n_quads vertices {clip pos, rgba}, 6*n_quads u32 indices (0,1,2, 2,1,3 + 4q).
    Quad q: centre uniform in NDC [-1,1]^2, rotation uniform, half-edge s ~ U[size_px] pixels, z ~ U(0.05,0.95) per quad
    + U(-0.01,0.01) per vertex, w ~ U[0.5,2] per vertex (clip = ndc*w), rgb ~ U[0,1], a ~ U[0.25,0.75]."""
    per = 5 + 4 * 6
    r = u01(seed, n_quads * per).reshape(n_quads, per)
    cx, cy = r[:, 0] * f32(2) - f32(1), r[:, 1] * f32(2) - f32(1)
    th = r[:, 2] * f32(2 * np.pi)
    s = f32(size_px[0]) + r[:, 3] * f32(size_px[1] - size_px[0])
    zq = f32(0.05) + r[:, 4] * f32(0.9)
    c, sn = np.cos(th).astype(f32), np.sin(th).astype(f32)
    corners = np.array([(-1, -1), (1, -1), (-1, 1), (1, 1)], dtype=f32)
    verts = np.zeros((n_quads, 4), dtype=VERTEX_P4C4)
    for k in range(4):
        pv = r[:, 5 + 6 * k: 11 + 6 * k]
        ox, oy = corners[k, 0] * s, corners[k, 1] * s
        px = (c * ox - sn * oy) * f32(2.0 / w)
        py = (sn * ox + c * oy) * f32(2.0 / h)
        z = zq + (pv[:, 0] * f32(0.02) - f32(0.01))
        wv = f32(0.5) + pv[:, 1] * f32(1.5)
        verts["pos"][:, k, 0] = (cx + px) * wv
        verts["pos"][:, k, 1] = (cy + py) * wv
        verts["pos"][:, k, 2] = z * wv
        verts["pos"][:, k, 3] = wv
        verts["rgba"][:, k, 0:3] = pv[:, 2:5]
        verts["rgba"][:, k, 3] = f32(0.25) + pv[:, 5] * f32(0.5)
    idx = (np.array([0, 1, 2, 2, 1, 3], dtype=np.uint32)[None, :] + (np.arange(n_quads, dtype=np.uint32) * 4)[:, None])
    return verts.reshape(-1), idx.reshape(-1)


# ---- C5: voxel icons --------------------------------------------------------------------------------------
_GRID = (12, 32, 6)
_FACES = [  # (normal, 4 corner offsets, counter-clockwise seen from outside, y up)
    ((1, 0, 0), [(1, 0, 0), (1, 1, 0), (1, 1, 1), (1, 0, 1)]),
    ((-1, 0, 0), [(0, 0, 0), (0, 0, 1), (0, 1, 1), (0, 1, 0)]),
    ((0, 1, 0), [(0, 1, 0), (0, 1, 1), (1, 1, 1), (1, 1, 0)]),
    ((0, -1, 0), [(0, 0, 0), (1, 0, 0), (1, 0, 1), (0, 0, 1)]),
    ((0, 0, 1), [(0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]),
    ((0, 0, -1), [(0, 0, 0), (0, 1, 0), (1, 1, 0), (1, 0, 0)]),
]


def voxel_icon_mesh(icon, seed_base=0xE0C50000):
    """Voxel grid 12x32x6; occupancy = inside a seeded capsule AND u01 < 0.85; colour from a 16-entry seeded
    palette; 1/16 of the voxels have alpha 0.5.  Mesh = 2 triangles per exposed face (indexed quads)."""
    gx, gy, gz = _GRID
    n = gx * gy * gz
    r = u01(seed_base + icon, 8 + 64 + 3 * n)
    # capsule: axis from a to b inside the grid, radius rad
    a = np.array([gx * (0.3 + 0.4 * r[0]), gy * (0.12 + 0.15 * r[1]), gz * (0.35 + 0.3 * r[2])])
    b = np.array([gx * (0.3 + 0.4 * r[3]), gy * (0.73 + 0.15 * r[4]), gz * (0.35 + 0.3 * r[5])])
    rad = 2.2 + 2.3 * r[6]
    palette = (r[8:8 + 64].reshape(16, 4) * 255.0).astype(np.uint8)
    palette[:, 3] = 255
    rv = r[72:].reshape(n, 3)
    ix, iy, iz = np.meshgrid(np.arange(gx), np.arange(gy), np.arange(gz), indexing="ij")
    cpos = np.stack([ix + 0.5, iy + 0.5, iz + 0.5], axis=-1).reshape(n, 3)
    ab = b - a
    t = np.clip(((cpos - a) @ ab) / (ab @ ab), 0.0, 1.0)
    dist = np.linalg.norm(cpos - (a + t[:, None] * ab), axis=1)
    occ = ((dist <= rad) & (rv[:, 0] < 0.85)).reshape(gx, gy, gz)
    colour_idx = (rv[:, 1] * 16).astype(np.int32).clip(0, 15).reshape(gx, gy, gz)
    glow = (rv[:, 2] < (1.0 / 16.0)).reshape(gx, gy, gz)
    pad = np.zeros((gx + 2, gy + 2, gz + 2), dtype=bool)
    pad[1:-1, 1:-1, 1:-1] = occ
    verts, idx = [], []
    half = np.array([gx, gy, gz], dtype=np.float64) * 0.5
    for nrm, corners in _FACES:
        nb = pad[1 + nrm[0]: 1 + nrm[0] + gx, 1 + nrm[1]: 1 + nrm[1] + gy, 1 + nrm[2]: 1 + nrm[2] + gz]
        exposed = np.argwhere(occ & ~nb)
        if exposed.size == 0:
            continue
        q = np.zeros((exposed.shape[0], 4), dtype=VERTEX_VOXEL)
        for k, off in enumerate(corners):
            q["pos"][:, k, :] = (exposed + np.array(off) - half).astype(f32)
        q["normal"][:] = np.array(nrm, dtype=f32)
        col = palette[colour_idx[exposed[:, 0], exposed[:, 1], exposed[:, 2]]].copy()
        col[glow[exposed[:, 0], exposed[:, 1], exposed[:, 2]], 3] = 128
        q["rgba"][:] = col[:, None, :]
        base = sum(v.shape[0] for v in verts)
        verts.append(q.reshape(-1))
        qi = np.arange(exposed.shape[0], dtype=np.uint32) * 4 + base
        idx.append((qi[:, None] + np.array([0, 1, 2, 0, 2, 3], dtype=np.uint32)[None, :]).reshape(-1))
    if not verts:
        return np.zeros(0, dtype=VERTEX_VOXEL), np.zeros(0, dtype=np.uint32)
    return np.concatenate(verts), np.concatenate(idx)


def voxel_icon_mvp(icon):
    """ortho-fit x rot_x(-0.6) x rot_y(0.785 + 0.01*icon): the rotated model fits NDC [-0.9,0.9]^2, depth 0.1..0.9."""
    rad = f32(0.5 * np.sqrt(sum(g * g for g in _GRID)))  # bounding-sphere radius of the grid: depth stays in 0.1..0.9
    s = f32(0.9) / f32(13.0)                             # the capsule (not the whole grid) fills the frame
    fit = vek.mul(vek.translation_3d((0.0, 0.0, 0.5)), vek.scaling_3d((s, s, f32(0.4) / rad)))
    return vek.mul(fit, vek.rotation_x(-0.6), vek.rotation_y(f32(0.785) + f32(0.01) * f32(icon)))


VOXEL_LIGHT_DIR = (np.array([-0.35, 0.6, -0.72], dtype=np.float64) / np.linalg.norm([-0.35, 0.6, -0.72])).astype(f32)


def voxel_icon_batch(n_icons, first_icon=0):
    """Concatenated geometry for icons [first_icon, first_icon+n): vertices, indices, draws (first, count,
    base_vertex, layer), uniform blocks (mvp + light_dir per icon)."""
    vs, is_, draws, ubs = [], [], [], []
    vbase = ibase = 0
    for k in range(n_icons):
        v, i = voxel_icon_mesh(first_icon + k)
        vs.append(v)
        is_.append(i)
        draws.append((ibase, i.size, vbase, k))
        vbase += v.size
        ibase += i.size
        mvp = voxel_icon_mvp(first_icon + k)
        ubs.append(np.ascontiguousarray(mvp.T).tobytes() + np.append(VOXEL_LIGHT_DIR, f32(0)).astype(f32).tobytes())
    return np.concatenate(vs), np.concatenate(is_), np.asarray(draws, dtype=np.int64).reshape(-1, 4), b"".join(ubs)
