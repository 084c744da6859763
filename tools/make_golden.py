"""Generates tests/golden/golden.npz from the CPU oracle at the BASELINE sizes.

The reference ships no golden vectors and cannot be run here (no rustc), so these are ORACLE-generated regression
pins (PARITY UNPINNED, see oracle/euc_oracle.hpp): the `not gpu` suite checks the oracle still reproduces them at the
sizes it finishes quickly, and the `gpu` suite compares the CUDA path against them at the full BASELINE sizes, where
re-running the CPU path inside a test would take too long.   python tools/make_golden.py
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import euc_b200 as e  # noqa: E402
from euc_b200 import scenes  # noqa: E402
from oracle import oracle  # noqa: E402


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def teapot(w, h, s, msaa):
    stream, u = scenes.teapot_stream(), scenes.teapot_uniforms(w, h, s)
    shadow = np.full((s, s), 1.0, np.float32)
    color = np.zeros((h, w), np.uint32)
    depth = np.full((h, w), 1.0, np.float32)
    f1 = oracle.render(e.TeapotShadow(u["shadow_mvp"]), stream, None, shadow, n_threads=0)["fragments"]
    aa = e.AaMode.Msaa(msaa) if msaa else None
    f2 = oracle.render(e.Teapot(u["m"], u["v"], u["p"], u["light_pos"], e.Sampler(shadow, e.abi.TEXEL_F32, e.abi.FILTER_LINEAR).clamped(),
                                u["light_vp"], u["cam_pos"], aa=aa), stream, color, depth, n_threads=0)["fragments"]
    return shadow, color, depth, f1, f2


def main():
    out = {}
    # C1
    sh, c, d, f1, f2 = teapot(640, 480, 512, 0)
    out.update(c1_shadow_crc=crc(sh), c1_depth_crc=crc(d), c1_color=c, c1_frags=np.array([f1, f2], np.uint64))
    # C2 (frame 250, tiled, uv x3)
    w, h = 1920, 1080
    verts, idx = scenes.cube_geometry(3.0)
    tex = scenes.rust_texture()
    color = np.full((h, w), 180, np.uint32)
    smp = e.Sampler(tex.view(np.uint32).reshape(tex.shape[0], tex.shape[1]), e.abi.TEXEL_RGBA8_TO_F32, e.abi.FILTER_LINEAR).tiled()
    f = oracle.render(e.Cube(scenes.cube_mvp(250, w, h), smp), e.IndexedVertices(idx, verts), color, None, n_threads=0)["fragments"]
    out.update(c2_color_crc=crc(color), c2_frags=np.uint64(f), c2_color_crop=color[400:656, 800:1056].copy())
    # C3
    sh, c, d, f1, f2 = teapot(3840, 2160, 2048, 1)
    out.update(c3_shadow_crc=crc(sh), c3_depth_crc=crc(d), c3_frags=np.array([f1, f2], np.uint64), c3_color_crop=c[700:1212, 1500:2012].copy(),
               c3_coverage_crc=crc(c != 0))
    # C4 full size
    w, h = 3840, 2160
    verts, idx = scenes.blend_tris(1 << 19, w, h)
    color = np.full((h, w), 0xFF000000, np.uint32)
    depth = np.full((h, w), 1.0, np.float32)
    f = oracle.render(e.BlendTris(), e.IndexedVertices(idx, verts), color, depth, n_threads=0)["fragments"]
    out.update(c4_color_crc=crc(color), c4_depth_crc=crc(depth), c4_frags=np.uint64(f), c4_color_crop=color[1000:1128, 1900:2028].copy())
    # C5: first 4 icons
    verts, idx, draws, ubs = scenes.voxel_icon_batch(4)
    cols, dcrc, fr = [], [], []
    for k in range(4):
        color = np.zeros((256, 256), np.uint32)
        depth = np.full((256, 256), 1.0, np.float32)
        first, count, base, _ = draws[k]
        st = oracle.render(e.VoxelIcon(scenes.voxel_icon_mvp(k), scenes.VOXEL_LIGHT_DIR), e.IndexedVertices(idx, verts), color, depth, draw=(first, count, base))
        cols.append(color); dcrc.append(crc(depth)); fr.append(st["fragments"])
    out.update(c5_color=np.stack(cols), c5_depth_crc=np.array(dcrc, np.uint32), c5_frags=np.array(fr, np.uint64))
    path = os.path.join(ROOT, "tests", "golden", "golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k, v in out.items():
        if np.ndim(v) <= 1:
            print(" ", k, v)


if __name__ == "__main__":
    main()
