"""Adds oracle-generated pins for 64 icons spread over the 4096-icon batch of BASELINE config 5 to tests/golden/golden.npz
(keys c5s_ids, c5s_color_crc, c5s_depth_crc, c5s_frags), leaving the other keys untouched.  bench.py and the GPU suite
verify the sharded batch against them.  Oracle-generated: PARITY UNPINNED (see oracle/euc_oracle.hpp).
    python tools/make_golden_c5.py"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import euc_b200 as e  # noqa: E402
from euc_b200 import scenes  # noqa: E402
from oracle import oracle  # noqa: E402

N_ICONS, N_SAMPLES = 4096, 64


def sample_ids():
    return (np.arange(N_SAMPLES, dtype=np.int64) * 64 + (np.arange(N_SAMPLES) * 37) % 64).astype(np.uint32)  # one per block of 64, varying offset


def main():
    path = os.path.join(ROOT, "tests", "golden", "golden.npz")
    out = dict(np.load(path))
    ids = sample_ids()
    ccrc, dcrc, fr = [], [], []
    for k in ids:
        verts, idx = scenes.voxel_icon_mesh(int(k))
        color = np.zeros((256, 256), np.uint32)
        depth = np.full((256, 256), 1.0, np.float32)
        st = oracle.render(e.VoxelIcon(scenes.voxel_icon_mvp(int(k)), scenes.VOXEL_LIGHT_DIR), e.IndexedVertices(idx, verts), color, depth)
        ccrc.append(zlib.crc32(color.tobytes())); dcrc.append(zlib.crc32(depth.tobytes())); fr.append(st["fragments"])
    out.update(c5s_ids=ids, c5s_color_crc=np.array(ccrc, np.uint32), c5s_depth_crc=np.array(dcrc, np.uint32), c5s_frags=np.array(fr, np.uint64))
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", N_SAMPLES, "icons, fragments", int(np.sum(fr)))


if __name__ == "__main__":
    main()
