"""One-GPU emulation of what one rank of an N-GPU C4 frame does: render only the rows of rank 0 of N and time the stages
for the development knob EUC_SPARSE_RECS.  usage: python tools/band_probe.py [N ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import euc_b200 as e
from euc_b200 import scenes, parallel

w, h = 3840, 2160
verts, idx = scenes.blend_tris(1 << 19, w, h)
for world in [int(a) for a in sys.argv[1:]] or [8, 4]:
    _, bands = parallel.row_band_slots(h, world)
    r0, r1 = bands[0]
    for grid in (1,):
        for sparse in (0, 1):
            os.environ["EUC_SPARSE_RECS"] = str(sparse)
            ctx = e.Context(0)
            geom = e.Geometry(verts, idx, ctx)
            color, depth = e.Buffer2d([w, h], np.uint32, ctx), e.Buffer2d([w, h], np.float32, ctx)
            pipe = e.BlendTris().freeze()

            def frame():
                color.clear_rows(0xFF000000, r0, r1)
                depth.clear_rows(1.0, r0, r1)
                pipe.render(geom, color, depth, rows=(r0, r1))
            for _ in range(5):
                frame()
            ctx.sync()
            ctx.set_profiling(True); ctx.get_profile()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record()
            for _ in range(50):
                frame()
            ctx.sync(); b.record(); torch.cuda.synchronize()
            prof = ctx.get_profile()
            ctx.set_profiling(False)
            crc = int(np.bitwise_xor.reduce(color.raw()[r0:r1].reshape(-1)))
            print(f"N={world} rows={r0}:{r1} sparse_recs={sparse} ms/frame={a.elapsed_time(b) / 50:.4f} "
                  f"stages={ {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]} } xor={crc:08x}", flush=True)
            del geom, color, depth, ctx
