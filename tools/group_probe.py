"""torchrun probe of the group path: times (a) band render without gather/barrier, (b) barrier alone, (c) render + root gather + barrier."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import euc_b200 as e
from euc_b200 import parallel, scenes
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = e.Context(lr)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
w, h = 3840, 2160
verts, idx = scenes.blend_tris(1 << 19, w, h)
geom = e.Geometry(verts, idx, ctx)
color = e.Buffer2d([w, h], np.uint32, ctx); depth = e.Buffer2d([w, h], np.float32, ctx)
grp = parallel.Group(ctx, f"probe_{os.environ['MASTER_PORT']}", rank, world)
peers = grp.share(color)
pipe = e.BlendTris().freeze()
r0, r1 = grp.rows(h)
def timed(fn, k=50):
    for _ in range(5): fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record(stream)
    for _ in range(k): fn()
    b.record(stream); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    return a.elapsed_time(b) / k, (t1 - t0) * 1e3 / k, (t2 - t0) * 1e3 / k
res = {}
res["band only"] = timed(lambda: pipe.render(geom, color, depth, rows=(r0, r1), clear=(0xFF000000, 1.0)))
res["barrier only"] = timed(lambda: grp.barrier())
res["band+barrier"] = timed(lambda: (pipe.render(geom, color, depth, rows=(r0, r1), clear=(0xFF000000, 1.0)), grp.barrier()))
res["group root"] = timed(lambda: grp.render(pipe, geom, peers, depth, gather=e.abi.GATHER_ROOT, clear=(0xFF000000, 1.0)))
res["group all"] = timed(lambda: grp.render(pipe, geom, peers, depth, gather=e.abi.GATHER_ALL, clear=(0xFF000000, 1.0)))
ctx.get_profile(reset=True); ctx.set_profiling(True)
res["root+profiling"] = timed(lambda: grp.render(pipe, geom, peers, depth, gather=e.abi.GATHER_ROOT, clear=(0xFF000000, 1.0)))
prof = ctx.get_profile(reset=True); ctx.set_profiling(False)
stages = {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]}
res["root again"] = timed(lambda: grp.render(pipe, geom, peers, depth, gather=e.abi.GATHER_ROOT, clear=(0xFF000000, 1.0)), k=200)
allres = [None] * world
dist.all_gather_object(allres, (rank, {k: round(v[0], 4) for k, v in res.items()}, stages))
if rank == 0:
    for k in res:
        print(f"{k:15s}", " ".join(f"{r[1][k]:.4f}" for r in allres), flush=True)
    for r in allres:
        print("stages rank", r[0], r[2], flush=True)
dist.barrier()
grp.close()
dist.destroy_process_group()
