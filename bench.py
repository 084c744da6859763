#!/usr/bin/env python
"""bench.py — frames/s of euc's `Pipeline::render` hot path on B200 (and of the CPU restatement beside it).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c1|c2|c3|c5] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...        (one rank per GPU; NCCL)

A "step" is one frame: clears + every pass of the workload.  Default workload = BASELINE config 4
("C4": 2^20-triangle indexed TriangleList, random depth + alpha blend, 3840x2160), the configuration the
north star's target is quoted on.  At N > 1 the C4 frame is split into screen-space row bands (one per rank, tile
aligned) and the colour rows are all-gathered over NCCL ("strong" scaling); C5 shards independent icon frames
across ranks with no collective ("weak").

value = whole-job frames/s with geometry and targets resident in HBM.
e2e   = frames/s through the host-facing call: per step the geometry is copied host->device from pinned memory and
        the finished colour buffer is read back device->host.
roofline = raster kernel: algorithmic bytes per frame / its mean launch duration (CUDA events around the
        launch, inside the timed region) against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
cpu_baseline = the C++ restatement of euc's render_par (oracle/, all host cores) on a bounded sample of the same frame.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--icons", type=int, default=4096, help="C5: icons in the whole job")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# workload definitions shared by both arms
# ---------------------------------------------------------------------------------------------------------
WORKLOADS = {
    "c1": dict(name="C1 teapot shadow(512^2)+phong 640x480", w=640, h=480, shadow=512, msaa=0),
    "c2": dict(name="C2 textured cube bilinear+tiled 1920x1080", w=1920, h=1080),
    "c3": dict(name="C3 teapot shadow(2048^2)+phong 3840x2160 msaa level 1", w=3840, h=2160, shadow=2048, msaa=1),
    "c4": dict(name="C4 2^20-triangle indexed TriangleList, random depth + alpha blend, 3840x2160", w=3840, h=2160, quads=1 << 19),
    "c5": dict(name="C5 voxel-icon batch 256x256, depth + blend", w=256, h=256),
}


def algorithmic_bytes(wl, args, scene):
    """compulsory I/O, every byte once: indices + vertices + sampled textures + written targets."""
    c = WORKLOADS[wl]
    w, h = c["w"], c["h"]
    if wl in ("c1", "c3"):
        s = c["shadow"]
        vb = scene["stream"].nbytes  # 6768 * 24
        return vb + s * s * 4 + vb + s * s * 4 + w * h * 8
    if wl == "c2":
        return scene["idx"].nbytes + scene["verts"].nbytes + scene["tex"].nbytes + w * h * 4
    if wl == "c4":
        return scene["idx"].nbytes + scene["verts"].nbytes + w * h * 8
    if wl == "c5":
        return scene["idx"].nbytes + scene["verts"].nbytes + scene["n_icons"] * w * h * 8
    raise ValueError(wl)


def build_scene(wl, args, rank=0, world=1):
    from euc_b200 import scenes
    c = WORKLOADS[wl]
    if wl in ("c1", "c3"):
        return dict(stream=scenes.teapot_stream(), u=scenes.teapot_uniforms(c["w"], c["h"], c["shadow"]))
    if wl == "c2":
        verts, idx = scenes.cube_geometry(uv_scale=3.0)
        return dict(verts=verts, idx=idx, tex=scenes.rust_texture(), mvp=scenes.cube_mvp(250, c["w"], c["h"]))
    if wl == "c4":
        verts, idx = scenes.blend_tris(c["quads"], c["w"], c["h"])
        return dict(verts=verts, idx=idx)
    if wl == "c5":
        per = args.icons // world
        verts, idx, draws, ubs = scenes.voxel_icon_batch(per, first_icon=rank * per)
        return dict(verts=verts, idx=idx, draws=draws, ubs=ubs, n_icons=per)


# ---------------------------------------------------------------------------------------------------------
# clocks sampler (NVML), during the timed regions
# ---------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.max_mhz, self.ok = [], set(), False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        while self.ok and not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the C++ restatement of euc (oracle/) on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_frame(wl, scene, rows=None, threads=0):
    """One frame of `wl` on the oracle (clears + all passes).  Returns (seconds, fragments).  rows=(r0, r1) restricts
    the final pass to the euc bands intersecting those rows (bounded sample)."""
    import euc_b200 as e
    from oracle import oracle
    c = WORKLOADS[wl]
    w, h = c["w"], c["h"]
    t0 = time.perf_counter()
    frags = 0
    if wl in ("c1", "c3"):
        s, u = c["shadow"], scene["u"]
        shadow = np.empty((s, s), np.float32); shadow.fill(1.0)
        color = np.zeros((h, w), np.uint32)
        depth = np.empty((h, w), np.float32); depth.fill(1.0)
        frags += oracle.render(e.TeapotShadow(u["shadow_mvp"]), scene["stream"], None, shadow, n_threads=threads)["fragments"]
        aa = e.AaMode.Msaa(c["msaa"]) if c["msaa"] else None
        frags += oracle.render(e.Teapot(u["m"], u["v"], u["p"], u["light_pos"], e.Sampler(shadow, e.abi.TEXEL_F32, e.abi.FILTER_LINEAR).clamped(),
                                        u["light_vp"], u["cam_pos"], aa=aa), scene["stream"], color, depth, n_threads=threads, rows=rows)["fragments"]
    elif wl == "c2":
        color = np.empty((h, w), np.uint32); color.fill(180)
        t = scene["tex"]
        smp = e.Sampler(t.view(np.uint32).reshape(t.shape[0], t.shape[1]), e.abi.TEXEL_RGBA8_TO_F32, e.abi.FILTER_LINEAR).tiled()
        frags += oracle.render(e.Cube(scene["mvp"], smp), e.IndexedVertices(scene["idx"], scene["verts"]), color, None, n_threads=threads, rows=rows)["fragments"]
    elif wl == "c4":
        color = np.empty((h, w), np.uint32); color.fill(0xFF000000)
        depth = np.empty((h, w), np.float32); depth.fill(1.0)
        frags += oracle.render(e.BlendTris(), e.IndexedVertices(scene["idx"], scene["verts"]), color, depth, n_threads=threads, rows=rows)["fragments"]
    elif wl == "c5":
        from concurrent.futures import ThreadPoolExecutor
        iv = e.IndexedVertices(scene["idx"], scene["verts"])
        n = scene["n_icons"] if rows is None else rows
        ubs = np.frombuffer(scene["ubs"], dtype=np.float32).reshape(-1, 20)

        def one(k, nthreads):
            color = np.zeros((h, w), np.uint32)
            depth = np.empty((h, w), np.float32); depth.fill(1.0)
            first, count, base, _ = scene["draws"][k]
            return oracle.render(e.VoxelIcon(ubs[k][:16].reshape(4, 4).T, ubs[k][16:19]), iv, color, depth, n_threads=nthreads, draw=(first, count, base))["fragments"]

        if threads == "frame-parallel":
            # one icon per host core, each rendered by a single thread walking euc's bands (ctypes releases the GIL)
            with ThreadPoolExecutor(max_workers=oracle.hardware_concurrency()) as ex:
                frags += sum(ex.map(lambda k: one(k, 1), range(n)))
        else:
            for k in range(n):
                frags += one(k, threads)
    return time.perf_counter() - t0, frags


def cpu_plan(wl, scene, budget_s=10.0):
    """Chooses a bounded sample of the frame for the CPU arm.  Returns (run, scale, description, cores): run() renders
    the sample once and returns seconds; one frame costs about run() * scale seconds."""
    from oracle import oracle
    cores = oracle.hardware_concurrency()
    c = WORKLOADS[wl]
    h = c["h"]
    if wl == "c5":
        n = min(scene["n_icons"], 256)
        # two ways a CPU user would run the batch: icon after icon with euc's own band threads (only 3 on a 256-row target),
        # or one icon per core; the faster one is the baseline
        t_seq = cpu_frame(wl, scene, rows=min(n, 32))[0] / min(n, 32)
        t_par = cpu_frame(wl, scene, rows=n, threads="frame-parallel")[0] / n
        if t_par <= t_seq:
            return (lambda: cpu_frame(wl, scene, rows=n, threads="frame-parallel")[0]), 1.0 / n, f"{n} icons of the batch, frame-parallel: one icon per host core, band structure unchanged", cores
        return (lambda: cpu_frame(wl, scene, rows=n)[0]), 1.0 / n, f"{n} icons of the batch, one after another, each by euc's band threads", cores
    if wl in ("c1", "c2"):
        return (lambda: cpu_frame(wl, scene)[0]), 1.0, "one full frame", cores
    # big frames: probe 1/8 of the rows (doubles as warm-up), take the whole frame if it fits the budget
    r1 = (h // 8) // 80 * 80
    t8, _ = cpu_frame(wl, scene, rows=(0, r1))
    if t8 * h / r1 <= budget_s:
        return (lambda: cpu_frame(wl, scene)[0]), 1.0, "one full frame", cores
    return ((lambda: cpu_frame(wl, scene, rows=(0, r1))[0]), h / r1,
            f"the euc bands covering rows [0,{r1}) of {h} (every band walks all primitives), scaled by {h}/{r1}", cores)


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = args.workload
    steps = args.steps if args.steps is not None else 3
    warm = args.warmup if args.warmup is not None else 1
    scene = build_scene(wl, args)
    run, scale, desc, cores = cpu_plan(wl, scene, budget_s=6.0)
    t_budget = time.perf_counter() + 150.0  # keep the whole arm within a few minutes
    for _ in range(warm):
        run()
        if time.perf_counter() > t_budget:
            break
    times = []
    for _ in range(steps):
        times.append(run())
        if time.perf_counter() > t_budget:
            break
    sec_per_frame = float(np.mean(times)) * scale
    value = 1.0 / sec_per_frame
    line = {
        "impl": "reference", "metric": "frames_per_s", "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": len(times),
        "warmup": warm, "ms_per_step": 1000.0 * float(np.mean(times)), "higher_is_better": True,
        "scaling": "weak" if wl == "c5" else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[wl]["name"], "target": f"{WORKLOADS[wl]['w']}x{WORKLOADS[wl]['h']}",
                   "note": "C++ restatement of euc's render_par (rustc unavailable here), all host cores; each step is the bounded sample"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import euc_b200 as e

    wl = args.workload
    c = WORKLOADS[wl]
    w, h = c["w"], c["h"]
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = e.Context(local_rank)
    stream = torch.cuda.Stream()  # a real (non-default) stream: shared by torch ops, NCCL waits and the raster library
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    scene = build_scene(wl, args, rank, world)
    steps = args.steps if args.steps is not None else {"c4": 100, "c3": 100, "c5": 10}.get(wl, 300)
    warm = max(3, args.warmup if args.warmup is not None else 5)

    h2d = d2h = 0
    frags_per_frame = None
    keep = []

    pipelined = None  # (list of per-slot e2e frame functions, list of torch streams) when e2e keeps 2 frames in flight
    if wl == "c4":
        # row partition: tile rows (16 px) split evenly; every rank owns one contiguous slot of the gather buffer
        from euc_b200 import parallel
        slot_rows, bands = parallel.row_band_slots(h, world)
        r0, r1 = bands[rank]
        pv = torch.from_numpy(scene["verts"].view(np.uint8)).pin_memory()
        pi = torch.from_numpy(scene["idx"].view(np.uint8)).pin_memory()
        pipe = e.BlendTris().freeze()
        keep += [pv, pi]

        fused = world > 1 and os.environ.get("EUC_GATHER", "p2p") != "nccl"

        def make_slot(cx, st):
            """One in-flight frame: its own context/stream, geometry, targets and host read-back buffer."""
            with torch.cuda.stream(st):
                host_out = torch.empty(h * w, dtype=torch.int32).pin_memory()
                token = torch.zeros(1, dtype=torch.int32, device="cuda")
            depth = e.Buffer2d([w, h], np.float32, cx)
            if world > 1:
                # geometry lives in torch tensors so that the e2e path can upload 1/world of it per rank (own PCIe link)
                # and all-gather the rest over NVLink
                with torch.cuda.stream(st):
                    dv = torch.from_numpy(scene["verts"].view(np.uint8)).cuda()
                    di = torch.from_numpy(scene["idx"].view(np.uint8)).cuda()
                geom = e.Geometry.wrap(dv.data_ptr(), scene["verts"].dtype.itemsize, scene["verts"].shape[0], di.data_ptr(), scene["idx"].size, cx)
                nv, ni = dv.numel() // world, di.numel() // world
                assert nv * world == dv.numel() and ni * world == di.numel() and nv % 16 == 0 and ni % 16 == 0
                keep.extend([dv, di])
            else:
                geom = e.Geometry(scene["verts"], scene["idx"], cx)
            if fused:
                # fused gather: every rank's raster kernel stores its colour rows into all peers' framebuffers (CUDA IPC
                # mappings, NVLink); a 4-byte all-reduce is the only collective (completion barrier)
                color = e.Buffer2d([w, h], np.uint32, cx)
                color.clear(0xFF000000)
                handles = [None] * world
                dist.all_gather_object(handles, color.ipc_export())
                mirrors = [e.Buffer2d.ipc_import(handles[r], [w, h], np.uint32, cx) for r in range(world) if r != rank]
                gather = color.as_torch()
                keep.extend(mirrors)
            else:
                gather = torch.empty(world * slot_rows * w, dtype=torch.int32, device="cuda")
                color = e.Buffer2d.wrap(gather.data_ptr(), [w, h], np.uint32, cx)
                mirrors = None
                my_slot = gather[rank * slot_rows * w:(rank + 1) * slot_rows * w]
            keep.extend([gather, host_out, color, depth, geom, token])

            def frame():
                # clears + render in one call (euc_render_clear: the tile kernels start from the clear values and write
                # every tile of the rendered rows; same bytes as clear(); clear(); render(), tests/test_fused_clear.py)
                if fused:
                    pipe.render(geom, color, depth, rows=(r0, r1), mirrors=mirrors, clear=(0xFF000000, 1.0))
                    dist.all_reduce(token)  # stream-ordered barrier: every rank's rows have landed everywhere
                else:
                    pipe.render(geom, color, depth, rows=(r0, r1), clear=(0xFF000000, 1.0))
                    if world > 1:
                        dist.all_gather_into_tensor(gather, my_slot)

            def frame_e2e():
                if world > 1:
                    # sharded upload: this rank's 1/world of the vertices and indices over its own PCIe link, then NVLink
                    dv[rank * nv:(rank + 1) * nv].copy_(pv[rank * nv:(rank + 1) * nv], non_blocking=True)
                    di[rank * ni:(rank + 1) * ni].copy_(pi[rank * ni:(rank + 1) * ni], non_blocking=True)
                    dist.all_gather_into_tensor(dv, dv[rank * nv:(rank + 1) * nv])
                    dist.all_gather_into_tensor(di, di[rank * ni:(rank + 1) * ni])
                    frame()
                    # sharded read-back: every rank returns its own rows of the frame to the host
                    host_out[r0 * w:r1 * w].copy_(gather[r0 * w:r1 * w], non_blocking=True)
                    if fused:
                        dist.all_reduce(token)  # peers must not write the next frame into rows that are still being read
                else:
                    geom.update(pv.data_ptr(), pi.data_ptr())
                    frame()
                    host_out.copy_(gather[: h * w], non_blocking=True)

            return frame, frame_e2e, gather

        frame, frame_e2e, gather = make_slot(ctx, stream)
        stream2 = torch.cuda.Stream()
        ctx2 = e.Context(local_rank)
        ctx2.set_stream(stream2.cuda_stream)
        _, frame_e2e_b, _ = make_slot(ctx2, stream2)
        keep += [ctx2, stream2]
        pipelined = ([frame_e2e, frame_e2e_b], [stream, stream2])
        n_slots = int(os.environ.get("EUC_E2E_SLOTS", "2")) if world == 1 else 2
        for _ in range(max(0, n_slots - 2)):
            # more frames in flight (measured: 2, 3 and 4 slots all give 1.634 ms per frame: the 80 MB upload at ~49 GB/s is
            # the bound, so the default stays at two)
            st_k = torch.cuda.Stream()
            cx_k = e.Context(local_rank)
            cx_k.set_stream(st_k.cuda_stream)
            _, fe_k, _ = make_slot(cx_k, st_k)
            keep += [cx_k, st_k]
            pipelined[0].append(fe_k)
            pipelined[1].append(st_k)

        def verify():
            """N-GPU (or 1-GPU) frame against the oracle-generated golden CRC at full size (bit-exact: this shader has no
            transcendental)."""
            import zlib
            frame()
            torch.cuda.synchronize()
            got = gather[: h * w].cpu().numpy().view(np.uint32)
            gold = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
            return bool(zlib.crc32(got.tobytes()) == int(gold["c4_color_crc"]))

        h2d, d2h = pv.numel() + pi.numel(), h * w * 4
        config_extra = {"partition": (f"{world} row bands of {slot_rows} rows; " + ("colour rows stored into every peer framebuffer by the raster kernel (CUDA IPC / NVLink), 4-byte all-reduce as barrier"
                                      if fused else "NCCL all_gather of colour rows")) if world > 1 else "single GPU",
                        "l2": "working set (80 MB geometry + 151 MB setup records + 66 MB targets) > 126 MB L2; no flush",
                        "e2e_pipeline": "2 frames in flight (one context / stream each): H2D of frame i+1 and D2H of frame i-1 overlap the kernels of frame i"
                                        + ("; every rank uploads 1/N of the geometry (NCCL all-gather over NVLink completes it) and reads back its own rows" if world > 1 else "")}
        if world > 1:
            h2d, d2h = (pv.numel() + pi.numel()) // world * world, h * w * 4  # whole-job bytes per step, spread over the ranks
    elif wl in ("c1", "c3"):
        s, u = c["shadow"], scene["u"]
        geom = e.Geometry(scene["stream"], None, ctx)
        shadow = e.Buffer2d([s, s], np.float32, ctx)
        color = e.Buffer2d([w, h], np.uint32, ctx)
        depth = e.Buffer2d([w, h], np.float32, ctx)
        aa = e.AaMode.Msaa(c["msaa"]) if c["msaa"] else None
        p1 = e.TeapotShadow(u["shadow_mvp"]).freeze()
        p2 = e.Teapot(u["m"], u["v"], u["p"], u["light_pos"], shadow.linear().clamped(), u["light_vp"], u["cam_pos"], aa=aa).freeze()
        empty = e.Empty()
        pv = torch.from_numpy(scene["stream"].view(np.uint8)).pin_memory()
        host_out = torch.empty(h * w, dtype=torch.int32).pin_memory()
        cptr, _ = color.device_ptr()
        keep += [pv, host_out]

        def frame(count=None):
            p1.render(geom, empty, shadow, clear=(None, 1.0))
            if count is not None:
                count.append(ctx.get_stats()["fragments"])
            p2.render(geom, color, depth, clear=(0, 1.0))

        def frame_e2e():
            geom.update(pv.data_ptr())
            frame()
            ctx._check(ctx._lib.euc_buf_download(ctx._p, color.handle, host_out.data_ptr(), h * w * 4))

        h2d, d2h = pv.numel(), h * w * 4
        config_extra = {"partition": "replicas" if world > 1 else "single GPU", "l2": "flush: 256 MB scratch write between timed frames" if wl == "c1" else "targets 100 MB + records; no flush"}
        if wl == "c3":
            # e2e with two frames in flight (the 33 MB read-back of frame i-1 overlaps the kernels of frame i); C1 keeps the
            # one-frame-at-a-time loop because its timed frames are separated by an L2 flush
            def make_teapot_slot(cx, st):
                with torch.cuda.stream(st):
                    out_s = torch.empty(h * w, dtype=torch.int32).pin_memory()
                g_s = e.Geometry(scene["stream"], None, cx)
                sh_s, c_s, d_s = e.Buffer2d([s, s], np.float32, cx), e.Buffer2d([w, h], np.uint32, cx), e.Buffer2d([w, h], np.float32, cx)
                q1 = e.TeapotShadow(u["shadow_mvp"]).freeze()
                q2 = e.Teapot(u["m"], u["v"], u["p"], u["light_pos"], sh_s.linear().clamped(), u["light_vp"], u["cam_pos"], aa=aa).freeze()
                dev = c_s.as_torch()
                keep.extend([out_s, g_s, sh_s, c_s, d_s, dev, q1, q2])

                def fe2e():
                    g_s.update(pv.data_ptr())
                    q1.render(g_s, empty, sh_s, clear=(None, 1.0))
                    q2.render(g_s, c_s, d_s, clear=(0, 1.0))
                    out_s.copy_(dev[: h * w], non_blocking=True)
                return fe2e

            stream2 = torch.cuda.Stream()
            ctx2 = e.Context(local_rank)
            ctx2.set_stream(stream2.cuda_stream)
            pipelined = ([make_teapot_slot(ctx, stream), make_teapot_slot(ctx2, stream2)], [stream, stream2])
            keep += [ctx2, stream2]
            config_extra["e2e_pipeline"] = "2 frames in flight (2 contexts / streams): D2H of frame i-1 overlaps the kernels of frame i"
    elif wl == "c2":
        geom = e.Geometry(scene["verts"], scene["idx"], ctx)
        tex = e.Buffer2d.from_array(scene["tex"], ctx)
        color = e.Buffer2d([w, h], np.uint32, ctx)
        pipe = e.Cube(scene["mvp"], tex.linear().tiled()).freeze()
        empty = e.Empty()
        host_out = torch.empty(h * w, dtype=torch.int32).pin_memory()
        pv = torch.from_numpy(scene["verts"].view(np.uint8)).pin_memory()
        pi = torch.from_numpy(scene["idx"].view(np.uint8)).pin_memory()
        keep += [pv, pi, host_out]

        def frame():
            pipe.render(geom, color, empty, clear=(180, None))

        def frame_e2e():
            geom.update(pv.data_ptr(), pi.data_ptr())
            frame()
            ctx._check(ctx._lib.euc_buf_download(ctx._p, color.handle, host_out.data_ptr(), h * w * 4))

        h2d, d2h = pv.numel() + pi.numel(), h * w * 4
        config_extra = {"partition": "replicas" if world > 1 else "single GPU", "l2": "flush: 256 MB scratch write between timed frames"}
    elif wl == "c5":
        n = scene["n_icons"]
        geom = e.Geometry(scene["verts"], scene["idx"], ctx)
        color = e.Buffer2d([w, h], np.uint32, ctx, layers=n)
        depth = e.Buffer2d([w, h], np.float32, ctx, layers=n)
        pipe = e.VoxelIcon(np.eye(4), e.scenes.VOXEL_LIGHT_DIR)
        pv = torch.from_numpy(scene["verts"].view(np.uint8)).pin_memory()
        pi = torch.from_numpy(scene["idx"].view(np.uint8)).pin_memory()
        host_out = torch.empty(n * h * w, dtype=torch.int32).pin_memory()
        keep += [pv, pi, host_out]

        def frame():
            pipe.render_batch(geom, scene["draws"], scene["ubs"], color, depth, clear=(0, 1.0))

        def make_icon_slot(cx, st, geom_s, color_s, depth_s):
            """One in-flight batch of the e2e loop: upload of the geometry, render, asynchronous read-back of all icons."""
            with torch.cuda.stream(st):
                out_s = torch.empty(n * h * w, dtype=torch.int32).pin_memory()
            dev = color_s.as_torch()
            keep.extend([out_s, dev, geom_s, color_s, depth_s])

            def fe2e():
                geom_s.update(pv.data_ptr(), pi.data_ptr())
                pipe.render_batch(geom_s, scene["draws"], scene["ubs"], color_s, depth_s, clear=(0, 1.0))
                out_s.copy_(dev[: n * h * w], non_blocking=True)
            return fe2e

        frame_e2e = make_icon_slot(ctx, stream, geom, color, depth)
        stream2 = torch.cuda.Stream()
        ctx2 = e.Context(local_rank)
        ctx2.set_stream(stream2.cuda_stream)
        frame_e2e_b = make_icon_slot(ctx2, stream2, e.Geometry(scene["verts"], scene["idx"], ctx2), e.Buffer2d([w, h], np.uint32, ctx2, layers=n),
                                     e.Buffer2d([w, h], np.float32, ctx2, layers=n))
        keep += [ctx2, stream2]
        pipelined = ([frame_e2e, frame_e2e_b], [stream, stream2])

        h2d, d2h = pv.numel() + pi.numel(), n * h * w * 4
        config_extra = {"partition": f"{n} icons per rank x {world} rank(s), no collective", "icons_per_step": n * world,
                        "l2": f"targets {n * w * h * 8 / 1e6:.0f} MB > 126 MB L2; no flush",
                        "e2e_pipeline": "2 batches in flight (2 contexts / streams): H2D of batch i+1 and D2H of batch i-1 overlap the kernels of batch i"}

    flush_buf = torch.empty(64 * 1024 * 1024, dtype=torch.int32, device="cuda") if wl in ("c1", "c2") else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k, flush):
        """k steps on the device clock: events bracket each step (so an L2 flush between steps is excluded)."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        barrier()
        if flush is None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(k):
                fn()
            b.record(stream)
            barrier()
            ms = a.elapsed_time(b)
        else:
            for a, b in evs:
                flush.fill_(1)
                a.record(stream)
                fn()
                b.record(stream)
            barrier()
            ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # fragment count of one frame (the reference's emit_fragment count), outside the timed region
    ctx.set_stats(True)
    first_pass = []
    if wl in ("c1", "c3"):
        frame(first_pass)  # two passes: stats are per render call
    else:
        frame()
    st = ctx.get_stats()
    ctx.set_stats(False)
    frags_per_frame = int(st["fragments"]) + sum(first_pass)
    tf = torch.tensor([frags_per_frame], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tf)  # bands / icon shards add up
    frags_per_frame = int(tf.item())

    for _ in range(warm):
        frame()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.get_profile(reset=True)
    ctx.set_profiling(True)
    l0 = ctx.launch_count()
    ms = timed(frame, steps, flush_buf)
    launches = ctx.launch_count() - l0
    prof = ctx.get_profile(reset=True)
    ctx.set_profiling(False)
    # e2e
    e2e_steps = max(4, min(steps, 50))
    if pipelined is None:
        for _ in range(2):
            frame_e2e()
        ms_e2e = timed(frame_e2e, e2e_steps, flush_buf)
    else:
        fns, sts = pipelined

        def run_pipelined(k):
            for i in range(k):
                with torch.cuda.stream(sts[i % len(sts)]):
                    fns[i % len(fns)]()

        run_pipelined(2 * len(fns))
        barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(sts[0])
        run_pipelined(e2e_steps)
        for st_o in sts[1:]:
            sts[0].wait_stream(st_o)
        eb.record(sts[0])
        barrier()
        tt = torch.tensor([ea.elapsed_time(eb)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_e2e = float(tt.item())
    clocks = sampler.result()

    verified = verify() if wl == "c4" else None  # collective inside: every rank calls it
    frames_per_step = 1 if wl != "c5" else scene["n_icons"] * world
    job_mult = 1 if wl in ("c4", "c5") else world  # replicas render independent frames
    value = frames_per_step * job_mult * steps / (ms / 1000.0)
    e2e_value = frames_per_step * job_mult * e2e_steps / (ms_e2e / 1000.0)

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"
        raster_ms, raster_calls = prof["raster"]
        alg = algorithmic_bytes(wl, args, scene)
        if wl in ("c1", "c3"):
            per_launch_bytes = alg / 2.0  # two raster launches (shadow, phong) share the frame's bytes
        else:
            per_launch_bytes = alg / (world if wl == "c4" else 1)
        avg_raster_s = (raster_ms / max(raster_calls, 1)) / 1000.0
        achieved = per_launch_bytes / avg_raster_s / 1e9 if avg_raster_s > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(wl)
        stage_ms = {k: (v[0] / max(v[1], 1)) for k, v in prof.items()}
        line = {
            "metric": "frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak" if wl == "c5" else ("strong" if wl == "c4" else "weak"),
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict({"workload": c["name"], "target": f"{w}x{h}", "frames_per_step": frames_per_step,
                            "clears": "every frame clears its targets; the clears are fused into the render calls (euc_render_clear)"}, **config_extra),
            "mfrag_per_s": frags_per_frame * job_mult * steps / (ms / 1000.0) / 1e6,
            "fragments_per_step": frags_per_frame,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "raster_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": per_launch_bytes, "avg_launch_ms": avg_raster_s * 1000.0,
                         "peak_source": peak_src, "frame_frac": (alg / ((ms / steps) / 1000.0) / 1e9) / peak if wl != "c5" else None},
            "stage_ms_per_launch": stage_ms,
        }
        if verified is not None:
            line["frame_matches_golden_crc"] = verified
        if world == 1 and not args.no_cpu_baseline:
            try:
                run, scale, desc, cores = cpu_plan(wl, scene)
                run()
                t_cpu = min(run() for _ in range(2)) * scale
                line["cpu_baseline"] = {"value": 1.0 / t_cpu, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc + "; best of 2 after 1 warm-up"}
            except Exception as ex:  # the oracle is a reported baseline, never a dependency of the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": None, "kind": "port", "sample": f"failed: {ex}"}
        print(json.dumps(line), flush=True)
    barrier()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
