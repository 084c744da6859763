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
    ap.add_argument("--no-icon-batch", action="store_true", help="C4 runs only: skip the 4096-icon batch (BASELINE config 5) that rides along")
    ap.add_argument("--icon-steps", type=int, default=None, help="batches timed for the icon batch that rides along (default: 20 per GPU, i.e. a timed region of about 65 ms)")
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
        if world == 1:
            b, e_ = 0, args.icons
        else:
            from euc_b200 import parallel
            b, e_ = parallel.frame_shards(args.icons, world)[rank]  # contiguous icon ranges, no collective (euc_group_frames)
        verts, idx, draws, ubs = scenes.voxel_icon_batch(e_ - b, first_icon=b)
        return dict(verts=verts, idx=idx, draws=draws, ubs=ubs, n_icons=e_ - b, first_icon=b)


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
        # NVML queries take a driver lock that kernel launches of every process on the box contend for (measured: a 20 ms
        # sampler on each of 8 ranks cost 35 % of an N = 8 frame rate, one 40 ms sampler 12 %): one sampler per job, one clock
        # query per 50 ms, throttle reasons every fourth sample
        k = 0
        while self.ok and not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                if k % 4 == 0:
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in self.REASONS.items():
                        if r & bit and name != "gpu_idle":
                            self.reasons.add(name)
            except Exception:
                pass
            k += 1
            time.sleep(0.05)

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
        "config": {"workload": WORKLOADS[wl]["name"], "target": f"{WORKLOADS[wl]['w']}x{WORKLOADS[wl]['h']}", "frames_per_step": 1 if wl != "c5" else args.icons,
                   "clears": "every frame clears its targets (buffer fills, as the reference does)", "partition": "host cores (euc's own band threads)",
                   "note": "C++ restatement of euc's render_par (rustc unavailable here), all host cores; each step is the bounded sample"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
_PINNED = []  # (context, address) of every pinned allocation of the current measurement


def pinned_array(ctx, arr):
    """A copy of `arr` in pinned host memory obtained through the C ABI (euc_host_alloc); returns (numpy view, address)."""
    import ctypes as C
    a = np.ascontiguousarray(arr)
    ptr = ctx.host_alloc(max(a.nbytes, 1))
    view = np.ctypeslib.as_array((C.c_uint8 * max(a.nbytes, 1)).from_address(ptr))[: a.nbytes].view(a.dtype).reshape(a.shape)
    view[...] = a
    _PINNED.append((ctx, ptr))
    return view, ptr


def pinned_empty(ctx, nbytes):
    import ctypes as C
    ptr = ctx.host_alloc(max(nbytes, 1))
    _PINNED.append((ctx, ptr))
    return np.ctypeslib.as_array((C.c_uint8 * max(nbytes, 1)).from_address(ptr)), ptr


def crc32(a):
    import zlib
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def lsb_diff(a, b):
    return int(np.abs(np.ascontiguousarray(a).view(np.uint8).astype(np.int16) - np.ascontiguousarray(b).view(np.uint8).astype(np.int16)).max())


def measure(wl, args, rank, world, local_rank, steps=None, warm=None, cpu_baseline=True):
    """Times workload `wl` on this job's GPUs; returns the JSON line (dict) on rank 0, None elsewhere.  torch supplies the
    stream, the events and (N > 1) the process-level barrier / max-reduction of the timings; everything the frame does goes
    through the C ABI (euc_b200.core / euc_b200.parallel are thin ctypes wrappers)."""
    import torch
    import torch.distributed as dist
    import euc_b200 as e
    from euc_b200 import parallel

    c = WORKLOADS[wl]
    w, h = c["w"], c["h"]
    ctx = e.Context(local_rank)
    stream = torch.cuda.Stream()  # a real (non-default) stream: shared by the timing events and the raster library
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    scene = build_scene(wl, args, rank, world)
    steps = steps if steps is not None else (args.steps if args.steps is not None else {"c4": 300, "c3": 200, "c5": 10}.get(wl, 300))
    warm = max(3, warm if warm is not None else (args.warmup if args.warmup is not None else 5))
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
    job = f"{os.environ.get('MASTER_PORT', '0')}_{os.environ.get('TORCHELASTIC_RUN_ID', 'solo')}_{os.getppid() if world > 1 else os.getpid()}_{wl}"

    h2d = d2h = 0
    keep = []
    slots = []      # e2e: (frame function, context) per in-flight slot
    closers = []
    gather_mode = {"root": e.abi.GATHER_ROOT, "all": e.abi.GATHER_ALL}[os.environ.get("EUC_GATHER", "root")]

    def second_context():
        st2 = torch.cuda.Stream()
        cx2 = e.Context(local_rank)
        cx2.set_stream(st2.cuda_stream)
        keep.extend([st2, cx2])
        return cx2, st2

    if wl == "c4":
        r0, r1 = parallel.row_band(h, rank, world)
        pipe = e.BlendTris().freeze()
        nv, ni = scene["verts"].shape[0], scene["idx"].size
        v_sh, i_sh = parallel.frame_shards(nv, world)[rank], parallel.frame_shards(ni, world)[rank]

        def make_slot(cx, k):
            """One in-flight frame: its own context / stream, geometry, targets, group and pinned host buffers."""
            pv, pv_ptr = pinned_array(cx, scene["verts"])
            pi, pi_ptr = pinned_array(cx, scene["idx"])
            out, out_ptr = pinned_empty(cx, h * w * 4)
            color = e.Buffer2d([w, h], np.uint32, cx)
            depth = e.Buffer2d([w, h], np.float32, cx)
            color.clear(0xFF000000)
            keep.extend([pv, pi, out, color, depth])
            state = {"ticket": None}
            if world > 1:
                grp = parallel.Group(cx, f"{job}_{k}", rank, world)
                peers = grp.share(color)
                geom = e.Geometry(scene["verts"], scene["idx"], cx)
                closers.append(grp.close)
                keep.extend([grp, peers, geom])

                def frame():
                    # this rank's row band; the raster kernel stores the band's colour rows into the root's framebuffer too
                    # (NVLink peer stores); device-side flag barrier; nothing waits on the host
                    grp.render(pipe, geom, peers, depth, gather=gather_mode, clear=(0xFF000000, 1.0))

                def frame_e2e():
                    if state["ticket"] is not None:
                        cx.ticket_wait(state["ticket"])
                    # sharded upload (this rank's 1/N of the vertices and indices over its own PCIe link), exchange over NVLink
                    geom.update_range(pv_ptr + v_sh[0] * scene["verts"].dtype.itemsize, v_sh[0], v_sh[1] - v_sh[0], pi_ptr + i_sh[0] * 4, i_sh[0], i_sh[1] - i_sh[0])
                    grp.allgather_geom(geom)
                    frame()
                    state["ticket"] = color.download_async(out_ptr + r0 * w * 4, rows=(r0, r1))  # every rank returns its own rows
            else:
                geom = e.Geometry(scene["verts"], scene["idx"], cx)
                host_geom = e.IndexedVertices(pi, pv)
                keep.extend([geom])

                def frame():
                    pipe.render(geom, color, depth, clear=(0xFF000000, 1.0))

                def frame_e2e():
                    if state["ticket"] is not None:
                        cx.ticket_wait(state["ticket"])
                    pipe.render(host_geom, color, depth, clear=(0xFF000000, 1.0))  # euc_render: host pointers in, H2D inside the call
                    state["ticket"] = color.download_async(out_ptr)

            def drain():
                if state["ticket"] is not None:
                    cx.ticket_wait(state["ticket"])
                    state["ticket"] = None
            return frame, frame_e2e, drain, color, out

        frame, fe_a, drain_a, color, host_out = make_slot(ctx, 0)
        ctx2, stream2 = second_context()
        _, fe_b, drain_b, _, _ = make_slot(ctx2, 1)
        slots = [(fe_a, stream, drain_a), (fe_b, stream2, drain_b)]

        def verify():
            """The (N-GPU) frame against the oracle-generated golden CRC at full size (bit-exact: no transcendental in this
            shader).  Under the root gather rank 0 holds the whole frame."""
            frame()
            torch.cuda.synchronize()
            barrier()
            ok = True
            if rank == 0:
                ok = crc32(color.raw()) == int(gold["c4_color_crc"])
            return {"frame_matches_golden_crc": bool(ok)}

        h2d, d2h = scene["verts"].nbytes + scene["idx"].nbytes, h * w * 4
        config_extra = {"partition": (f"{world} row bands (tile aligned); the raster kernel stores each band's colour rows into "
                                      + ("the root's framebuffer" if gather_mode == e.abi.GATHER_ROOT else "every peer's framebuffer")
                                      + " (CUDA IPC / NVLink), device-side flag barrier (euc_group_render)") if world > 1 else "single GPU",
                        "l2": "working set (80 MB geometry + 151 MB setup records + 66 MB targets) > 126 MB L2; no flush",
                        "e2e_pipeline": "2 frames in flight (one context / stream each): H2D of frame i+1 and D2H of frame i-1 overlap the kernels of frame i"
                                        + ("; every rank uploads 1/N of the geometry (euc_group_allgather_geom completes it over NVLink) and reads back its own rows" if world > 1
                                           else "; euc_render with pinned host pointers, euc_buf_download_async")}
    elif wl in ("c1", "c3"):
        s, u = c["shadow"], scene["u"]
        aa = e.AaMode.Msaa(c["msaa"]) if c["msaa"] else None
        empty = e.Empty()

        def make_slot(cx):
            pv, pv_ptr = pinned_array(cx, scene["stream"])
            out, out_ptr = pinned_empty(cx, h * w * 4)
            geom = e.Geometry(scene["stream"], None, cx)
            shadow = e.Buffer2d([s, s], np.float32, cx)
            color = e.Buffer2d([w, h], np.uint32, cx)
            depth = e.Buffer2d([w, h], np.float32, cx)
            p1 = e.TeapotShadow(u["shadow_mvp"]).freeze()
            p2 = e.Teapot(u["m"], u["v"], u["p"], u["light_pos"], shadow.linear().clamped(), u["light_vp"], u["cam_pos"], aa=aa).freeze()
            keep.extend([pv, out, geom, shadow, color, depth, p1, p2])
            state = {"ticket": None}

            def frame(count=None):
                p1.render(geom, empty, shadow, clear=(None, 1.0))
                if count is not None:
                    count.append(cx.get_stats()["fragments"])
                p2.render(geom, color, depth, clear=(0, 1.0))

            def frame_e2e():
                if state["ticket"] is not None:
                    cx.ticket_wait(state["ticket"])
                p1.render(pv, empty, shadow, clear=(None, 1.0))   # euc_render: the vertex stream comes from pinned host memory
                p2.render(pv, color, depth, clear=(0, 1.0))
                state["ticket"] = color.download_async(out_ptr)

            def drain():
                if state["ticket"] is not None:
                    cx.ticket_wait(state["ticket"])
                    state["ticket"] = None
            return frame, frame_e2e, drain, (shadow, color, depth)

        frame, fe_a, drain_a, (shadow, color, depth) = make_slot(ctx)
        slots = [(fe_a, stream, drain_a)]
        if wl == "c3":
            ctx2, stream2 = second_context()
            _, fe_b, drain_b, _ = make_slot(ctx2)
            slots.append((fe_b, stream2, drain_b))

        def verify():
            frame()
            torch.cuda.synchronize()
            gs, gc, gd = shadow.raw(), color.raw(), depth.raw()
            res = {"shadow_crc_ok": crc32(gs) == int(gold[f"{wl}_shadow_crc"]), "depth_crc_ok": crc32(gd) == int(gold[f"{wl}_depth_crc"])}
            if wl == "c1":
                res["colour_max_lsb"] = lsb_diff(gc, gold["c1_color"])
            else:
                res["coverage_crc_ok"] = crc32(gc != 0) == int(gold["c3_coverage_crc"])
                res["colour_max_lsb"] = lsb_diff(gc[700:1212, 1500:2012], gold["c3_color_crop"])
            res["frame_matches_golden_crc"] = bool(all(v for k, v in res.items() if k.endswith("_ok")) and res["colour_max_lsb"] <= 1)
            return res

        h2d, d2h = 2 * scene["stream"].nbytes, h * w * 4
        config_extra = {"partition": "replicas" if world > 1 else "single GPU", "l2": "flush: 256 MB scratch write between timed frames" if wl == "c1" else "targets 100 MB + records; no flush",
                        "e2e_pipeline": ("2 frames in flight (2 contexts / streams); " if wl == "c3" else "") + "euc_render with the vertex stream in pinned host memory (both passes), euc_buf_download_async"}
    elif wl == "c2":
        geom = e.Geometry(scene["verts"], scene["idx"], ctx)
        tex = e.Buffer2d.from_array(scene["tex"], ctx)
        color = e.Buffer2d([w, h], np.uint32, ctx)
        pipe = e.Cube(scene["mvp"], tex.linear().tiled()).freeze()
        empty = e.Empty()
        pv, _ = pinned_array(ctx, scene["verts"])
        pi, _ = pinned_array(ctx, scene["idx"])
        out, out_ptr = pinned_empty(ctx, h * w * 4)
        host_geom = e.IndexedVertices(pi, pv)
        keep += [pv, pi, out]
        state = {"ticket": None}

        def frame():
            pipe.render(geom, color, empty, clear=(180, None))

        def fe_a():
            if state["ticket"] is not None:
                ctx.ticket_wait(state["ticket"])
            pipe.render(host_geom, color, empty, clear=(180, None))
            state["ticket"] = color.download_async(out_ptr)

        def drain_a():
            if state["ticket"] is not None:
                ctx.ticket_wait(state["ticket"])
                state["ticket"] = None
        slots = [(fe_a, stream, drain_a)]

        def verify():
            frame()
            torch.cuda.synchronize()
            return {"frame_matches_golden_crc": crc32(color.raw()) == int(gold["c2_color_crc"])}

        h2d, d2h = scene["verts"].nbytes + scene["idx"].nbytes, h * w * 4
        config_extra = {"partition": "replicas" if world > 1 else "single GPU", "l2": "flush: 256 MB scratch write between timed frames"}
    elif wl == "c5":
        n = scene["n_icons"]
        first_icon = scene["first_icon"]
        pipe = e.VoxelIcon(np.eye(4), e.scenes.VOXEL_LIGHT_DIR)

        def make_slot(cx):
            pv, pv_ptr = pinned_array(cx, scene["verts"])
            pi, pi_ptr = pinned_array(cx, scene["idx"])
            out, out_ptr = pinned_empty(cx, n * h * w * 4)
            geom = e.Geometry(scene["verts"], scene["idx"], cx)
            color = e.Buffer2d([w, h], np.uint32, cx, layers=n)
            depth = e.Buffer2d([w, h], np.float32, cx, layers=n)
            keep.extend([pv, pi, out, geom, color, depth])
            state = {"ticket": None}

            def frame():
                pipe.render_batch(geom, scene["draws"], scene["ubs"], color, depth, clear=(0, 1.0))

            def frame_e2e():
                if state["ticket"] is not None:
                    cx.ticket_wait(state["ticket"])
                geom.update(pv_ptr, pi_ptr)
                frame()
                state["ticket"] = color.download_async(out_ptr)

            def drain():
                if state["ticket"] is not None:
                    cx.ticket_wait(state["ticket"])
                    state["ticket"] = None
            return frame, frame_e2e, drain, (color, depth)

        frame, fe_a, drain_a, (color, depth) = make_slot(ctx)
        ctx2, stream2 = second_context()
        _, fe_b, drain_b, _ = make_slot(ctx2)
        slots = [(fe_a, stream, drain_a), (fe_b, stream2, drain_b)]

        def verify():
            """The sampled icons of tests/golden that fall into this rank's shard: depth CRC (bit-exact) and colour CRC."""
            frame()
            torch.cuda.synchronize()
            ids = gold["c5s_ids"].astype(np.int64)
            mine = [(k, int(i) - first_icon) for k, i in enumerate(ids) if first_icon <= i < first_icon + n]
            checked = bad_depth = bad_colour = 0
            if mine:
                cptr, _ = color.device_ptr()
                ct, dt = color.as_torch().view(n, h * w), depth.as_torch().view(n, h * w)
                for k, loc in mine:
                    cimg = ct[loc].cpu().numpy().view(np.uint32)
                    dimg = dt[loc].cpu().numpy()
                    checked += 1
                    bad_depth += int(crc32(dimg) != int(gold["c5s_depth_crc"][k]))
                    bad_colour += int(crc32(cimg) != int(gold["c5s_color_crc"][k]))
            t = torch.tensor([checked, bad_depth, bad_colour], dtype=torch.int64, device="cuda")
            if world > 1:
                dist.all_reduce(t)
            checked, bad_depth, bad_colour = (int(x) for x in t.tolist())
            return {"icons_checked": checked, "icons_depth_crc_mismatch": bad_depth, "icons_colour_crc_mismatch": bad_colour,
                    "frame_matches_golden_crc": bool(checked > 0 and bad_depth == 0 and bad_colour == 0)}

        h2d, d2h = (scene["verts"].nbytes + scene["idx"].nbytes), n * h * w * 4
        tot = torch.tensor([h2d, d2h], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(tot)
        h2d, d2h = (int(x) for x in tot.tolist())  # whole-job bytes per step
        config_extra = {"partition": f"{args.icons} icons in contiguous shards over {world} rank(s) ({n} on rank 0), no collective (euc_group_frames)", "icons_per_step": args.icons,
                        "l2": f"targets {n * w * h * 8 / 1e6:.0f} MB per rank > 126 MB L2; no flush",
                        "e2e_pipeline": "2 batches in flight (2 contexts / streams): H2D of batch i+1 and D2H of batch i-1 overlap the kernels of batch i"}

    flush_buf = torch.empty(64 * 1024 * 1024, dtype=torch.int32, device="cuda") if wl in ("c1", "c2") else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms):
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    trace = os.environ.get("EUC_BENCH_TRACE")  # diagnostics: host time of every call of a timed loop, collector runs

    def timed(fn, k, flush):
        """k steps on the device clock: events bracket each step (so an L2 flush between steps is excluded).  The cyclic
        garbage collector is kept out of the loop: a collection that finds device objects of an earlier measurement frees
        them (cudaFree / cudaFreeHost wait for the device) in the middle of the timed region."""
        import gc
        gc.collect()
        barrier()
        gc.disable()
        try:
            return timed_inner(fn, k, flush)
        finally:
            gc.enable()

    def timed_inner(fn, k, flush):
        if flush is None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            host = []
            a.record(stream)
            for _ in range(k):
                t0 = time.perf_counter()
                fn()
                host.append(time.perf_counter() - t0)
            b.record(stream)
            barrier()
            ms = a.elapsed_time(b)
            if trace and rank == 0:
                print(f"[trace] {wl} k={k} device {ms:.3f} ms, host per call (ms): " + " ".join(f"{1e3 * t:.2f}" for t in host[:64]), file=sys.stderr, flush=True)
        else:
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
            for a, b in evs:
                flush.fill_(1)
                a.record(stream)
                fn()
                b.record(stream)
            barrier()
            ms = sum(a.elapsed_time(b) for a, b in evs)
        return reduce_max(ms)

    # fragment count of one frame (the reference's emit_fragment count), outside the timed region
    ctx.set_stats(True)
    first_pass = []
    if wl in ("c1", "c3"):
        frame(first_pass)  # two passes: stats are per render call
    else:
        frame()
    st = ctx.get_stats()
    ctx.set_stats(False)
    tf = torch.tensor([int(st["fragments"]) + sum(first_pass)], dtype=torch.float64, device="cuda")
    if world > 1 and wl in ("c4", "c5"):
        dist.all_reduce(tf)  # bands / icon shards add up
    frags_per_frame = int(tf.item())

    for _ in range(warm):
        frame()
    barrier()
    waits0 = ctx.blocking_waits()
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("EUC_BENCH_NOSAMPLER"):  # one sampler per job: NVML queries from 8 processes disturb each other's launches
        sampler.start()
    # per-kernel durations: CUDA events around every launch, inside the timed region at N = 1 (the roofline is quoted there);
    # at N > 1 the event pairs between the kernels cost a few per cent of a 0.15 ms frame, so they are taken from a
    # separate short run and the timed region stays clean
    prof_in_region = world == 1 and not os.environ.get("EUC_BENCH_NOPROF")
    ctx.get_profile(reset=True)
    ctx.set_profiling(prof_in_region)
    l0 = ctx.launch_count()
    ms = timed(frame, steps, flush_buf)
    launches = ctx.launch_count() - l0
    host_waits = ctx.blocking_waits() - waits0
    prof_steps = steps
    if not prof_in_region:
        ctx.set_profiling(True)
        prof_steps = 20
        timed(frame, prof_steps, flush_buf)
    prof = ctx.get_profile(reset=True)
    ctx.set_profiling(False)

    # CUDA-graph replay of the same frame (single-draw renders are pure kernel launches: capturable)
    graph_ms = None
    if wl in ("c1", "c2", "c3") or (wl == "c4" and world == 1):
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                frame()
            torch.cuda.set_stream(stream)
            for _ in range(3):
                g.replay()

            def replay():
                g.replay()
            graph_ms = timed(replay, steps, flush_buf) / steps
            keep.append(g)
        except Exception as ex:  # capture is an extra; the frame loop above is the measurement
            graph_ms = f"unavailable: {str(ex)[:120]}"
            torch.cuda.set_stream(stream)

    # e2e: the same frames through the host-facing calls, inputs from pinned host memory, result read back
    e2e_steps = max(4, min(steps, 50))

    def run_slots(k):
        for i in range(k):
            fn, st_i, _ = slots[i % len(slots)]
            with torch.cuda.stream(st_i):
                fn()

    run_slots(2 * len(slots))
    for _, _, dr in slots:
        dr()
    barrier()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if flush_buf is None:
        import gc
        gc.collect()
        gc.disable()  # as in timed(): no collector run inside the timed region
        try:
            ea.record(stream)
            run_slots(e2e_steps)
            for _, st_o, _ in slots[1:]:
                stream.wait_stream(st_o)
            eb.record(stream)
            for _, _, dr in slots:
                dr()
        finally:
            gc.enable()
        barrier()
        ms_e2e = reduce_max(ea.elapsed_time(eb))
    else:
        fn0, _, dr0 = slots[0]

        def one():
            fn0()
        ms_e2e = timed(one, e2e_steps, flush_buf)
        dr0()
    clocks = sampler.result()

    verified = verify()  # collective inside (N > 1): every rank calls it
    frames_per_step = 1 if wl != "c5" else args.icons
    job_mult = 1 if wl in ("c4", "c5") else world  # replicas render independent frames
    value = frames_per_step * job_mult * steps / (ms / 1000.0)
    e2e_value = frames_per_step * job_mult * e2e_steps / (ms_e2e / 1000.0)

    line = None
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"
        alg = algorithmic_bytes(wl, args, scene)
        stage_ms = {k: (v[0] / max(v[1], 1)) for k, v in prof.items()}
        # the kernel that dominates the frame (raster for C1/C2/C4/C5, resolve for C3): its launches share the frame's bytes
        dom = max(prof.items(), key=lambda kv: kv[1][0])[0]
        dom_ms, dom_calls = prof[dom]
        dom_per_frame = max(1, round(dom_calls / prof_steps))
        per_launch_bytes = alg / dom_per_frame / (world if wl == "c4" else 1)
        avg_s = (dom_ms / max(dom_calls, 1)) / 1000.0
        achieved = per_launch_bytes / avg_s / 1e9 if avg_s > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath) and world == 1:  # ncu captures are single-GPU, whole-frame launches
            tj = json.load(open(tpath))
            traffic = tj.get(f"c5_{args.icons}_{dom}") if wl == "c5" else tj.get(f"{wl}_{dom}")
        kernel_names = {"raster": "raster_kernel", "resolve": "resolve_kernel", "setup": "setup_kernel", "alloc": "alloc_tiles_kernel", "fill": "fill_kernel", "classify": "band_classify_kernel"}
        line = {
            "metric": "frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong" if wl in ("c4", "c5") else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict({"workload": c["name"], "target": f"{w}x{h}", "frames_per_step": frames_per_step,
                            "clears": "every frame clears its targets; the clears are fused into the render calls (euc_render_clear)"}, **config_extra),
            "mfrag_per_s": frags_per_frame * job_mult * steps / (ms / 1000.0) / 1e6,
            "fragments_per_step": frags_per_frame,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "ms_per_step": ms_e2e / e2e_steps, "api": "C ABI: euc_render / euc_geom_update(_range) with pinned host pointers, euc_buf_download(_rows)_async + euc_ticket_wait"},
            "gpu_launches": int(launches),
            "host_waits_in_timed_region": int(host_waits),
            "cuda_graph_ms_per_step": graph_ms,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kernel_names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": per_launch_bytes, "avg_launch_ms": avg_s * 1000.0, "launches_per_frame": dom_per_frame,
                         "peak_source": peak_src, "frame_frac": (alg / ((ms / steps) / 1000.0) / 1e9) / peak / (1 if wl != "c5" else 1),
                         "limiter": "instruction issue, not HBM (ncu: issue-active ~80 %, DRAM < 10 % of peak; profiles/)"},
            "parity": {"checked_against": "oracle/ (C++ restatement of euc's render path) via the golden CRCs of tests/golden", "oracle_pinned": False,
                       "why": "the reference is a Rust crate; no rustc / cargo here, so real euc never ran (DESIGN.md section 6)"},
            "stage_ms_per_launch": stage_ms,
            "stage_timing": "CUDA events around every kernel launch, " + ("inside the timed region" if prof_in_region else f"separate run of {prof_steps} frames (N > 1: the timed region carries no event pairs)"),
        }
        line.update(verified)
        if world == 1 and cpu_baseline and not args.no_cpu_baseline:
            try:
                run, scale, desc, cores = cpu_plan(wl, scene)
                run()
                t_cpu = min(run() for _ in range(2)) * scale
                line["cpu_baseline"] = {"value": 1.0 / t_cpu, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc + "; best of 2 after 1 warm-up"}
            except Exception as ex:  # the oracle is a reported baseline, never a dependency of the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": None, "kind": "port", "sample": f"failed: {ex}"}
    barrier()
    for cl in closers:
        cl()
    for _, _, dr in slots:
        dr()
    del slots, keep
    while _PINNED:
        cx, ptr = _PINNED.pop()
        cx.host_free(ptr)
    return line


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))  # process plumbing: barriers and max-reductions of timings
    line = measure(args.workload, args, rank, world, local_rank)
    if args.workload == "c4" and not args.no_icon_batch:
        # BASELINE config 5 rides along in the same record: 4096 icons strong-sharded over the job's GPUs
        ib = measure("c5", args, rank, world, local_rank, steps=args.icon_steps if args.icon_steps is not None else 20 * world, warm=3, cpu_baseline=False)
        if rank == 0:
            line["icon_batch"] = {"workload": ib["config"]["workload"], "icons": args.icons, "n_gpus": world, "metric": "icons_per_s", "value": ib["value"], "unit": "icons/s",
                                  "ms_per_batch": ib["ms_per_step"], "e2e": ib["e2e"], "scaling": "strong", "partition": ib["config"]["partition"],
                                  "gpu_launches": ib["gpu_launches"], "stage_ms_per_launch": ib["stage_ms_per_launch"], "fragments_per_batch": ib["fragments_per_step"],
                                  "host_waits_in_timed_region": ib["host_waits_in_timed_region"], "crc_ok": ib["frame_matches_golden_crc"], "icons_checked": ib["icons_checked"],
                                  "icons_depth_crc_mismatch": ib["icons_depth_crc_mismatch"], "icons_colour_crc_mismatch": ib["icons_colour_crc_mismatch"],
                                  "roofline": ib["roofline"], "steps": ib["steps"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
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
