// euc_b200.cu — C ABI (include/euc_b200.h) over the sm_100a kernels in kernels.cuh.
// There is no CPU fallback: every entry point that does work launches CUDA kernels on the context's stream.
//
// Build (see __graft_entry__.build()):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -shared -Xcompiler -fPIC
#include "kernels.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <unordered_map>
#include <array>
#include <vector>

using namespace eucb;

namespace {

struct Buf {
    void* d = nullptr;
    uint32_t w = 0, h = 0, layers = 0, texel = 0;
    size_t bytes = 0;
    bool owned = true;
    bool ipc = false;  // mapped from another process (cudaIpcOpenMemHandle)
};
struct Geom {
    uint8_t* verts = nullptr;
    uint32_t stride = 0, n_verts = 0;
    uint32_t* idx = nullptr;
    uint32_t n_idx = 0;
    bool owned = true;
    // smallest / largest index, when the library has seen the indices on the host (euc_geom_create, small euc_render
    // streams): a render whose draws are in range by these bounds cannot fail with EUC_E_OUT_OF_BOUNDS and therefore has
    // nothing to wait for.  Unknown bounds: the device checks, and the error is deferred (see euc_set_async).
    bool bounds_known = false;
    uint32_t idx_min = 0, idx_max = 0;
    // A draw table (batch rendering: ranges + base vertices) that the global bounds cannot prove in range is checked once on
    // the device; the same table on the same indices need not be checked again.
    uint64_t version = 0;
    mutable uint64_t verified_hash = 0, verified_version = ~0ull;
};
uint64_t fnv1a(const void* data, size_t n) {
    const uint8_t* b = (const uint8_t*)data;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h ? h : 1;
}
// Index bounds of a host index array (one pass; auto-vectorised)
void index_bounds(const uint32_t* idx, size_t n, uint32_t& lo, uint32_t& hi) {
    uint32_t a = 0xffffffffu, b = 0u;
    for (size_t i = 0; i < n; ++i) { a = idx[i] < a ? idx[i] : a; b = idx[i] > b ? idx[i] : b; }
    lo = n ? a : 0u; hi = b;
}
constexpr size_t HOST_SCAN_MAX_INDICES = 1u << 16;  // euc_render / euc_geom_update scan host indices up to this many per call
struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

namespace { struct PipeOps; }
struct euc_user_pipes;
struct euc_group;
struct euc_ctx {
    euc_group* group = nullptr;  // multi-GPU group this context belongs to (group.inc)
    std::unordered_map<int, PipeOps*> builtin;  // kernel launchers of the built-in pipelines used so far (key: pipeline id * 2 + lines)
    euc_user_pipes* user = nullptr;  // pipelines compiled at run time (runtime_pipeline.inc)
    int dev = 0;
    cudaStream_t own = nullptr, stream = nullptr;
    std::string err;
    uint64_t next_handle = 1;
    std::unordered_map<uint64_t, Buf> bufs;
    std::unordered_map<uint64_t, Geom> geoms;
    Scratch recs, bbox, tile_count, tile_range, tile_list, draws, uniforms, tmp_verts, tmp_idx, winner, ovf, ext;
    unsigned long long* counters = nullptr;      // device, CTR_WORDS words (kernels.cuh: Params::counters); word 15 = sticky flags
    unsigned long long* counters_host = nullptr;  // pinned
    // Asynchronous renders (euc_set_async, default on): no host wait inside a render call.  The last raster warp of a render
    // publishes its summary into `summary` (mapped pinned memory); the host reads it at the start of later calls.
    bool async = true;
    volatile unsigned long long* summary = nullptr;      // host view
    unsigned long long* summary_dev = nullptr;           // device view of the same memory
    unsigned long long seq = 0, seen_seq = 0;
    uint64_t ovf_want = 0;                                // overflow-buffer entries wanted by past renders
    int deferred_code = EUC_OK;                           // error of an earlier asynchronous render, reported by the next call
    std::string deferred_msg;
    int launch_err = 0;                                   // first failed driver-API launch of a run-time pipeline (CUresult)
    uint64_t blocking_waits = 0;                          // host waits inside render calls (diagnostics: 0 in steady state)
    void* stage_host[2] = {nullptr, nullptr}; size_t stage_cap[2] = {0, 0}; cudaEvent_t stage_ev[2] = {nullptr, nullptr};  // pinned staging of batch tables
    bool stage_used[2] = {false, false}; int stage_next = 0;
    bool stats = false;
    struct ClearReq { bool px = false, z = false; uint32_t px_value = 0, z_value = 0; } next_clear;  // euc_render_clear: consumed by the next render
    int sparse_recs = -1;  // EUC_SPARSE_RECS (development): -1 = automatic, 0 / 1 = force
    // Smallest group whose ranks classify the primitives together (sort-middle of ids, group.inc) instead of each setting up
    // the whole stream; 0 = never.  EUC_GROUP_CLS_MIN_WORLD overrides (tests use 2).
    uint32_t group_cls_min_world = 0;
    uint64_t ovf_fixed = 0;  // EUC_OVF_ENTRIES (tests): fixed size of the bin-overflow buffer instead of the adaptive one
    euc_render_stats last{};
    bool stats_on_device = false;  // the fragment counter of the last render lives in counters[1]
    int sm_count = 148;
    uint64_t launches = 0;
    std::unordered_map<uint64_t, cudaEvent_t> tickets;  // pending asynchronous read-backs
    uint64_t next_ticket = 1;
    bool profiling = false;
    cudaEvent_t ev_counts = nullptr, ev_setup = nullptr;
    cudaStream_t aux = nullptr;  // counter read-back
    struct BinHint { uint32_t cap = 128; bool verified = false; uint32_t n_tris = 0; };  // cap 0 = use the exact path; n_tris: primitives of the last checked render
    std::unordered_map<uint32_t, BinHint> bin_hint;       // per tile-count: bin size of the fast path
    std::vector<std::array<cudaEvent_t, 2>> pending[EUC_STAGE_COUNT];  // recorded, not yet read
    std::vector<cudaEvent_t> ev_pool;
    float prof_ms[EUC_STAGE_COUNT] = {};
    uint64_t prof_calls[EUC_STAGE_COUNT] = {};
};

namespace {

int fail(euc_ctx* c, int code, const char* fmt, ...) {
    if (c) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        c->err = buf;
    }
    return code;
}
#define CU(call)                                                                                                  \
    do {                                                                                                          \
        cudaError_t e_ = (call);                                                                                  \
        if (e_ != cudaSuccess) return fail(ctx, e_ == cudaErrorMemoryAllocation ? EUC_E_OOM : EUC_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

constexpr int CTR_WORDS = 16;        // device counters of a context (kernels.cuh: Params::counters)
constexpr int CTR_CLEAR_WORDS = 12;  // zeroed per render; word 15 keeps the sticky flags until the host has seen them

// Reads the summary of the most recent finished asynchronous render (mapped pinned memory, no CUDA call): grows the bin /
// overflow sizes for later renders and latches device-detected errors, which the next API call reports.
void poll_summary(euc_ctx* ctx) {
    if (!ctx->summary) return;
    if (ctx->summary[6]) {  // a group barrier gave up waiting for a peer
        ctx->summary[6] = 0;
        if (ctx->deferred_code == EUC_OK) { ctx->deferred_code = EUC_E_CUDA; ctx->deferred_msg = "a group barrier timed out: a peer rank never arrived"; }
    }
    const unsigned long long s0 = ctx->summary[0];
    if (s0 == ctx->seen_seq) return;
    const unsigned long long flags = ctx->summary[1], longest = ctx->summary[2], ovf = ctx->summary[3], tiles = ctx->summary[5];
    if (ctx->summary[0] != s0) return;  // a newer render is publishing right now: look again at the next call
    ctx->seen_seq = s0;
    auto it = ctx->bin_hint.find((uint32_t)tiles);
    if (it != ctx->bin_hint.end() && it->second.cap) {
        const unsigned long long want = ((longest + longest / 4 + 32) + 31) / 32 * 32;
        if (want > it->second.cap) it->second.cap = (want * tiles * 4 <= (1ull << 30)) ? (uint32_t)want : 0u;
    }
    if (ovf * 2 > ctx->ovf_want) ctx->ovf_want = ovf * 2;
    if (flags) {
        // bit 2 (overflow buffer exhausted) needs no report: the affected tiles were rendered by scanning all primitives, and
        // the buffer is larger from now on
        if ((flags & 1ull) && ctx->deferred_code == EUC_OK) { ctx->deferred_code = EUC_E_OUT_OF_BOUNDS; ctx->deferred_msg = "vertex index out of range in an earlier asynchronous render (that render drew nothing)"; }
        cudaMemsetAsync(ctx->counters + 15, 0, sizeof(unsigned long long), ctx->stream);
    }
}

struct DeviceGuard {  // every entry point works on its context's device, whatever device the calling thread had current
    int prev = -1; bool switched = false;
    explicit DeviceGuard(euc_ctx* ctx) { if (ctx && cudaGetDevice(&prev) == cudaSuccess && prev != ctx->dev) switched = cudaSetDevice(ctx->dev) == cudaSuccess; }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

int ensure(euc_ctx* ctx, Scratch& s, size_t bytes, bool zero_new = false) {
    if (bytes <= s.cap) return EUC_OK;
    {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(ctx->stream, &st);
        if (st != cudaStreamCaptureStatusNone) return fail(ctx, EUC_E_UNSUPPORTED, "a scratch buffer must grow: run this render once outside stream capture");
    }
    size_t cap = bytes + bytes / 4 + 4096;
    if (s.p) {
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaFree(s.p));
        s.p = nullptr;
        s.cap = 0;
    }
    CU(cudaMalloc(&s.p, cap));
    s.cap = cap;
    if (zero_new) CU(cudaMemsetAsync(s.p, 0, cap, ctx->stream));
    return EUC_OK;
}

cudaEvent_t get_event(euc_ctx* ctx) {
    if (!ctx->ev_pool.empty()) { cudaEvent_t e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
struct StageTimer {  // brackets one kernel launch with events when profiling is on
    euc_ctx* ctx; int stage; cudaEvent_t a = nullptr;
    StageTimer(euc_ctx* c, int s) : ctx(c), stage(s) {
        ++ctx->launches;
        if (ctx->profiling) { a = get_event(ctx); cudaEventRecord(a, ctx->stream); }
    }
    ~StageTimer() {
        if (a) { cudaEvent_t b = get_event(ctx); cudaEventRecord(b, ctx->stream); ctx->pending[stage].push_back({a, b}); }
    }
};
void drain_profile(euc_ctx* ctx) {  // stream must be synchronised
    for (int s = 0; s < EUC_STAGE_COUNT; ++s) {
        for (auto& pr : ctx->pending[s]) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, pr[0], pr[1]) == cudaSuccess) { ctx->prof_ms[s] += ms; ctx->prof_calls[s] += 1; }
            ctx->ev_pool.push_back(pr[0]);
            ctx->ev_pool.push_back(pr[1]);
        }
        ctx->pending[s].clear();
    }
}

template <class P> struct PipeInfo {
    static constexpr bool needs_sampler = false;
    static constexpr int sampler_format = -1;
    static constexpr bool vec4_loads = true;
};
template <> struct PipeInfo<PipeTeapotShadow> { static constexpr bool needs_sampler = false; static constexpr int sampler_format = -1; static constexpr bool vec4_loads = false; };
template <> struct PipeInfo<PipeWireframe> { static constexpr bool needs_sampler = false; static constexpr int sampler_format = -1; static constexpr bool vec4_loads = false; };
template <> struct PipeInfo<PipeTeapotPhong> { static constexpr bool needs_sampler = true; static constexpr int sampler_format = EUC_TEXEL_F32; static constexpr bool vec4_loads = false; };
template <> struct PipeInfo<PipeTexCube> { static constexpr bool needs_sampler = true; static constexpr int sampler_format = EUC_TEXEL_RGBA8_TO_F32; static constexpr bool vec4_loads = true; };

struct RenderCall {
    const euc_pipeline_desc* desc;
    const Geom* geom;
    const euc_batch_draw* draws;
    uint32_t n_draws;
    const void* uniforms;  // host: n_draws blocks (batch) or one block
    bool batch;
    euc_buf pixel, depth;
    uint32_t row_begin, row_end;
    const euc_buf* mirrors = nullptr;
    uint32_t n_mirrors = 0;
    bool maybe_oob_sync = false;  // the host knows the index bounds and they do not prove every draw in range: checked render
    bool group_cls = false;       // euc_group_render: the ranks classify 1/world of the primitives each and exchange ids (group.inc)
};

// What the render driver needs from a pipeline: sizes, flags, and how to launch its kernels on ctx->stream.
struct PipeOps {
    uint32_t rec_bytes = 0, vertex_bytes = 0, uniform_bytes = 0;
    bool has_fragment = false, defer = false, needs_sampler = false, vec4_loads = false;
    int sampler_format = -1;
    std::function<void(const Params&, uint32_t blocks)> setup;
    std::function<void(const Params&, uint32_t blocks)> group_classify;  // the same across a group (optional)
    std::function<void(const Params&, bool msaa, uint32_t blocks, uint32_t n_tiles)> raster;
    std::function<void(const Params&, bool msaa, uint32_t grid)> resolve;
    std::function<int(bool msaa)> resident;
    int resident_cache[2] = {0, 0};  // per context (the smem attribute must be set on every device)
};

// Fills rows [row_begin, row_end) of every layer of a buffer (Target::clear restricted to a row range).
static int clear_rows_impl(euc_ctx* ctx, const Buf& b, uint32_t v, uint32_t row_begin, uint32_t row_end) {
    row_end = std::min(row_end, b.h);
    if (row_begin >= row_end || b.w == 0) return EUC_OK;
    if (row_begin == 0 && row_end == b.h) {  // whole buffer: one launch over all layers
        const size_t n = b.bytes / 4, vec = (n + 3) / 4;
        if (n == 0) return EUC_OK;
        const unsigned blocks = (unsigned)std::min<size_t>((vec + 255) / 256, (size_t)ctx->sm_count * 16);
        ++ctx->launches;
        fill_u32_kernel<<<blocks, 256, 0, ctx->stream>>>((uint32_t*)b.d, n, v);
        CU(cudaGetLastError());
        return EUC_OK;
    }
    for (uint32_t l = 0; l < b.layers; ++l) {
        uint32_t* base = (uint32_t*)b.d + ((size_t)l * b.h + row_begin) * b.w;
        const size_t n = (size_t)(row_end - row_begin) * b.w;
        const size_t vec = (n + 3) / 4;
        const unsigned blocks = (unsigned)std::min<size_t>((vec + 255) / 256, (size_t)ctx->sm_count * 16);
        ++ctx->launches;
        fill_u32_kernel<<<blocks, 256, 0, ctx->stream>>>(base, n, v);
    }
    CU(cudaGetLastError());
    return EUC_OK;
}

// A euc_render_clear request belongs to the next render call only: whatever way that call ends, the request is gone.
struct DropClear {
    euc_ctx* ctx;
    ~DropClear() { if (ctx) ctx->next_clear = euc_ctx::ClearReq{}; }
};

// group.inc: fills the cls_* / surv_* fields of prm for a render whose primitives the group classifies together; launches
// the group barrier that carries this rank's per-destination counts.
int group_cls_prepare(euc_ctx* ctx, Params& prm);
int group_cls_barrier(euc_ctx* ctx);

// Pipeline-agnostic render driver.  `ops` describes the pipeline (record size, flags) and launches its kernels: template
// instantiations for the built-in pipelines, NVRTC-compiled modules for pipelines registered at run time.
int render_driver(euc_ctx* ctx, const RenderCall& rc, Params& prm, uint32_t n_tiles, const PipeOps& ops) {
    const euc_pipeline_desc& d = *rc.desc;
    if (rc.geom->stride < ops.vertex_bytes || (rc.geom->stride & 3u)) return fail(ctx, EUC_E_INVALID, "vertex stride %u too small / unaligned for pipeline %d", rc.geom->stride, d.pipeline_id);
    if (ops.vec4_loads && (rc.geom->stride & 15u)) return fail(ctx, EUC_E_INVALID, "vertex stride must be a multiple of 16 for pipeline %d", d.pipeline_id);
    if (d.uniform_bytes < ops.uniform_bytes) return fail(ctx, EUC_E_INVALID, "uniform block too small (%u < %u)", d.uniform_bytes, ops.uniform_bytes);
    if (ops.needs_sampler && prm.pixel_write) {
        if (!prm.samp[0].data) return fail(ctx, EUC_E_INVALID, "pipeline %d needs sampler 0", d.pipeline_id);
        if (ops.sampler_format >= 0 && prm.samp[0].format != ops.sampler_format) return fail(ctx, EUC_E_INVALID, "sampler 0 has the wrong texel format for pipeline %d", d.pipeline_id);
    }
    int rcode;
    if ((rcode = ensure(ctx, ctx->recs, (size_t)prm.n_tris * ops.rec_bytes)) != EUC_OK) return rcode;
    prm.recs = (uint32_t*)ctx->recs.p;

    if ((prm.clear_mask & 1u) && !ops.has_fragment) {
        // no fragment stage: the tile kernel never touches the colour target, so its clear is a plain fill
        Buf pxb{}; pxb.d = prm.pixel; pxb.w = prm.w; pxb.h = prm.h; pxb.layers = prm.layers; pxb.bytes = (size_t)prm.w * prm.h * prm.layers * 4;
        int crc = clear_rows_impl(ctx, pxb, prm.clear_px, prm.row_begin, prm.row_end);
        if (crc != EUC_OK) return crc;
        prm.clear_mask &= ~1u;
    }
    const uint32_t tri_blocks = (prm.n_tris + 127) / 128;
    const uint32_t rblocks = (n_tiles + RASTER_WARPS - 1) / RASTER_WARPS;
    const bool msaa = prm.msaa_level > 0 && ops.has_fragment && prm.pixel_write;
    const int resident = ops.resident(msaa);  // CTAs of the raster kernel that fit one SM (sets its smem attribute once)
    if (resident <= 0) return fail(ctx, EUC_E_CUDA, "raster kernel occupancy query failed");
    // persistent grid: one resident set of CTAs (or one warp per tile if there are fewer tiles); warps take tiles from a
    // ticket counter (counters[4], zeroed per render).  A "balanced" grid (every warp the same number of tiles, fewer warps)
    // was measured against this with the 6-CTA kernel and lost at every size: 3 % on the full C4 frame, 9 % on one eighth
    // of it (tools/band_probe.py): early finishers that pick up the remaining tiles run them at low occupancy, i.e. fast.
    const uint32_t ty_lo = prm.row_begin / TILE, ty_hi = (std::min(prm.row_end, prm.h) + TILE - 1) / TILE;
    const uint64_t n_active = (uint64_t)(ty_hi - ty_lo) * prm.tiles_x * prm.layers;
    uint32_t pblocks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((n_active + RASTER_WARPS - 1) / RASTER_WARPS, (uint64_t)ctx->sm_count * resident));
    if (prm.n_mirrors) {
        // fused gather (peer stores over NVLink at every tile's write-back): here the balanced grid wins (N = 8: raster
        // 0.133 vs 0.149 ms) because fewer warps finish their tiles, and burst their rows onto the links, at the same time
        const uint64_t max_warps = (uint64_t)ctx->sm_count * resident * RASTER_WARPS;
        const uint64_t tiles_per_warp = std::max<uint64_t>(1, (n_active + max_warps - 1) / max_warps);
        const uint64_t want_warps = (n_active + tiles_per_warp - 1) / tiles_per_warp;
        pblocks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((want_warps + RASTER_WARPS - 1) / RASTER_WARPS, (uint64_t)ctx->sm_count * resident));
    }
    (void)rblocks;
    const bool resolve = ops.defer && prm.pixel_write;  // deferred pipelines: raster records winners, resolve_kernel shades
    if (resolve) {
        const size_t wb = (size_t)prm.w * prm.h * prm.layers * 4;
        if (wb > ctx->winner.cap) {  // new allocation: fill with NO_WINNER once; resolve_kernel keeps it clean afterwards
            if ((rcode = ensure(ctx, ctx->winner, wb)) != EUC_OK) return rcode;
            CU(cudaMemsetAsync(ctx->winner.p, 0xff, ctx->winner.cap, ctx->stream));
        }
        prm.winner = (uint32_t*)ctx->winner.p;
    }
    auto launch_raster = [&]() {
        ctx->stats_on_device = true;
        { StageTimer t(ctx, EUC_STAGE_RASTER); ops.raster(prm, msaa, pblocks, n_tiles); }
        if (resolve) {
            const uint32_t row_end = std::min(prm.row_end, prm.h), rows = row_end - prm.row_begin;
            uint64_t nblocks;
            if (msaa) {  // resolve_kernel<MSAA>: one thread per cell of the shading grid, CTAs of 32 x 4 cells (same arithmetic as the kernel)
                const uint32_t cs = 1u << prm.msaa_level;
                const uint32_t cells_x = (prm.w + cs - 1) >> prm.msaa_level, cells_per_band = (prm.group_rows + cs - 1) >> prm.msaa_level;
                const uint32_t bands = rows ? (row_end - 1) / prm.group_rows - prm.row_begin / prm.group_rows + 1 : 0u;
                nblocks = (uint64_t)((cells_x + 31) / 32) * (((uint64_t)bands * cells_per_band + 3) / 4) * prm.layers;
            } else {
                nblocks = (uint64_t)((prm.w + 31) / 32) * ((rows + 3) / 4) * prm.layers;
            }
            // light shaders: grid-stride over a machine-sized grid (most blocks hold no winner); heavy MSAA shading: one CTA
            // per block so that the hardware balances the expensive blocks dynamically
            const uint32_t grid = msaa ? (uint32_t)std::min<uint64_t>(nblocks, 0x7fffffffull) : (uint32_t)std::min<uint64_t>(nblocks, (uint64_t)ctx->sm_count * 16);
            StageTimer t(ctx, EUC_STAGE_RESOLVE);
            if (grid > 0) ops.resolve(prm, msaa, grid);
        }
    };
    auto fetch_counters = [&]() -> int {
        // The host needs the pair count / flags that setup (and alloc) produced.  The tiny D2H copy runs on an auxiliary
        // stream that waits for the kernels queued so far, so the raster kernel queued next on the main stream does
        // not sit behind the copy engine's latency.  The caller waits on ev_counts before it returns.
        CU(cudaEventRecord(ctx->ev_setup, ctx->stream));
        CU(cudaStreamWaitEvent(ctx->aux, ctx->ev_setup, 0));
        CU(cudaMemcpyAsync(ctx->counters_host, ctx->counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->aux));
        CU(cudaEventRecord(ctx->ev_counts, ctx->aux));
        return EUC_OK;
    };
    cudaStreamCaptureStatus cap_st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(ctx->stream, &cap_st);
    const bool capturing = cap_st != cudaStreamCaptureStatusNone;

    // ---- fast path: fixed-capacity bins.  setup appends primitive ids straight into tile*cap + slot; raster follows.
    // The capacity is a per-context hint (longest list seen, with headroom).
    //   * asynchronous (steady state): a pair that does not fit its bin goes to the overflow buffer and the raster warp of
    //     that tile collects it from there, so the result is right whatever the hint was and the host waits for nothing.
    //     The render's summary (longest list, overflow volume, flags) reaches the host through mapped pinned memory and
    //     sizes later renders.
    //   * checked (first render of a target shape, renders that could read a vertex out of range, euc_set_async(0)): the
    //     host waits for setup's flags; an overflowing bin makes raster exit at once and the render is redone on the exact
    //     path below, which also measures the longest list.
    euc_ctx::BinHint hint;
    {
        auto it = ctx->bin_hint.find(n_tiles);
        if (it != ctx->bin_hint.end()) hint = it->second;
    }
    const uint32_t cap = hint.cap;
    const bool fast = cap > 0 && (size_t)n_tiles * cap * 4 <= ((size_t)1 << 30);
    // asynchronous only for a shape that a checked render has seen: same tile grid AND same primitive count (a frame loop over
    // one scene); another scene on the same target is checked once first, so that an arbitrarily denser scene is sized by
    // the exact path instead of overrunning the overflow buffer
    const bool go_async = fast && ctx->async && hint.verified && hint.n_tris == prm.n_tris && !rc.maybe_oob_sync;
    if (capturing && !go_async)
        return fail(ctx, EUC_E_UNSUPPORTED, "this render needs a host check (first render of a target shape, or index bounds the host cannot prove): run it once outside stream capture");
    if (fast) {
        if ((rcode = ensure(ctx, ctx->tile_list, (size_t)n_tiles * cap * 4)) != EUC_OK) return rcode;
        prm.tile_list = (uint32_t*)ctx->tile_list.p;
        prm.list_capacity = (uint32_t)std::min<size_t>(ctx->tile_list.cap / 4, 0xfffffff0u);
        prm.bin_cap = cap;
        if (ctx->async) {
            // sized on the checked render already: the asynchronous renders that follow (possibly inside stream capture) find them
            const uint64_t want = ctx->ovf_fixed ? ctx->ovf_fixed : std::max<uint64_t>({(uint64_t)1 << 18, (uint64_t)prm.n_tris * 2, ctx->ovf_want});
            if ((rcode = ensure(ctx, ctx->ovf, (size_t)want * sizeof(uint2))) != EUC_OK) return rcode;
            if ((rcode = ensure(ctx, ctx->ext, (size_t)want * 4)) != EUC_OK) return rcode;
        }
        if (go_async) {
            prm.ovf = (uint2*)ctx->ovf.p;
            prm.ext = (uint32_t*)ctx->ext.p;
            prm.ovf_cap = (uint32_t)std::min<uint64_t>(std::min(ctx->ovf.cap / sizeof(uint2), ctx->ext.cap / 4), 0x7fffffffull);
            prm.summary = ctx->summary_dev;
            prm.seq = ++ctx->seq;
        }
        // a rank's band of a large frame: list the primitives that meet the band first (light kernel, full occupancy), set up those
        // (inside a group the ranks share that work: each classifies 1/world of the primitives and sends the ids where they belong)
        const bool gcls = rc.group_cls && ops.group_classify && prm.n_tris > (1u << 16);
        if (gcls) {
            if ((rcode = group_cls_prepare(ctx, prm)) != EUC_OK) return rcode;
            prm.sparse_recs = 1;
        }
        CU(cudaMemsetAsync(ctx->counters, 0, CTR_CLEAR_WORDS * sizeof(unsigned long long), ctx->stream));
        if (gcls) {
            CU(cudaMemsetAsync(prm.cls_counts, 0, EUC_MAX_GROUP * sizeof(uint32_t), ctx->stream));
            { StageTimer t(ctx, EUC_STAGE_CLASSIFY); ops.group_classify(prm, (prm.cls_n + 255) / 256); }
            if ((rcode = group_cls_barrier(ctx)) != EUC_OK) return rcode;
        }
        // list mode: a machine-sized grid strides over the list (its length is only known on the device)
        const uint32_t setup_blocks = gcls ? std::min<uint32_t>(tri_blocks, (uint32_t)ctx->sm_count * 5u) : tri_blocks;
        { StageTimer t(ctx, EUC_STAGE_SETUP); ops.setup(prm, setup_blocks); }
        if (go_async) {
            launch_raster();
            CU(cudaGetLastError());
            if (ctx->launch_err) { const int le = ctx->launch_err; ctx->launch_err = 0; return fail(ctx, EUC_E_CUDA, "kernel launch of a run-time pipeline failed (CUresult %d)", le); }
            ctx->last.primitives = prm.n_tris;
            ctx->last.binned_pairs = 0;  // on the device; euc_get_stats reads it
            ctx->last.fragments = 0;
            return EUC_OK;
        }
        if ((rcode = fetch_counters()) != EUC_OK) return rcode;
        launch_raster();
        CU(cudaGetLastError());
        if (ctx->launch_err) { const int le = ctx->launch_err; ctx->launch_err = 0; return fail(ctx, EUC_E_CUDA, "kernel launch of a run-time pipeline failed (CUresult %d)", le); }
        ++ctx->blocking_waits;
        CU(cudaEventSynchronize(ctx->ev_counts));  // waits for setup only; raster is already queued behind it
        ctx->last.primitives = prm.n_tris;
        ctx->last.binned_pairs = ctx->counters_host[0];
        ctx->last.fragments = 0;
        if (ctx->counters_host[3] & 1ull) {
            CU(cudaMemsetAsync(ctx->tile_count.p, 0, (size_t)n_tiles * 4, ctx->stream));
            return fail(ctx, EUC_E_OUT_OF_BOUNDS, "vertex index out of range");
        }
        if (!(ctx->counters_host[3] & 2ull)) {
            hint.verified = true;  // this shape fits bins of `cap`: later renders run asynchronously
            hint.n_tris = prm.n_tris;
            ctx->bin_hint[n_tiles] = hint;
            return EUC_OK;
        }
        // overflow: raster skipped itself.  Reset the tile counters and fall through to the exact path, which also
        // measures the longest list for the next render's capacity.
        CU(cudaMemsetAsync(ctx->tile_count.p, 0, (size_t)n_tiles * 4, ctx->stream));
    }

    // ---- exact path: count (setup) -> alloc -> fill -> raster.  The pair list is sized optimistically (grow-only) so
    // that fill and raster can be queued before the host knows the pair count: the GPU never waits for the host.
    // alloc_tiles flags a list that is too small, fill/raster then exit immediately, and the host re-launches them.
    prm.bin_cap = 0;
    prm.ovf_cap = 0;
    prm.summary = nullptr;
    prm.survivors = nullptr;
    if (rc.group_cls) prm.sparse_recs = ctx->sparse_recs >= 0 ? (uint32_t)ctx->sparse_recs : ((uint64_t)(prm.row_end - prm.row_begin) * 2 < prm.h ? 1u : 0u);
    if ((rcode = ensure(ctx, ctx->tile_list, std::max<size_t>((size_t)prm.n_tris * 3, 1u << 16) * 4)) != EUC_OK) return rcode;
    prm.tile_list = (uint32_t*)ctx->tile_list.p;
    prm.list_capacity = (uint32_t)std::min<size_t>(ctx->tile_list.cap / 4, 0xfffffff0u);

    CU(cudaMemsetAsync(ctx->counters, 0, CTR_CLEAR_WORDS * sizeof(unsigned long long), ctx->stream));
    auto launch_fill_raster = [&]() {
        { StageTimer t(ctx, EUC_STAGE_FILL); fill_kernel<<<tri_blocks, 128, 0, ctx->stream>>>(prm); }
        launch_raster();
    };
    { StageTimer t(ctx, EUC_STAGE_SETUP); ops.setup(prm, tri_blocks); }
    { StageTimer t(ctx, EUC_STAGE_ALLOC); alloc_tiles_kernel<<<(n_tiles + 255) / 256, 256, 0, ctx->stream>>>(prm, n_tiles); }
    if ((rcode = fetch_counters()) != EUC_OK) return rcode;
    // counters[1] held the longest list for the host; raster accumulates the fragment count there.  The reset must
    // follow the read-back, which runs on the auxiliary stream.
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_counts, 0));
    CU(cudaMemsetAsync(ctx->counters + 1, 0, sizeof(unsigned long long), ctx->stream));
    launch_fill_raster();
    CU(cudaGetLastError());
    if (ctx->launch_err) { const int le = ctx->launch_err; ctx->launch_err = 0; return fail(ctx, EUC_E_CUDA, "kernel launch of a run-time pipeline failed (CUresult %d)", le); }
    ++ctx->blocking_waits;
    CU(cudaEventSynchronize(ctx->ev_counts));  // waits for setup + alloc only; fill and raster are already queued behind
    const unsigned long long pairs = ctx->counters_host[0];
    ctx->last.primitives = prm.n_tris;
    ctx->last.binned_pairs = pairs;
    ctx->last.fragments = 0;
    if (ctx->counters_host[3] & 1ull) {
        // reference: slice index panic (index.rs:53).  fill/raster skipped themselves; restore the all-zero tile counters.
        CU(cudaMemsetAsync(ctx->tile_count.p, 0, (size_t)n_tiles * 4, ctx->stream));
        return fail(ctx, EUC_E_OUT_OF_BOUNDS, "vertex index out of range");
    }
    {   // capacity hint for the fast path of the next render with this tile configuration
        const unsigned long long longest = ctx->counters_host[1];
        const unsigned long long want = ((longest + longest / 4 + 32) + 31) / 32 * 32;
        hint.cap = (want * n_tiles * 4 <= (1ull << 30)) ? (uint32_t)want : 0u;
        hint.verified = hint.cap != 0;
        hint.n_tris = prm.n_tris;
        ctx->bin_hint[n_tiles] = hint;
    }
    if (ctx->counters_host[3] & 2ull) {
        if (pairs > 0xfffffff0ull) {
            CU(cudaMemsetAsync(ctx->tile_count.p, 0, (size_t)n_tiles * 4, ctx->stream));
            return fail(ctx, EUC_E_UNSUPPORTED, "too many (tile, primitive) pairs: %llu", pairs);
        }
        if ((rcode = ensure(ctx, ctx->tile_list, (size_t)pairs * 4)) != EUC_OK) return rcode;
        prm.tile_list = (uint32_t*)ctx->tile_list.p;
        prm.list_capacity = (uint32_t)std::min<size_t>(ctx->tile_list.cap / 4, 0xfffffff0u);
        CU(cudaMemsetAsync(ctx->counters + 3, 0, 2 * sizeof(unsigned long long), ctx->stream));  // flags and tile ticket
        launch_fill_raster();
        CU(cudaGetLastError());
    }
    return EUC_OK;
}

// PipeOps of a built-in pipeline: template instantiations of the kernels.  One set per context and pipeline (the shared
// memory attribute of the raster kernel is per device, and two contexts may be used in turn by one thread).
template <class P, bool LINES = false> const PipeOps& builtin_ops(euc_ctx* ctx, int key) {
    auto found = ctx->builtin.find(key);
    if (found != ctx->builtin.end()) return *found->second;
    PipeOps* po = new PipeOps();
    PipeOps& ops = *po;
    using PI = PipeInfo<P>;
    constexpr bool DEFER = P::HAS_FRAGMENT && P::BLEND_IGNORES_OLD;
    ops.rec_bytes = RecLayout<P>::BYTES;
    ops.vertex_bytes = P::VERTEX_BYTES;
    ops.uniform_bytes = std::is_same<P, PipeBlendTris>::value ? 0u : (uint32_t)sizeof(typename P::Uniforms);
    ops.has_fragment = P::HAS_FRAGMENT; ops.defer = DEFER;
    ops.needs_sampler = PI::needs_sampler; ops.sampler_format = PI::sampler_format; ops.vec4_loads = PI::vec4_loads;
    ops.setup = [ctx](const Params& prm, uint32_t blocks) {
        if (LINES) setup_lines_kernel<P><<<blocks, 128, 0, ctx->stream>>>(prm);
        else if (prm.survivors) setup_kernel<P, true><<<blocks, 128, 0, ctx->stream>>>(prm);
        else setup_kernel<P, false><<<blocks, 128, 0, ctx->stream>>>(prm);
    };
    if (!LINES) ops.group_classify = [ctx](const Params& prm, uint32_t blocks) { group_classify_kernel<P><<<blocks, 256, 0, ctx->stream>>>(prm); };
    ops.raster = [ctx](const Params& prm, bool msaa, uint32_t blocks, uint32_t n_tiles) {
        auto kern = msaa ? raster_kernel<P, true, DEFER, LINES> : raster_kernel<P, false, DEFER, LINES>;
        kern<<<blocks, RASTER_WARPS * 32, raster_smem_bytes<P, DEFER>(), ctx->stream>>>(prm, n_tiles);
    };
    ops.resolve = [ctx](const Params& prm, bool msaa, uint32_t grid) {
        if (msaa) resolve_kernel<P, true, LINES><<<grid, 128, 0, ctx->stream>>>(prm);
        else resolve_kernel<P, false, LINES><<<grid, 128, 0, ctx->stream>>>(prm);
    };
    ops.resident = [po](bool msaa) -> int {
        int* res = po->resident_cache;
        if (!res[msaa]) {
            auto kern = msaa ? raster_kernel<P, true, DEFER, LINES> : raster_kernel<P, false, DEFER, LINES>;
            const size_t smem = raster_smem_bytes<P, DEFER>();
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, RASTER_WARPS * 32, smem) != cudaSuccess) return -1;
            res[msaa] = std::max(nb, 1);
        }
        return res[msaa];
    };
    ctx->builtin[key] = po;
    return ops;
}

template <class P, bool LINES = false> int render_typed(euc_ctx* ctx, const RenderCall& rc, Params& prm, uint32_t n_tiles) {
    return render_driver(ctx, rc, prm, n_tiles, builtin_ops<P, LINES>(ctx, rc.desc->pipeline_id * 2 + (LINES ? 1 : 0)));
}

}  // namespace
#include "runtime_pipeline.inc"
namespace {

// Error of an earlier asynchronous render, if the host has learnt of one: reported (once) by the next call.
int take_deferred(euc_ctx* ctx) {
    poll_summary(ctx);
    if (ctx->deferred_code == EUC_OK) return EUC_OK;
    const int c = ctx->deferred_code;
    ctx->deferred_code = EUC_OK;
    return fail(ctx, c, "%s", ctx->deferred_msg.c_str());
}

// Pinned staging for the per-draw tables of a batch (two slots, used in turn): the copies to the device are then truly
// asynchronous; a slot is rewritten only after the copy that last read it has finished, so the host runs up to two
// batches ahead of the device.  Returns the slot's base address in *out.
int stage_tables(euc_ctx* ctx, const void* a, size_t na, const void* b, size_t nb, size_t b_off, uint8_t** out) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(ctx->stream, &st);
    if (st != cudaStreamCaptureStatusNone) return fail(ctx, EUC_E_UNSUPPORTED, "multi-draw renders stage their tables through the host and cannot be captured into a CUDA graph");
    const int k = ctx->stage_next;
    ctx->stage_next ^= 1;
    const size_t total = b_off + nb;
    if (ctx->stage_ev[k] && ctx->stage_used[k]) CU(cudaEventSynchronize(ctx->stage_ev[k]));
    if (total > ctx->stage_cap[k]) {
        if (ctx->stage_host[k]) { CU(cudaFreeHost(ctx->stage_host[k])); ctx->stage_host[k] = nullptr; ctx->stage_cap[k] = 0; }
        CU(cudaMallocHost(&ctx->stage_host[k], total + total / 2 + 4096));
        ctx->stage_cap[k] = total + total / 2 + 4096;
    }
    if (!ctx->stage_ev[k]) CU(cudaEventCreateWithFlags(&ctx->stage_ev[k], cudaEventDisableTiming));
    if (na) std::memcpy(ctx->stage_host[k], a, na);
    if (nb) std::memcpy((uint8_t*)ctx->stage_host[k] + b_off, b, nb);
    *out = (uint8_t*)ctx->stage_host[k];
    return EUC_OK;
}

int render_common(euc_ctx* ctx, const RenderCall& rc_in) {
    ctx->last = euc_render_stats{};  // a render that returns before launching anything (quirks, empty targets) reports zeros
    ctx->stats_on_device = false;
    RenderCall rc = rc_in;
    if (!rc.desc || !rc.geom) return fail(ctx, EUC_E_INVALID, "null desc/geom");
    { const int dc = take_deferred(ctx); if (dc != EUC_OK) return dc; }
    const euc_pipeline_desc& d = *rc.desc;
    const UserPipe* user_pipe = nullptr;
    if (d.pipeline_id >= EUC_PIPE_USER_BASE) {
        if (ctx->user) { auto it = ctx->user->pipes.find(d.pipeline_id); if (it != ctx->user->pipes.end()) user_pipe = it->second; }
        if (!user_pipe) return fail(ctx, EUC_E_INVALID, "unknown run-time pipeline id %d", d.pipeline_id);
    } else
    if (d.pipeline_id < 0 || d.pipeline_id >= EUC_PIPE_COUNT) return fail(ctx, EUC_E_INVALID, "unknown pipeline_id %d", d.pipeline_id);
    if (d.primitive_kind < 0 || d.primitive_kind > EUC_PRIM_LINE_TRIANGLE_LIST) return fail(ctx, EUC_E_INVALID, "unknown primitive kind %d", d.primitive_kind);
    const bool lines = d.primitive_kind != EUC_PRIM_TRIANGLE_LIST;  // every pipeline renders lines too (primitives.rs:49-104 is generic over the pipeline)
    if (d.cull_mode < 0 || d.cull_mode > 2 || d.depth_test < 0 || d.depth_test > 3) return fail(ctx, EUC_E_INVALID, "bad cull/depth mode");

    const bool shadow = d.pipeline_id == EUC_PIPE_TEAPOT_SHADOW;
    const bool pixel_write = d.pixel_write != 0;
    const bool uses_depth = d.depth_test != EUC_DEPTH_NONE || d.depth_write != 0;
    const Buf* pb = nullptr;
    const Buf* db = nullptr;
    if (rc.pixel) {
        auto it = ctx->bufs.find(rc.pixel);
        if (it == ctx->bufs.end()) return fail(ctx, EUC_E_INVALID, "unknown pixel buffer handle");
        pb = &it->second;
    }
    if (rc.depth) {
        auto it = ctx->bufs.find(rc.depth);
        if (it == ctx->bufs.end()) return fail(ctx, EUC_E_INVALID, "unknown depth buffer handle");
        db = &it->second;
    }
    // euc_render_clear: whatever part of the requested clear the kernels of this render do not perform themselves is a
    // plain fill of the rendered rows, issued before them (also when the render turns out to draw nothing)
    euc_ctx::ClearReq clr = ctx->next_clear;
    ctx->next_clear = euc_ctx::ClearReq{};
    auto plain_clear = [&](bool px, bool z) -> int {
        int c = EUC_OK;
        if (px && clr.px && pb) { clr.px = false; if ((c = clear_rows_impl(ctx, *pb, clr.px_value, rc.row_begin, rc.row_end)) != EUC_OK) return c; }
        if (z && clr.z && db) { clr.z = false; if ((c = clear_rows_impl(ctx, *db, clr.z_value, rc.row_begin, rc.row_end)) != EUC_OK) return c; }
        return c;
    };
    // pipeline.rs:256-270
    if (!pixel_write && !uses_depth) return plain_clear(true, true);
    uint32_t w = 0, h = 0, layers = 1;
    auto size_of = [](const Buf* b, uint32_t& bw, uint32_t& bh, uint32_t& bl) { bw = b ? b->w : 0; bh = b ? b->h : 0; bl = b ? b->layers : 1; };
    if (pixel_write && uses_depth) {
        uint32_t pw, ph, pl, dw, dh, dl;
        size_of(pb, pw, ph, pl);
        size_of(db, dw, dh, dl);
        if (pw != dw || ph != dh || (pb && db && pl != dl))
            return fail(ctx, EUC_E_SIZE_MISMATCH, "Pixel target size is compatible with depth target size: [%u, %u] vs [%u, %u]", pw, ph, dw, dh);
        w = pw; h = ph; layers = pb ? pl : dl;
    } else if (pixel_write) {
        size_of(pb, w, h, layers);
    } else {
        size_of(db, w, h, layers);
    }
    if (w == 0 || h == 0) return plain_clear(true, true);  // Empty target: size [0,0] -> needed_threads == 0 -> nothing happens
    const uint32_t msaa = (uint32_t)std::min(std::max(d.msaa_level, 0), 6);  // pipeline.rs:291-294
    // pipeline.rs:329-330.  width > 20000*2^msaa makes group_rows 0 and the reference divides by zero.
    const uint64_t group_rows64 = 20000ull * (1ull << msaa) / std::max<uint64_t>(w, 1);
    if (group_rows64 == 0) return fail(ctx, EUC_E_UNSUPPORTED, "target width %u > 20000*2^msaa: the reference panics (division by zero, pipeline.rs:330)", w);
    if (w > 65535u || h > 65535u) return fail(ctx, EUC_E_UNSUPPORTED, "target larger than 65535 in a dimension");
    const uint32_t group_rows = (uint32_t)std::min<uint64_t>(group_rows64, 0x7fffffffull);
    if (h / group_rows == 0) return plain_clear(true, true);  // needed_threads == 0: the reference renders nothing (pipeline.rs:330,337)

    uint32_t row_begin = rc.row_begin, row_end = std::min(rc.row_end, h);
    if (row_begin >= row_end) return plain_clear(true, true);  // an empty band (a rank beyond the last tile row) renders nothing
    if (row_begin % TILE) return fail(ctx, EUC_E_INVALID, "row_begin must be a multiple of %d", TILE);

    // draws
    std::vector<DrawDev> dd(rc.n_draws);
    uint64_t tri_total = 0;
    const uint32_t stream_len = rc.geom->idx ? rc.geom->n_idx : rc.geom->n_verts;
    const Geom& gm = *rc.geom;
    for (uint32_t i = 0; i < rc.n_draws; ++i) {
        const euc_batch_draw& b = rc.draws[i];
        if ((uint64_t)b.first + b.count > stream_len) return fail(ctx, EUC_E_OUT_OF_BOUNDS, "draw %u reads past the end of the vertex stream", i);
        // can this draw read a vertex out of range?  (index.rs:53 panics; here: EUC_E_OUT_OF_BOUNDS from a checked render)
        if (b.count) {
            const int64_t lo = gm.idx ? (int64_t)gm.idx_min : (int64_t)b.first, hi = gm.idx ? (int64_t)gm.idx_max : (int64_t)b.first + b.count - 1;
            const bool known = gm.idx ? gm.bounds_known : true;
            if (known && (lo + b.base_vertex < 0 || hi + b.base_vertex >= (int64_t)gm.n_verts)) rc.maybe_oob_sync = true;
        }
        if (b.layer >= layers) return fail(ctx, EUC_E_INVALID, "draw %u targets layer %u of %u", i, b.layer, layers);
        // primitives per draw; a trailing partial primitive is dropped (pipeline.rs:283).  LineTriangleList turns every
        // collected triangle into three lines (primitives.rs:56-76).
        const uint32_t nprim = d.primitive_kind == EUC_PRIM_TRIANGLE_LIST ? b.count / 3 : (d.primitive_kind == EUC_PRIM_LINE_LIST ? b.count / 2 : (b.count / 3) * 3);
        dd[i] = DrawDev{b.first, b.count, b.base_vertex, b.layer, (uint32_t)tri_total, nprim};
        tri_total += nprim;
    }
    uint64_t draws_hash = 0;
    if (rc.maybe_oob_sync) {
        draws_hash = fnv1a(rc.draws, (size_t)rc.n_draws * sizeof(euc_batch_draw));
        if (gm.verified_hash == draws_hash && gm.verified_version == gm.version) rc.maybe_oob_sync = false;  // this table was checked on these indices
    }
    if (tri_total == 0) return plain_clear(true, true);
    if (tri_total > 0x7fffffffull) return fail(ctx, EUC_E_UNSUPPORTED, "too many primitives");

    Params prm{};
    prm.w = w; prm.h = h; prm.layers = layers;
    prm.tiles_x = (w + TILE - 1) / TILE; prm.tiles_y = (h + TILE - 1) / TILE;
    prm.row_begin = row_begin; prm.row_end = row_end;
    prm.group_rows = group_rows; prm.msaa_level = msaa;
    prm.pixel = pb ? (uint32_t*)pb->d : nullptr;
    prm.depth = db ? (float*)db->d : nullptr;
    prm.depth_test = d.depth_test; prm.depth_write = d.depth_write != 0; prm.pixel_write = pixel_write && !shadow; prm.uses_depth = uses_depth;
    if (prm.pixel_write && !prm.pixel) return plain_clear(true, true);
    if (uses_depth && !prm.depth) {
        // Empty depth target reads 0.0 and drops writes (texture.rs:312-317); size [0,0] only reaches here when
        // pixel_write is false, which returned above.  With both targets, sizes would have mismatched.
        return plain_clear(true, true);
    }
    {   // fused clear: colour when this render writes pixels, depth when it uses the depth target; the rest is filled now
        const bool fuse_px = clr.px && prm.pixel_write && prm.pixel, fuse_z = clr.z && uses_depth && prm.depth;
        const euc_ctx::ClearReq req = clr;
        const int c = plain_clear(!fuse_px, !fuse_z);
        if (c != EUC_OK) return c;
        prm.clear_mask = (fuse_px ? 1u : 0u) | (fuse_z ? 2u : 0u);
        prm.clear_px = req.px_value;
        std::memcpy(&prm.clear_z, &req.z_value, 4);
    }
    prm.zclip = d.z_clip_enabled != 0; prm.zmin = d.z_clip_min; prm.zmax = d.z_clip_max;
    prm.cull = d.cull_mode; prm.flip_y = d.y_axis_up ? -1.0f : 1.0f;
    prm.prim_kind = d.primitive_kind;
    prm.cta_bin = tri_total <= 65536 ? 1u : 0u;
    // row-restricted renders (multi-GPU bands) that keep less than half of the rows: most records are dropped, so the live
    // ones are stored by their own lanes instead of one bulk store per warp that would write all 32 slots
    prm.sparse_recs = ctx->sparse_recs >= 0 ? (uint32_t)ctx->sparse_recs : ((uint64_t)(prm.row_end - prm.row_begin) * 2 < prm.h ? 1u : 0u);
    prm.vertices = rc.geom->verts; prm.vstride = rc.geom->stride; prm.n_vertices = rc.geom->n_verts;
    prm.indices = rc.geom->idx;
    prm.n_draws = rc.n_draws; prm.n_tris = (uint32_t)tri_total;
    prm.stats = ctx->stats ? 1 : 0;
    prm.counters = ctx->counters;

    if (rc.n_mirrors > EUC_MAX_MIRRORS) return fail(ctx, EUC_E_INVALID, "at most %d mirrors", EUC_MAX_MIRRORS);
    for (uint32_t i = 0; i < rc.n_mirrors; ++i) {
        auto it = ctx->bufs.find(rc.mirrors[i]);
        if (it == ctx->bufs.end()) return fail(ctx, EUC_E_INVALID, "mirror %u: unknown buffer handle", i);
        if (!pb || it->second.w != pb->w || it->second.h != pb->h || it->second.layers != pb->layers)
            return fail(ctx, EUC_E_SIZE_MISMATCH, "mirror %u does not have the pixel target's size", i);
        prm.mirrors[i] = (uint32_t*)it->second.d;
    }
    prm.n_mirrors = prm.pixel_write ? rc.n_mirrors : 0;
    for (int i = 0; i < EUC_MAX_SAMPLERS; ++i) {
        const euc_sampler_desc& s = d.samplers[i];
        prm.samp[i] = SamplerDev{nullptr, 0, 0, s.format, s.filter, s.wrap};
        if (s.buf) {
            auto it = ctx->bufs.find(s.buf);
            if (it == ctx->bufs.end()) return fail(ctx, EUC_E_INVALID, "sampler %d: unknown buffer handle", i);
            if (it->second.w == 0 || it->second.h == 0) return fail(ctx, EUC_E_INVALID, "sampler %d: empty texture (texture.rs:63-66)", i);
            if (s.filter < 0 || s.filter > 1 || s.wrap < 0 || s.wrap > 3 || s.format < 0 || s.format > 1) return fail(ctx, EUC_E_INVALID, "sampler %d: bad filter/wrap/format", i);
            prm.samp[i].data = it->second.d;
            prm.samp[i].w = it->second.w;
            prm.samp[i].h = it->second.h;
        }
    }

    int rcode;
    prm.draw0 = dd[0];
    prm.draws = nullptr;
    const size_t draw_bytes = dd.size() * sizeof(DrawDev), draw_pad = (draw_bytes + 255) / 256 * 256;
    const bool batch_uniforms = rc.batch && d.uniform_bytes != 0 && rc.uniforms;
    const size_t ub = batch_uniforms ? (size_t)d.uniform_bytes * rc.n_draws : 0;
    if (batch_uniforms && (d.uniform_bytes & 15u)) return fail(ctx, EUC_E_INVALID, "batch uniform blocks must be a multiple of 16 bytes");
    if (rc.n_draws > 1 || batch_uniforms) {
        // per-draw tables go through pinned staging: asynchronous copies, one for the draws, one for the uniform blocks
        if ((rcode = ensure(ctx, ctx->draws, draw_bytes)) != EUC_OK) return rcode;
        if (ub && (rcode = ensure(ctx, ctx->uniforms, ub)) != EUC_OK) return rcode;
        uint8_t* stg = nullptr;
        const int slot = ctx->stage_next;
        if ((rcode = stage_tables(ctx, dd.data(), draw_bytes, batch_uniforms ? rc.uniforms : nullptr, ub, draw_pad, &stg)) != EUC_OK) return rcode;
        CU(cudaMemcpyAsync(ctx->draws.p, stg, draw_bytes, cudaMemcpyHostToDevice, ctx->stream));
        if (ub) CU(cudaMemcpyAsync(ctx->uniforms.p, stg + draw_pad, ub, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaEventRecord(ctx->stage_ev[slot], ctx->stream));
        ctx->stage_used[slot] = true;
        prm.draws = (const DrawDev*)ctx->draws.p;
    }
    if (rc.batch) {
        if (!batch_uniforms) {
            prm.uniforms = nullptr;
        } else {
            prm.uniforms = (const uint8_t*)ctx->uniforms.p;
            prm.uniform_stride = d.uniform_bytes;
        }
    } else {
        if (d.uniform_bytes > sizeof(prm.uni_inline)) return fail(ctx, EUC_E_INVALID, "uniform block larger than %zu bytes", sizeof(prm.uni_inline));
        if (d.uniform_bytes && !rc.uniforms) return fail(ctx, EUC_E_INVALID, "uniform_bytes > 0 but uniforms is NULL");
        if (d.uniform_bytes) std::memcpy(prm.uni_inline, rc.uniforms, d.uniform_bytes);
        prm.uniforms = nullptr;
    }
    // the staged copies above read pageable host memory synchronously w.r.t. the caller; dd lives until return.
    const uint32_t n_tiles = prm.tiles_x * prm.tiles_y * layers;
    if ((rcode = ensure(ctx, ctx->bbox, (size_t)prm.n_tris * sizeof(uint2))) != EUC_OK) return rcode;
    if ((rcode = ensure(ctx, ctx->tile_count, (size_t)n_tiles * 4, true)) != EUC_OK) return rcode;
    if ((rcode = ensure(ctx, ctx->tile_range, (size_t)n_tiles * sizeof(uint2))) != EUC_OK) return rcode;
    prm.tri_bbox = (uint2*)ctx->bbox.p;
    prm.tile_count = (uint32_t*)ctx->tile_count.p;
    prm.tile_range = (uint2*)ctx->tile_range.p;

    if (user_pipe) return render_driver(ctx, rc, prm, n_tiles, lines ? user_pipe->ops_lines : user_pipe->ops);
    const uint64_t waits_before = ctx->blocking_waits;
#define EUC_DISPATCH(P) rcode = lines ? render_typed<P, true>(ctx, rc, prm, n_tiles) : render_typed<P>(ctx, rc, prm, n_tiles); break
    switch (d.pipeline_id) {
        case EUC_PIPE_TEAPOT_SHADOW: EUC_DISPATCH(PipeTeapotShadow);
        case EUC_PIPE_TEAPOT_PHONG: EUC_DISPATCH(PipeTeapotPhong);
        case EUC_PIPE_TEX_CUBE: EUC_DISPATCH(PipeTexCube);
        case EUC_PIPE_BLEND_TRIS: EUC_DISPATCH(PipeBlendTris);
        case EUC_PIPE_VOXEL_ICON: EUC_DISPATCH(PipeVoxelIcon);
        case EUC_PIPE_VERTEX_COLOR: EUC_DISPATCH(PipeVertexColor);
        case EUC_PIPE_WIREFRAME: EUC_DISPATCH(PipeWireframe);
        default: rcode = EUC_E_INVALID;
    }
#undef EUC_DISPATCH
    // a checked render that found every index in range: its draw table need not be checked again on these indices
    if (rcode == EUC_OK && rc.maybe_oob_sync && ctx->blocking_waits > waits_before) { gm.verified_hash = draws_hash; gm.verified_version = gm.version; }
    // dd is pageable: make sure the async copy consumed it (render_typed synchronises; early outs do not)
    return rcode;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
extern "C" {

int euc_abi_version(void) { return EUC_B200_ABI_VERSION; }

int euc_init(int device_ordinal, euc_ctx** out_ctx) {
    if (!out_ctx) return EUC_E_INVALID;
    *out_ctx = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return EUC_E_CUDA;  // no CPU fallback
    if (device_ordinal < 0 || device_ordinal >= n) return EUC_E_INVALID;
    if (cudaSetDevice(device_ordinal) != cudaSuccess) return EUC_E_CUDA;
    euc_ctx* ctx = new euc_ctx();
    ctx->dev = device_ordinal;
    if (cudaStreamCreateWithFlags(&ctx->own, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return EUC_E_CUDA; }
    ctx->stream = ctx->own;
    if (cudaMalloc(&ctx->counters, CTR_WORDS * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(ctx->counters, 0, CTR_WORDS * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMallocHost(&ctx->counters_host, 8 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaHostAlloc((void**)&ctx->summary, 8 * sizeof(unsigned long long), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&ctx->summary_dev, (void*)ctx->summary, 0) != cudaSuccess) {
        delete ctx;
        return EUC_E_CUDA;
    }
    for (int i = 0; i < 8; ++i) ctx->summary[i] = 0;
    if (const char* e = getenv("EUC_ASYNC")) ctx->async = atoi(e) != 0;
    if (cudaEventCreateWithFlags(&ctx->ev_counts, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_setup, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return EUC_E_CUDA; }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device_ordinal);
    if (const char* e = getenv("EUC_SPARSE_RECS")) ctx->sparse_recs = atoi(e);
    if (const char* e = getenv("EUC_GROUP_CLS_MIN_WORLD")) ctx->group_cls_min_world = (uint32_t)std::max(atoi(e), 0);
    if (const char* e = getenv("EUC_OVF_ENTRIES")) ctx->ovf_fixed = (uint64_t)std::max(atoll(e), 0ll);
    *out_ctx = ctx;
    return EUC_OK;
}

int euc_group_destroy(euc_ctx* ctx);

int euc_shutdown(euc_ctx* ctx) {
    if (!ctx) return EUC_E_INVALID;
    cudaSetDevice(ctx->dev);
    if (ctx->group) euc_group_destroy(ctx);
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->bufs) { if (kv.second.ipc) cudaIpcCloseMemHandle(kv.second.d); else if (kv.second.owned) cudaFree(kv.second.d); }
    drain_profile(ctx);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    for (auto& kv : ctx->geoms) if (kv.second.owned) { cudaFree(kv.second.verts); cudaFree(kv.second.idx); }
    Scratch* ss[] = {&ctx->recs, &ctx->bbox, &ctx->tile_count, &ctx->tile_range, &ctx->tile_list, &ctx->draws, &ctx->uniforms, &ctx->tmp_verts, &ctx->tmp_idx, &ctx->winner,
                     &ctx->ovf, &ctx->ext};
    for (Scratch* s : ss) cudaFree(s->p);
    for (auto& kv : ctx->tickets) cudaEventDestroy(kv.second);
    for (auto& kv : ctx->builtin) delete kv.second;
    if (ctx->user) {
        for (auto& kv : ctx->user->pipes) { if (rt_api().ok) rt_api().ModuleUnload(kv.second->mod); delete kv.second; }
        delete ctx->user;
    }
    cudaEventDestroy(ctx->ev_counts);
    cudaEventDestroy(ctx->ev_setup);
    cudaStreamDestroy(ctx->aux);
    cudaFree(ctx->counters);
    cudaFreeHost(ctx->counters_host);
    if (ctx->summary) cudaFreeHost((void*)ctx->summary);
    for (int k = 0; k < 2; ++k) { if (ctx->stage_host[k]) cudaFreeHost(ctx->stage_host[k]); if (ctx->stage_ev[k]) cudaEventDestroy(ctx->stage_ev[k]); }
    cudaStreamDestroy(ctx->own);
    delete ctx;
    return EUC_OK;
}

const char* euc_last_error(euc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int euc_set_stream(euc_ctx* ctx, void* cuda_stream) {
    if (!ctx) return EUC_E_INVALID;
    DeviceGuard dg(ctx);
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own;
    return EUC_OK;
}

int euc_sync(euc_ctx* ctx) {
    if (!ctx) return EUC_E_INVALID;
    DeviceGuard dg(ctx);
    CU(cudaStreamSynchronize(ctx->stream));
    return take_deferred(ctx);  // an error a finished asynchronous render left behind
}

int euc_set_async(euc_ctx* ctx, int enabled) {
    if (!ctx) return EUC_E_INVALID;
    ctx->async = enabled != 0;
    return EUC_OK;
}

uint64_t euc_blocking_waits(euc_ctx* ctx) { return ctx ? ctx->blocking_waits : 0; }

int euc_set_stats(euc_ctx* ctx, int enabled) {
    if (!ctx) return EUC_E_INVALID;
    ctx->stats = enabled != 0;
    return EUC_OK;
}

int euc_get_stats(euc_ctx* ctx, euc_render_stats* out) {
    if (!ctx || !out) return EUC_E_INVALID;
    DeviceGuard dg(ctx);
    if (ctx->stats_on_device) {
        CU(cudaMemcpyAsync(ctx->counters_host, ctx->counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->last.binned_pairs = ctx->counters_host[0];
        ctx->last.fragments = ctx->counters_host[1];
        const int dc = take_deferred(ctx);
        if (dc != EUC_OK) return dc;
    }
    *out = ctx->last;
    return EUC_OK;
}

int euc_buf_create(euc_ctx* ctx, uint32_t width, uint32_t height, uint32_t layers, uint32_t texel_bytes, euc_buf* out) {
    if (!ctx || !out) return EUC_E_INVALID;
    if (texel_bytes != 4) return fail(ctx, EUC_E_UNSUPPORTED, "only 4-byte texels (u32 colour, f32 depth, RGBA8 texture) are supported");
    if (layers == 0) return fail(ctx, EUC_E_INVALID, "layers must be >= 1");
    DeviceGuard dg(ctx);
    Buf b;
    b.w = width; b.h = height; b.layers = layers; b.texel = texel_bytes;
    b.bytes = (size_t)width * height * layers * texel_bytes;  // Buffer::fill_with: len = product of sizes (buffer.rs:74-75)
    if (b.bytes) CU(cudaMalloc(&b.d, b.bytes));
    uint64_t hnd = ctx->next_handle++;
    ctx->bufs[hnd] = b;
    *out = hnd;
    return EUC_OK;
}

int euc_buf_wrap(euc_ctx* ctx, void* device_ptr, uint32_t width, uint32_t height, uint32_t layers, uint32_t texel_bytes, euc_buf* out) {
    if (!ctx || !out || !device_ptr) return EUC_E_INVALID;
    if (texel_bytes != 4 || layers == 0) return fail(ctx, EUC_E_UNSUPPORTED, "only 4-byte texels, layers >= 1");
    if (((uintptr_t)device_ptr & 15u) != 0) return fail(ctx, EUC_E_INVALID, "wrapped device memory must be 16-byte aligned");
    Buf b;
    b.d = device_ptr; b.w = width; b.h = height; b.layers = layers; b.texel = texel_bytes; b.owned = false;
    b.bytes = (size_t)width * height * layers * texel_bytes;
    uint64_t hnd = ctx->next_handle++;
    ctx->bufs[hnd] = b;
    *out = hnd;
    return EUC_OK;
}

int euc_set_profiling(euc_ctx* ctx, int enabled) {
    if (!ctx) return EUC_E_INVALID;
    ctx->profiling = enabled != 0;
    return EUC_OK;
}

int euc_get_profile(euc_ctx* ctx, float* ms, uint64_t* calls, int reset) {
    if (!ctx) return EUC_E_INVALID;
    DeviceGuard dg(ctx);
    CU(cudaStreamSynchronize(ctx->stream));
    drain_profile(ctx);
    for (int s = 0; s < EUC_STAGE_COUNT; ++s) {
        if (ms) ms[s] = ctx->prof_ms[s];
        if (calls) calls[s] = ctx->prof_calls[s];
        if (reset) { ctx->prof_ms[s] = 0.f; ctx->prof_calls[s] = 0; }
    }
    return EUC_OK;
}

uint64_t euc_launch_count(euc_ctx* ctx) { return ctx ? ctx->launches : 0; }

int euc_buf_destroy(euc_ctx* ctx, euc_buf buf) {
    if (!ctx) return EUC_E_INVALID;
    auto it = ctx->bufs.find(buf);
    if (it == ctx->bufs.end()) return fail(ctx, EUC_E_INVALID, "unknown buffer handle");
    DeviceGuard dg(ctx);
    CU(cudaStreamSynchronize(ctx->stream));
    if (it->second.ipc) CU(cudaIpcCloseMemHandle(it->second.d));
    else if (it->second.d && it->second.owned) CU(cudaFree(it->second.d));
    ctx->bufs.erase(it);
    return EUC_OK;
}

int euc_buf_clear(euc_ctx* ctx, euc_buf buf, const void* texel) {
    if (!ctx || !texel) return EUC_E_INVALID;
    auto it = ctx->bufs.find(buf);
    if (it == ctx->bufs.end()) return fail(ctx, EUC_E_INVALID, "unknown buffer handle");
    const Buf& b = it->second;
    uint32_t v;
    std::memcpy(&v, texel, 4);
    DeviceGuard dg(ctx);
    return clear_rows_impl(ctx, b, v, 0, b.h);
}

int euc_buf_clear_rows(euc_ctx* ctx, euc_buf buf, const void* texel, uint32_t row_begin, uint32_t row_end) {
    if (!ctx || !texel) return EUC_E_INVALID;
    auto it = ctx->bufs.find(buf);
    if (it == ctx->bufs.end()) return fail(ctx, EUC_E_INVALID, "unknown buffer handle");
    uint32_t v;
    std::memcpy(&v, texel, 4);
    DeviceGuard dg(ctx);
    return clear_rows_impl(ctx, it->second, v, row_begin, row_end);
}

int euc_render_clear(euc_ctx* ctx, const void* pixel_texel, const void* depth_texel) {
    if (!ctx) return EUC_E_INVALID;
    ctx->next_clear = euc_ctx::ClearReq{};
    if (pixel_texel) { ctx->next_clear.px = true; std::memcpy(&ctx->next_clear.px_value, pixel_texel, 4); }
    if (depth_texel) { ctx->next_clear.z = true; std::memcpy(&ctx->next_clear.z_value, depth_texel, 4); }
    return EUC_OK;
}

int euc_buf_upload(euc_ctx* ctx, euc_buf buf, const void* host, size_t bytes) {
    if (!ctx || !host) return EUC_E_INVALID;
    auto it = ctx->bufs.find(buf);
    if (it == ctx->bufs.end()) return fail(ctx, EUC_E_INVALID, "unknown buffer handle");
    if (bytes != it->second.bytes) return fail(ctx, EUC_E_SIZE_MISMATCH, "upload of %zu bytes into a buffer of %zu bytes", bytes, it->second.bytes);
    DeviceGuard dg(ctx);
    if (bytes) CU(cudaMemcpyAsync(it->second.d, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return EUC_OK;
}

int euc_buf_download(euc_ctx* ctx, euc_buf buf, void* host, size_t bytes) {
    if (!ctx || !host) return EUC_E_INVALID;
    auto it = ctx->bufs.find(buf);
    if (it == ctx->bufs.end()) return fail(ctx, EUC_E_INVALID, "unknown buffer handle");
    if (bytes != it->second.bytes) return fail(ctx, EUC_E_SIZE_MISMATCH, "download of %zu bytes from a buffer of %zu bytes", bytes, it->second.bytes);
    DeviceGuard dg(ctx);
    if (bytes) CU(cudaMemcpyAsync(host, it->second.d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return take_deferred(ctx);  // the bytes are there; an error means an earlier asynchronous render did not draw
}

int euc_host_alloc(euc_ctx* ctx, size_t bytes, void** out_ptr) {
    if (!ctx || !out_ptr) return EUC_E_INVALID;
    DeviceGuard dg(ctx);
    CU(cudaMallocHost(out_ptr, std::max<size_t>(bytes, 1)));
    return EUC_OK;
}

int euc_host_free(euc_ctx* ctx, void* ptr) {
    if (!ctx) return EUC_E_INVALID;
    DeviceGuard dg(ctx);
    if (ptr) CU(cudaFreeHost(ptr));
    return EUC_OK;
}

int euc_buf_download_async(euc_ctx* ctx, euc_buf buf, void* host, size_t bytes, uint64_t* out_ticket) {
    if (!ctx || !host || !out_ticket) return EUC_E_INVALID;
    auto it = ctx->bufs.find(buf);
    if (it == ctx->bufs.end()) return fail(ctx, EUC_E_INVALID, "unknown buffer handle");
    if (bytes != it->second.bytes) return fail(ctx, EUC_E_SIZE_MISMATCH, "download of %zu bytes from a buffer of %zu bytes", bytes, it->second.bytes);
    DeviceGuard dg(ctx);
    if (bytes) CU(cudaMemcpyAsync(host, it->second.d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    cudaEvent_t ev = nullptr;
    CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU(cudaEventRecord(ev, ctx->stream));
    const uint64_t t = ctx->next_ticket++;
    ctx->tickets[t] = ev;
    *out_ticket = t;
    return EUC_OK;
}

int euc_ticket_wait(euc_ctx* ctx, uint64_t ticket) {
    if (!ctx) return EUC_E_INVALID;
    auto it = ctx->tickets.find(ticket);
    if (it == ctx->tickets.end()) return fail(ctx, EUC_E_INVALID, "unknown or already consumed ticket");
    cudaEvent_t ev = it->second;
    ctx->tickets.erase(it);
    DeviceGuard dg(ctx);
    cudaError_t e = cudaEventSynchronize(ev);
    cudaEventDestroy(ev);
    if (e != cudaSuccess) return fail(ctx, EUC_E_CUDA, "cudaEventSynchronize: %s", cudaGetErrorString(e));
    return take_deferred(ctx);
}

int euc_buf_device_ptr(euc_ctx* ctx, euc_buf buf, void** out_ptr, size_t* out_bytes) {
    if (!ctx || !out_ptr) return EUC_E_INVALID;
    auto it = ctx->bufs.find(buf);
    if (it == ctx->bufs.end()) return fail(ctx, EUC_E_INVALID, "unknown buffer handle");
    *out_ptr = it->second.d;
    if (out_bytes) *out_bytes = it->second.bytes;
    return EUC_OK;
}

int euc_buf_size(euc_ctx* ctx, euc_buf buf, uint32_t* w, uint32_t* h, uint32_t* layers) {
    if (!ctx) return EUC_E_INVALID;
    auto it = ctx->bufs.find(buf);
    if (it == ctx->bufs.end()) return fail(ctx, EUC_E_INVALID, "unknown buffer handle");
    if (w) *w = it->second.w;
    if (h) *h = it->second.h;
    if (layers) *layers = it->second.layers;
    return EUC_OK;
}

int euc_geom_create(euc_ctx* ctx, const void* vertices, uint32_t vertex_stride, uint32_t n_vertices, const uint32_t* indices,
                    uint32_t n_indices, euc_geom* out) {
    if (!ctx || !out || (!vertices && n_vertices) || vertex_stride == 0) return EUC_E_INVALID;
    DeviceGuard dg(ctx);
    Geom g;
    g.stride = vertex_stride; g.n_verts = n_vertices; g.n_idx = indices ? n_indices : 0;
    const size_t vb = (size_t)vertex_stride * n_vertices;
    CU(cudaMalloc((void**)&g.verts, std::max<size_t>(vb, 16)));
    if (vb) CU(cudaMemcpyAsync(g.verts, vertices, vb, cudaMemcpyHostToDevice, ctx->stream));
    if (indices) {
        CU(cudaMalloc((void**)&g.idx, std::max<size_t>((size_t)n_indices * 4, 16)));
        if (n_indices) CU(cudaMemcpyAsync(g.idx, indices, (size_t)n_indices * 4, cudaMemcpyHostToDevice, ctx->stream));
        index_bounds(indices, n_indices, g.idx_min, g.idx_max);  // once per geometry: its renders then need no host check
        g.bounds_known = true;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    uint64_t hnd = ctx->next_handle++;
    ctx->geoms[hnd] = g;
    *out = hnd;
    return EUC_OK;
}

int euc_geom_wrap(euc_ctx* ctx, void* device_vertices, uint32_t vertex_stride, uint32_t n_vertices, void* device_indices, uint32_t n_indices,
                  euc_geom* out) {
    if (!ctx || !out || !device_vertices || vertex_stride == 0) return EUC_E_INVALID;
    if (((uintptr_t)device_vertices & 15u) || ((uintptr_t)device_indices & 3u)) return fail(ctx, EUC_E_INVALID, "wrapped geometry must be 16-byte (vertices) / 4-byte (indices) aligned");
    Geom g;
    g.verts = (uint8_t*)device_vertices; g.stride = vertex_stride; g.n_verts = n_vertices;
    g.idx = (uint32_t*)device_indices; g.n_idx = device_indices ? n_indices : 0; g.owned = false;
    uint64_t hnd = ctx->next_handle++;
    ctx->geoms[hnd] = g;
    *out = hnd;
    return EUC_OK;
}

int euc_geom_update(euc_ctx* ctx, euc_geom geom, const void* vertices, const uint32_t* indices) {
    if (!ctx) return EUC_E_INVALID;
    auto it = ctx->geoms.find(geom);
    if (it == ctx->geoms.end()) return fail(ctx, EUC_E_INVALID, "unknown geometry handle");
    Geom& g = it->second;
    DeviceGuard dg(ctx);
    if (vertices && g.n_verts) CU(cudaMemcpyAsync(g.verts, vertices, (size_t)g.stride * g.n_verts, cudaMemcpyHostToDevice, ctx->stream));
    if (indices) {
        if (!g.idx) return fail(ctx, EUC_E_INVALID, "geometry has no index buffer");
        if (g.n_idx) CU(cudaMemcpyAsync(g.idx, indices, (size_t)g.n_idx * 4, cudaMemcpyHostToDevice, ctx->stream));
        // large index arrays are not re-scanned per update (a frame loop re-uploading its geometry must not pay a host pass
        // over it): the device checks every index and an out-of-range one is reported by a later call (euc_set_async)
        g.bounds_known = g.n_idx <= HOST_SCAN_MAX_INDICES;
        if (g.bounds_known) index_bounds(indices, g.n_idx, g.idx_min, g.idx_max);
        ++g.version;
    }
    return EUC_OK;
}

int euc_geom_update_range(euc_ctx* ctx, euc_geom geom, const void* vertices, uint32_t first_vertex, uint32_t n_vertices, const uint32_t* indices,
                          uint32_t first_index, uint32_t n_indices) {
    if (!ctx) return EUC_E_INVALID;
    auto it = ctx->geoms.find(geom);
    if (it == ctx->geoms.end()) return fail(ctx, EUC_E_INVALID, "unknown geometry handle");
    Geom& g = it->second;
    if ((uint64_t)first_vertex + n_vertices > g.n_verts || (uint64_t)first_index + n_indices > g.n_idx) return fail(ctx, EUC_E_OUT_OF_BOUNDS, "range past the end of the geometry");
    DeviceGuard dg(ctx);
    if (vertices && n_vertices) CU(cudaMemcpyAsync(g.verts + (size_t)first_vertex * g.stride, vertices, (size_t)n_vertices * g.stride, cudaMemcpyHostToDevice, ctx->stream));
    if (indices && n_indices) {
        CU(cudaMemcpyAsync(g.idx + first_index, indices, (size_t)n_indices * 4, cudaMemcpyHostToDevice, ctx->stream));
        g.bounds_known = false;  // a part of the indices changed: the device checks (deferred error, see euc_set_async)
        ++g.version;
    }
    return EUC_OK;
}

int euc_buf_download_rows_async(euc_ctx* ctx, euc_buf buf, void* host, uint32_t row_begin, uint32_t row_end, uint64_t* out_ticket) {
    if (!ctx || !host || !out_ticket) return EUC_E_INVALID;
    auto it = ctx->bufs.find(buf);
    if (it == ctx->bufs.end()) return fail(ctx, EUC_E_INVALID, "unknown buffer handle");
    const Buf& b = it->second;
    if (b.layers != 1) return fail(ctx, EUC_E_UNSUPPORTED, "row ranges are read back from single-layer buffers");
    row_end = std::min(row_end, b.h);
    DeviceGuard dg(ctx);
    if (row_begin < row_end && b.w) CU(cudaMemcpyAsync(host, (const uint8_t*)b.d + (size_t)row_begin * b.w * 4, (size_t)(row_end - row_begin) * b.w * 4, cudaMemcpyDeviceToHost, ctx->stream));
    cudaEvent_t ev = nullptr;
    CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU(cudaEventRecord(ev, ctx->stream));
    const uint64_t t = ctx->next_ticket++;
    ctx->tickets[t] = ev;
    *out_ticket = t;
    return EUC_OK;
}

int euc_geom_destroy(euc_ctx* ctx, euc_geom geom) {
    if (!ctx) return EUC_E_INVALID;
    auto it = ctx->geoms.find(geom);
    if (it == ctx->geoms.end()) return fail(ctx, EUC_E_INVALID, "unknown geometry handle");
    DeviceGuard dg(ctx);
    CU(cudaStreamSynchronize(ctx->stream));
    if (it->second.owned) {
        CU(cudaFree(it->second.verts));
        if (it->second.idx) CU(cudaFree(it->second.idx));
    }
    ctx->geoms.erase(it);
    return EUC_OK;
}

int euc_render_geom_rows(euc_ctx* ctx, const euc_pipeline_desc* desc, euc_geom geom, euc_buf pixel, euc_buf depth, uint32_t row_begin,
                         uint32_t row_end) {
    DropClear drop_clear{ctx};
    if (!ctx || !desc) return EUC_E_INVALID;
    auto it = ctx->geoms.find(geom);
    if (it == ctx->geoms.end()) return fail(ctx, EUC_E_INVALID, "unknown geometry handle");
    DeviceGuard dg(ctx);
    const Geom& g = it->second;
    euc_batch_draw one{0, g.idx ? g.n_idx : g.n_verts, 0, 0};
    RenderCall rc{desc, &g, &one, 1, desc->uniforms, false, pixel, depth, row_begin, row_end};
    return render_common(ctx, rc);
}

int euc_buf_ipc_export(euc_ctx* ctx, euc_buf buf, void* handle_out) {
    if (!ctx || !handle_out) return EUC_E_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == EUC_IPC_HANDLE_BYTES, "IPC handle size");
    auto it = ctx->bufs.find(buf);
    if (it == ctx->bufs.end()) return fail(ctx, EUC_E_INVALID, "unknown buffer handle");
    if (!it->second.owned || it->second.bytes < (2u << 20))
        return fail(ctx, EUC_E_UNSUPPORTED, "only buffers created by euc_buf_create with at least 2 MiB can be exported (allocation base == buffer base)");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, it->second.d));
    std::memcpy(handle_out, &h, sizeof h);
    return EUC_OK;
}

int euc_buf_ipc_import(euc_ctx* ctx, const void* handle, uint32_t width, uint32_t height, uint32_t layers, uint32_t texel_bytes, euc_buf* out) {
    if (!ctx || !handle || !out) return EUC_E_INVALID;
    if (texel_bytes != 4 || layers == 0) return fail(ctx, EUC_E_UNSUPPORTED, "only 4-byte texels, layers >= 1");
    DeviceGuard dg(ctx);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    void* ptr = nullptr;
    CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    Buf b;
    b.d = ptr; b.w = width; b.h = height; b.layers = layers; b.texel = texel_bytes; b.owned = false; b.ipc = true;
    b.bytes = (size_t)width * height * layers * texel_bytes;
    uint64_t hnd = ctx->next_handle++;
    ctx->bufs[hnd] = b;
    *out = hnd;
    return EUC_OK;
}

int euc_render_geom_rows_mirrored(euc_ctx* ctx, const euc_pipeline_desc* desc, euc_geom geom, euc_buf pixel, euc_buf depth, uint32_t row_begin,
                                  uint32_t row_end, const euc_buf* mirrors, uint32_t n_mirrors) {
    DropClear drop_clear{ctx};
    if (!ctx || !desc || (n_mirrors && !mirrors)) return EUC_E_INVALID;
    auto it = ctx->geoms.find(geom);
    if (it == ctx->geoms.end()) return fail(ctx, EUC_E_INVALID, "unknown geometry handle");
    DeviceGuard dg(ctx);
    const Geom& g = it->second;
    euc_batch_draw one{0, g.idx ? g.n_idx : g.n_verts, 0, 0};
    RenderCall rc{desc, &g, &one, 1, desc->uniforms, false, pixel, depth, row_begin, row_end, mirrors, n_mirrors};
    return render_common(ctx, rc);
}

int euc_pipeline_register(euc_ctx* ctx, const char* source, const char* struct_name, int32_t* out_pipeline_id) {
    if (!ctx || !source || !struct_name || !out_pipeline_id) return EUC_E_INVALID;
    DeviceGuard dg(ctx);
    if (!ctx->user) ctx->user = new euc_user_pipes();
    return register_pipeline(ctx, *ctx->user, source, struct_name, out_pipeline_id);
}

const char* euc_pipeline_log(euc_ctx* ctx) { return (ctx && ctx->user) ? ctx->user->log.c_str() : ""; }

int euc_render_geom(euc_ctx* ctx, const euc_pipeline_desc* desc, euc_geom geom, euc_buf pixel, euc_buf depth) {
    DropClear drop_clear{ctx};
    return euc_render_geom_rows(ctx, desc, geom, pixel, depth, 0, 0xffffffffu);
}

int euc_render(euc_ctx* ctx, const euc_pipeline_desc* desc, const void* vertices, uint32_t vertex_stride, uint32_t n_vertices,
               const uint32_t* indices, uint32_t n_indices, euc_buf pixel, euc_buf depth) {
    DropClear drop_clear{ctx};
    if (!ctx || !desc || (!vertices && n_vertices) || vertex_stride == 0) return EUC_E_INVALID;
    DeviceGuard dg(ctx);
    int rcode;
    const size_t vb = (size_t)vertex_stride * n_vertices;
    if ((rcode = ensure(ctx, ctx->tmp_verts, std::max<size_t>(vb, 16))) != EUC_OK) return rcode;
    if (vb) CU(cudaMemcpyAsync(ctx->tmp_verts.p, vertices, vb, cudaMemcpyHostToDevice, ctx->stream));
    Geom g;
    g.verts = (uint8_t*)ctx->tmp_verts.p; g.stride = vertex_stride; g.n_verts = n_vertices;
    if (indices) {
        if ((rcode = ensure(ctx, ctx->tmp_idx, std::max<size_t>((size_t)n_indices * 4, 16))) != EUC_OK) return rcode;
        if (n_indices) CU(cudaMemcpyAsync(ctx->tmp_idx.p, indices, (size_t)n_indices * 4, cudaMemcpyHostToDevice, ctx->stream));
        g.idx = (uint32_t*)ctx->tmp_idx.p; g.n_idx = n_indices;
        g.bounds_known = n_indices <= HOST_SCAN_MAX_INDICES;
        if (g.bounds_known) index_bounds(indices, n_indices, g.idx_min, g.idx_max);
    }
    euc_batch_draw one{0, g.idx ? g.n_idx : g.n_verts, 0, 0};
    RenderCall rc{desc, &g, &one, 1, desc->uniforms, false, pixel, depth, 0, 0xffffffffu};
    return render_common(ctx, rc);
}

int euc_render_batch(euc_ctx* ctx, const euc_pipeline_desc* desc, euc_geom geom, const euc_batch_draw* draws, uint32_t n_draws,
                     const void* uniforms, euc_buf pixel, euc_buf depth) {
    DropClear drop_clear{ctx};
    if (!ctx || !desc || (!draws && n_draws)) return EUC_E_INVALID;
    auto it = ctx->geoms.find(geom);
    if (it == ctx->geoms.end()) return fail(ctx, EUC_E_INVALID, "unknown geometry handle");
    DeviceGuard dg(ctx);
    RenderCall rc{desc, &it->second, draws, n_draws, uniforms, true, pixel, depth, 0, 0xffffffffu};
    return render_common(ctx, rc);
}

}  // extern "C"

#include "group.inc"
