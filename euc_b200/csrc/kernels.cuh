// kernels.cuh — sm_100a kernels of the raster back end.
//
//   K1+K2  setup_kernel<P>      vertex shade (per stream vertex, 128-bit loads) + primitive assembly + triangle
//                               setup (src/pipeline.rs:273-289, src/index.rs:52-54, triangles.rs:54-173) + per-tile
//                               population count.  One thread per triangle; emits a fixed-size setup record.
//   K3     alloc_tiles_kernel   gives every non-empty 16x16 tile a private slice of the pair list
//          fill_kernel          writes (tile <- triangle) pairs
//   K4/K5  raster_kernel<P>     one warp per tile; restores submission order inside the tile's list (pipeline.rs:581),
//                               then: lane = half a tile row (8 px); depth (and the winner id) of the tile live in
//                               registers, colour in shared memory, for the whole list; setup records are staged in
//                               shared memory with cp.async; coverage, depth test, fragment shade (incl. euc's
//                               coarse-shading "MSAA"), blend in submission order (triangles.rs:219-303,
//                               pipeline.rs:514-578).
//          resolve_kernel<P>    deferred pipelines: fragment + blend once per pixel for the winning primitive
//   K7     fill_u32_kernel      Target::clear (buffer.rs:213-218)
//
// Bit-exactness notes (DESIGN.md §"Exactness"): the per-pixel weights are euc's *sequentially accumulated* chain
// started at row_range[0] of the (triangle,row,band); lanes replay the chain up to their segment.
#pragma once
#include "shaders.cuh"

namespace eucb {

constexpr int TILE = 16;
constexpr int REC_BASE_WORDS = 24;

template <class P> struct RecLayout {
    static constexpr int VPAD = (3 * P::V + 3) / 4 * 4;
    static constexpr int WORDS = REC_BASE_WORDS + VPAD;
    static constexpr int BYTES = WORDS * 4;
};

// Record word offsets
enum : int {
    R_O = 0,     // w_hom_origin[3]
    R_DX = 3,    // w_hom_dx[3]
    R_DY = 6,    // w_hom_dy[3]
    R_ZH = 9,    // verts_hom[i][2]
    R_VY = 12,   // verts_by_y: a.x a.y b.x b.y c.x c.y
    R_BBX = 18,  // x0 | x1 << 16     (bounds clamped to [0, w])
    R_BBY = 19,  // y0 | y1 << 16     (bounds clamped to [0, h]; per-band clamp happens per row)
    R_FLAGS = 20,  // bit0: all three vertices pass the z clip
    R_DRAW = 21,
    R_TRI = 22,  // primitive index inside this render call
    R_VAR = 24
};

struct DrawDev {
    uint32_t first, count;
    int32_t base_vertex;
    uint32_t layer;
    uint32_t tri_begin;  // prefix of triangle counts
    uint32_t n_tris;
};

struct Params {
    // target
    uint32_t w, h, layers;
    uint32_t tiles_x, tiles_y;  // tiles per layer
    uint32_t row_begin, row_end;  // rendered rows (multi-GPU bands); row_begin % 16 == 0
    uint32_t group_rows;          // euc band height (pipeline.rs:329)
    uint32_t msaa_level;
    // fused clear (euc_render_clear): the render behaves as if its targets' rendered rows had been filled with these values
    // first; tiles start from the constants instead of loading, and every tile of the rendered rows is written back
    uint32_t clear_mask;  // bit 0: colour, bit 1: depth
    uint32_t clear_px;
    float clear_z;
    uint32_t* pixel;
    float* depth;
    uint32_t* mirrors[EUC_MAX_MIRRORS];  // peer framebuffers that receive this render's colour rows (fused gather)
    uint32_t n_mirrors;
    uint32_t* winner;  // deferred pipelines: per-pixel id of the last primitive whose fragment passed (NO_WINNER = none)
    // modes (pipeline.rs:178-209)
    int32_t depth_test, depth_write, pixel_write, uses_depth;
    int32_t zclip;
    float zmin, zmax;
    int32_t cull;
    float flip_y;
    // geometry
    const uint8_t* vertices;
    uint32_t vstride, n_vertices;
    const uint32_t* indices;
    const DrawDev* draws;   // n_draws > 1: device table
    DrawDev draw0;          // n_draws == 1: the draw itself (no table, no upload: a render is then pure kernel launches)
    uint32_t n_draws;
    uint32_t n_tris;
    const uint8_t* uniforms;  // device array of n_draws blocks (batch) or nullptr -> uni_inline
    uint32_t uniform_stride;
    // intermediates
    uint32_t* recs;
    uint2* tri_bbox;
    uint32_t* tile_count;   // zero on entry, zero again after fill
    uint2* tile_range;      // (offset, n) per tile
    uint32_t* tile_list;
    uint32_t list_capacity;
    int32_t prim_kind;      // euc_primitive_kind
    uint32_t cta_bin;       // 1: primitives covering > 256 tiles are binned by the whole CTA (few, huge primitives)
    uint32_t sparse_recs;   // 1: live records are stored by their own lanes (row-restricted renders drop most primitives)
    // Row-restricted renders of many primitives (a rank's band of a large frame): band_classify_kernel first lists the
    // primitives whose vertical bounds meet the band (survivors[0 .. counters[9])), and the set-up kernel walks that list.
    uint32_t* survivors;
    // The list may come in slices, one per source rank of a group (group_classify_kernel): slice s holds surv_counts[s]
    // ids from survivors + s * surv_stride on.  surv_counts == nullptr: one slice, its length in counters[9].
    const uint32_t* surv_counts;
    uint32_t surv_slices, surv_stride;
    // Classification across a group (sort-middle of primitive ids): this rank looks at primitives [cls_first, cls_first +
    // cls_n) of the frame and appends each to the list slice `cls_rank` of every rank whose band it meets (peer stores).
    uint32_t* const* cls_lists;   // every rank's `survivors` base as mapped here
    uint32_t* cls_counts;         // local fill of this rank's slice in each destination list
    uint32_t cls_rank, cls_world, cls_band_rows, cls_first, cls_n;
    uint32_t bin_cap;       // > 0: fixed-capacity bins (tile t owns list[t*bin_cap ..]); setup appends directly, no alloc/fill pass
    // Bin overflow handled on the device (renders that never wait for the host): a pair that does not fit its tile's bin
    // goes to `ovf` (tile, primitive); the raster warp of such a tile collects its pairs into a slice of `ext`.
    // ovf_cap == 0: an overflowing bin flags the render instead (bit 1) and the host redoes it on the exact path.
    uint2* ovf;
    uint32_t* ext;
    uint32_t ovf_cap;
    // Render summary for the host, written by the last raster warp into mapped pinned memory (no host call needed to
    // learn it): [0] sequence number (written last), [1] flags, [2] longest tile list, [3] overflow pairs, [4] pairs,
    // [5] tiles of the render
    volatile unsigned long long* summary;
    unsigned long long seq;
    unsigned long long* counters;  // [0] pairs, [1] fragments, [2] list cursor, [3] error flags, [4] tile ticket, [5] overflow
                                   // pairs, [6] ext cursor | warps done << 32, [7] longest tile list
    int32_t stats;
    SamplerDev samp[EUC_MAX_SAMPLERS];
    alignas(16) uint8_t uni_inline[320];
};

// -------------------------------------------------------------------------------------------------------
// K7 clear
// -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fill_u32_kernel(uint32_t* __restrict__ dst, size_t n, uint32_t v) {
    // scalar head up to the first 16-byte boundary (a row range of a target whose width is not a multiple of 4), 128-bit
    // body, scalar tail; head and tail are written by the first threads of block 0
    const size_t head = min(n, (size_t)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u) >> 2));
    uint32_t* const body = dst + head;
    const size_t nb = n - head;
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    const uint4 vv = make_uint4(v, v, v, v);
    for (; i + 4 <= nb; i += stride) *reinterpret_cast<uint4*>(body + i) = vv;
    if (blockIdx.x == 0) {
        if (threadIdx.x < head) dst[threadIdx.x] = v;
        if (threadIdx.x < (nb & 3)) body[(nb & ~(size_t)3) + threadIdx.x] = v;
    }
}

// -------------------------------------------------------------------------------------------------------
// tile traversal shared by the count and fill passes
// -------------------------------------------------------------------------------------------------------
struct TileRect { uint32_t tx0, ty0, ntx, nty, layer_base; };

// Error / abort flags of a render (counters[3])
constexpr unsigned long long FLAG_OOB = 1ull;       // a vertex index was out of range
constexpr unsigned long long FLAG_BINS = 2ull;      // a bin (or the pair list of the exact path) was too small: the host redoes the render
constexpr unsigned long long FLAG_OVF_FULL = 4ull;  // (informative) the overflow buffer was too small: some tiles were rendered by scanning all primitives
constexpr uint32_t TILE_LOST = 0x80000000u;         // tile_count bit: pairs of this tile were dropped, its list is incomplete

// Fast-path append of (tile <- primitive) when slot `sl` of the tile's bin has been taken.
__device__ __forceinline__ void bin_store(const Params& p, uint32_t tile, uint32_t sl, uint32_t tri, bool& over) {
    if (sl < p.bin_cap) {
        p.tile_list[(size_t)tile * p.bin_cap + sl] = tri;
    } else if (p.ovf_cap) {
        const unsigned long long k = atomicAdd(p.counters + 5, 1ull);
        if (k < (unsigned long long)p.ovf_cap) {
            p.ovf[k] = make_uint2(tile, tri);
        } else {
            // Not even the overflow buffer has room (a scene far denser than the ones this target has seen).  The pair is
            // dropped and the tile marked: its raster warp ignores the lists and tests every primitive of the render against
            // the tile instead (slow, but right, and nobody has to wait for the host).
            atomicOr(p.tile_count + tile, TILE_LOST);
            over = true;
        }
    } else {
        over = true;
    }
}

__device__ __forceinline__ bool tile_rect(const Params& p, uint2 bb, uint32_t layer, TileRect& r) {
    uint32_t x0 = bb.x & 0xffffu, x1 = bb.x >> 16, y0 = bb.y & 0xffffu, y1 = bb.y >> 16;
    y0 = max(y0, p.row_begin);
    y1 = min(y1, p.row_end);
    if (x1 <= x0 || y1 <= y0) return false;
    r.tx0 = x0 / TILE; r.ty0 = y0 / TILE;
    r.ntx = (x1 - 1) / TILE - r.tx0 + 1;
    r.nty = (y1 - 1) / TILE - r.ty0 + 1;
    r.layer_base = layer * p.tiles_x * p.tiles_y;
    return true;
}

// Calls op(tile_index, tri) for every tile of every lane's rectangle.  Small rectangles (<= 8 tiles) are walked by
// their own lane; larger ones by the whole warp, one primitive at a time; huge ones (> 256 tiles, e.g. a cube face at
// 1080p) by the whole CTA through a shared-memory list.  Must be reached by every thread of the CTA.
template <class OP> __device__ __forceinline__ void for_each_tile(const Params& p, bool valid, const TileRect& r, uint32_t tri, OP op) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t nt = valid ? r.ntx * r.nty : 0u;
    bool small = nt <= 8u;
    if (valid && small) {
        for (uint32_t j = 0; j < r.nty; ++j)
            for (uint32_t i = 0; i < r.ntx; ++i) op(r.layer_base + (r.ty0 + j) * p.tiles_x + r.tx0 + i, tri);
    }
    if (p.cta_bin) {  // uniform: the host enables it for renders with few primitives (barriers cost ~15 us at 2^20)
        constexpr uint32_t HUGE_SLOTS = 32;  // per CTA; further huge primitives take the warp path below
        __shared__ uint32_t huge_n;
        __shared__ uint4 huge_a[HUGE_SLOTS];  // tx0, ty0, ntx, nt
        __shared__ uint2 huge_b[HUGE_SLOTS];  // layer_base, tri
        bool huge = valid && nt > 256u;
        if (threadIdx.x == 0) huge_n = 0u;
        __syncthreads();
        if (huge) {
            const uint32_t k = atomicAdd(&huge_n, 1u);
            if (k < HUGE_SLOTS) {
                huge_a[k] = make_uint4(r.tx0, r.ty0, r.ntx, nt);
                huge_b[k] = make_uint2(r.layer_base, tri);
            } else {
                huge = false;
            }
        }
        __syncthreads();
        const uint32_t hn = min(huge_n, HUGE_SLOTS);
        for (uint32_t e = 0; e < hn; ++e) {
            const uint4 a = huge_a[e];
            const uint2 b = huge_b[e];
            for (uint32_t i = threadIdx.x; i < a.w; i += blockDim.x) {
                const uint32_t j = i / a.z, k = i - j * a.z;
                op(b.x + (a.y + j) * p.tiles_x + a.x + k, b.y);
            }
        }
        __syncthreads();
        if (huge) { valid = false; nt = 0u; }
    }
    uint32_t big = __ballot_sync(0xffffffffu, valid && !small);
    while (big) {
        int src = __ffs(big) - 1;
        big &= big - 1;
        uint32_t tx0 = __shfl_sync(0xffffffffu, r.tx0, src), ty0 = __shfl_sync(0xffffffffu, r.ty0, src);
        uint32_t ntx = __shfl_sync(0xffffffffu, r.ntx, src), n = __shfl_sync(0xffffffffu, nt, src);
        uint32_t lb = __shfl_sync(0xffffffffu, r.layer_base, src), t = __shfl_sync(0xffffffffu, tri, src);
        for (uint32_t i = lane; i < n; i += 32u) {
            uint32_t j = i / ntx, k = i - j * ntx;
            op(lb + (ty0 + j) * p.tiles_x + tx0 + k, t);
        }
    }
}

__device__ __forceinline__ uint32_t find_draw(const Params& p, uint32_t tri) {
    if (p.n_draws == 1) return 0;
    uint32_t lo = 0, hi = p.n_draws - 1;  // last draw with tri_begin <= tri
    while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (__ldg(&p.draws[mid].tri_begin) <= tri) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ DrawDev draw_of(const Params& p, uint32_t d) { return p.n_draws == 1 ? p.draw0 : p.draws[d]; }

template <class P> __device__ __forceinline__ const typename P::Uniforms& uniforms_of(const Params& p, uint32_t draw) {
    const uint8_t* base = p.uniforms ? p.uniforms + (size_t)draw * p.uniform_stride : p.uni_inline;
    return *reinterpret_cast<const typename P::Uniforms*>(base);
}

// Bulk (TMA, 1-D) store of a contiguous shared-memory block to global memory (SASS: UBLKCP.G.S)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"((uint32_t)__cvta_generic_to_shared(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// -------------------------------------------------------------------------------------------------------
// K1 + K2: vertex shade, assemble, set up.  triangles.rs:54-173.
// Records are written to a shared-memory stage (one slot per thread) and leave the SM as ONE bulk store per warp:
// 32 consecutive primitives' records are contiguous in global memory, so the store is fully coalesced, whereas
// per-thread 16-byte pieces of 144-byte records touch ~22 sectors per request.
// -------------------------------------------------------------------------------------------------------
// -------------------------------------------------------------------------------------------------------
// Fast-path binning of primitives that cover many tiles (fixed-capacity bins: slot = atomicAdd(count), list[slot] = id).
// An append is a round trip to L2 followed by a dependent store, so a walk costs the round trips it serialises, not
// the tiles it visits: both walks below issue eight appends per thread before they consume the first slot.
// -------------------------------------------------------------------------------------------------------
// Warp level: the (primitive, tile) pairs of the warp's primitives are flattened; item f belongs to the lane whose
// inclusive prefix is the first one above f, and the 32 lanes take 256 items per round whatever the primitives' sizes.
// Must be reached by all 32 lanes.
__device__ __forceinline__ void bin_big_warp(const Params& p, bool big, const TileRect& r, uint32_t nt, uint32_t tri, bool& over, uint32_t& npairs) {
    const uint32_t lane = threadIdx.x & 31u, full = 0xffffffffu;
    const uint32_t n = big ? nt : 0u;
    uint32_t incl = n;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { const uint32_t v = __shfl_up_sync(full, incl, s); if (lane >= (uint32_t)s) incl += v; }
    const uint32_t excl = incl - n, total = __shfl_sync(full, incl, 31);
    if (total == 0u) return;
    npairs += n;
    for (uint32_t base = 0; base < total; base += 256u) {
        uint32_t tl[8], sl[8], id[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t f = base + (uint32_t)k * 32u + lane;
            uint32_t j = 0;  // number of lanes whose inclusive prefix is <= f == the owner of item f (32 when f >= total)
#pragma unroll
            for (uint32_t st = 16; st > 0; st >>= 1) { if (__shfl_sync(full, incl, (j + st - 1u) & 31u) <= f) j += st; }
            j &= 31u;
            const uint32_t l = f - __shfl_sync(full, excl, j), ntx = max(__shfl_sync(full, r.ntx, j), 1u);
            const uint32_t tx0 = __shfl_sync(full, r.tx0, j), ty0 = __shfl_sync(full, r.ty0, j), lb = __shfl_sync(full, r.layer_base, j);
            const uint32_t row = l / ntx, col = l - row * ntx;
            tl[k] = lb + (ty0 + row) * p.tiles_x + tx0 + col;
            id[k] = __shfl_sync(full, tri, j);
            sl[k] = 0u;
            if (f < total) sl[k] = atomicAdd(p.tile_count + tl[k], 1u);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (base + (uint32_t)k * 32u + lane < total) {
                bin_store(p, tl[k], sl[k], id[k], over);
            }
        }
    }
}

// CTA level, for the few primitives that cover more than 256 tiles (a cube face at 1080p covers thousands): up to 32 of
// them per CTA are walked by all its threads, one primitive after the other.  Returns true for a lane whose primitive
// was taken.  Must be reached by every thread of the CTA.
__device__ __forceinline__ bool bin_huge_cta(const Params& p, bool huge, const TileRect& r, uint32_t nt, uint32_t tri, bool& over, uint32_t& npairs) {
    constexpr uint32_t HUGE_SLOTS = 32;
    __shared__ uint32_t huge_n;
    __shared__ uint4 huge_a[HUGE_SLOTS];  // tx0, ty0, ntx, nt
    __shared__ uint2 huge_b[HUGE_SLOTS];  // layer_base, tri
    if (threadIdx.x == 0) huge_n = 0u;
    __syncthreads();
    if (huge) {
        const uint32_t k = atomicAdd(&huge_n, 1u);
        if (k < HUGE_SLOTS) {
            huge_a[k] = make_uint4(r.tx0, r.ty0, r.ntx, nt);
            huge_b[k] = make_uint2(r.layer_base, tri);
            npairs += nt;
        } else {
            huge = false;
        }
    }
    __syncthreads();
    const uint32_t hn = min(huge_n, HUGE_SLOTS);
    for (uint32_t e = 0; e < hn; ++e) {
        const uint4 a = huge_a[e];
        const uint2 b = huge_b[e];
        for (uint32_t base = 0; base < a.w; base += 8u * blockDim.x) {
            uint32_t tl[8], sl[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t i = base + (uint32_t)k * blockDim.x + threadIdx.x;
                const uint32_t row = i / a.z, col = i - row * a.z;
                tl[k] = b.x + (a.y + row) * p.tiles_x + a.x + col;
                sl[k] = 0u;
                if (i < a.w) sl[k] = atomicAdd(p.tile_count + tl[k], 1u);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (base + (uint32_t)k * blockDim.x + threadIdx.x < a.w) {
                    bin_store(p, tl[k], sl[k], b.y, over);
                }
            }
        }
    }
    return huge;
}

constexpr uint32_t SETUP_LOCAL_TILES = 32;  // primitives covering up to this many tiles are binned by their own thread
#ifndef EUC_SETUP_MIN_CTAS
#define EUC_SETUP_MIN_CTAS 1
#endif
template <class P> __device__ __forceinline__ void setup_body(const Params& p, const uint32_t tri, const bool live, uint32_t (*rec_stage)[RecLayout<P>::WORDS]) {
    using L = RecLayout<P>;
    uint2 bbox = make_uint2(0u, 0u);
    uint32_t layer = 0;
    bool oob = false;
    TileRect r;
    bool valid = false;
    uint32_t nt = 0;
    uint32_t sl[8];  // bin slots of a small primitive (fast path), taken early and consumed after the record store

    if (live) {
        const uint32_t d = find_draw(p, tri);
        const DrawDev dr = draw_of(p, d);
        layer = dr.layer;
        const typename P::Uniforms& u = uniforms_of<P>(p, d);
        const uint32_t s0 = dr.first + 3u * (tri - dr.tri_begin);

        float hx[3], hy[3], hz[3], hw[3];
        float var[3][P::V > 0 ? P::V : 1];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            // index.rs:52-54: &verts[idx]; every stream element re-runs the vertex shader (no vertex cache)
            long long vi = p.indices ? (long long)__ldg(p.indices + s0 + i) + dr.base_vertex : (long long)(s0 + i) + dr.base_vertex;
            if (vi < 0 || vi >= (long long)p.n_vertices) { oob = true; vi = 0; }
            float4 clip;
            P::vertex(u, p.vertices + (size_t)vi * p.vstride, clip, var[i]);
            hx[i] = clip.x * 1.0f;       // triangles.rs:61  (flip[0] == 1.0)
            hy[i] = clip.y * p.flip_y;
            hz[i] = clip.z;
            hw[i] = clip.w;
        }
        // Row-restricted renders (multi-GPU bands): every rank runs setup over all primitives, so the ones that miss its
        // rows leave right after the vertex stage, on a y-only evaluation of the same bounds as below (:64, :110-139).
        bool offband_early = false;
        if (p.row_begin > 0u || p.row_end < p.h) {
            const float size_yf = (float)p.h;
            const float s0 = size_yf * ((hy[0] / hw[0]) * -0.5f + 0.5f), s1 = size_yf * ((hy[1] / hw[1]) * -0.5f + 0.5f),
                        s2 = size_yf * ((hy[2] / hw[2]) * -0.5f + 0.5f);
            const uint32_t eby0 = r_as_usize_clamped(r_min(r_min(s0, s1), s2) + 0.0f, 0u, p.h);
            const uint32_t eby1 = r_as_usize_clamped(r_max(r_max(s0, s1), s2) + 1.0f, 0u, p.h);
            offband_early = eby1 <= p.row_begin || eby0 >= p.row_end || eby1 <= eby0;
        }
        float ex[3] = {0.0f, 0.0f, 0.0f}, ey[3] = {0.0f, 0.0f, 0.0f}, ez[3] = {0.0f, 0.0f, 0.0f};
        float winding = 0.0f;
        bool culled = offband_early;
        if (!offband_early) {
            // triangles.rs:64
#pragma unroll
            for (int i = 0; i < 3; ++i) { ex[i] = hx[i] / hw[i]; ey[i] = hy[i] / hw[i]; ez[i] = hz[i] / hw[i]; }
            // triangles.rs:67-70: cross(e1-e0, e2-e0).z
            const float ax = ex[1] - ex[0], ay = ey[1] - ey[0];
            const float bx = ex[2] - ex[0], by = ey[2] - ey[0];
            winding = ax * by - ay * bx;
            if (p.cull != EUC_CULL_NONE) {
                const float cull_dir = p.cull == EUC_CULL_BACK ? 1.0f : -1.0f;
                culled = winding * cull_dir < 0.0f;  // :73-77
            }
        }
        if (oob) culled = true;
        // :78-80 reverse vertex order when winding >= 0 (conditional swap of vertices 0 and 2)
        const bool rev = !culled && winding >= 0.0f;
        if (rev) {
            float t;
            t = hx[0]; hx[0] = hx[2]; hx[2] = t;  t = hy[0]; hy[0] = hy[2]; hy[2] = t;
            t = hz[0]; hz[0] = hz[2]; hz[2] = t;  t = hw[0]; hw[0] = hw[2]; hw[2] = t;
            t = ex[0]; ex[0] = ex[2]; ex[2] = t;  t = ey[0]; ey[0] = ey[2]; ey[2] = t;
            t = ez[0]; ez[0] = ez[2]; ez[2] = t;
        }

        // :110-111 verts_screen and :114-139 bounds (clamped to the whole target here; the band clamp is applied per row in
        // the raster kernel).  They only need the divided coordinates, so they come first: the bin appends below are then
        // in flight during the rest of the set-up arithmetic and the record store.
        float scx[3] = {0.0f, 0.0f, 0.0f}, scy[3] = {0.0f, 0.0f, 0.0f};
        const float size_x = (float)p.w, size_y = (float)p.h;
        if (!culled) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                scx[i] = size_x * (ex[i] * 0.5f + 0.5f);
                scy[i] = size_y * (ey[i] * -0.5f + 0.5f);
            }
            const uint32_t bx0 = r_as_usize_clamped(r_min(r_min(scx[0], scx[1]), scx[2]) + 0.0f, 0u, p.w);
            const uint32_t by0 = r_as_usize_clamped(r_min(r_min(scy[0], scy[1]), scy[2]) + 0.0f, 0u, p.h);
            const uint32_t bx1 = r_as_usize_clamped(r_max(r_max(scx[0], scx[1]), scx[2]) + 1.0f, 0u, p.w);
            const uint32_t by1 = r_as_usize_clamped(r_max(r_max(scy[0], scy[1]), scy[2]) + 1.0f, 0u, p.h);
            bbox = make_uint2(bx0 | (bx1 << 16), by0 | (by1 << 16));
            // row-restricted renders (multi-GPU bands): primitives that miss this rank's rows are dropped here
            if (by1 <= p.row_begin || by0 >= p.row_end || bx1 <= bx0 || by1 <= by0) bbox = make_uint2(0u, 0u);
        }
        p.tri_bbox[tri] = bbox;  // read by the exact path's fill pass and by raster warps that scan all primitives (TILE_LOST)
        valid = tile_rect(p, bbox, layer, r);
        nt = valid ? r.ntx * r.nty : 0u;
        if (p.bin_cap && valid && nt <= SETUP_LOCAL_TILES) {
            // fast path, primitives binned by their own thread: every tile owns bin_cap slots; the appends of (the first)
            // eight tiles are issued back to back so that their round trips overlap, and the slots are consumed after the
            // record has been written
            uint32_t jj = 0, ii = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if ((uint32_t)k < nt) {
                    sl[k] = atomicAdd(p.tile_count + (r.layer_base + (r.ty0 + jj) * p.tiles_x + r.tx0 + ii), 1u);
                    if (++ii == r.ntx) { ii = 0; ++jj; }
                }
            }
        }

        uint32_t* rec = rec_stage[threadIdx.x];
        if (bbox.x | bbox.y) {  // neither culled nor off-band
            // :86-102 coords_to_weights
            const float a0 = hx[0], a1 = hy[0], a3 = hw[0];
            const float b0 = hx[1], b1 = hy[1], b3 = hw[1];
            const float c0 = hx[2], c1 = hy[2], c2 = hw[2];  // c = [c.x, c.y, c.w]
            const float ca0 = a0 - c0, ca1 = a1 - c1, ca2 = a3 - c2;
            const float cb0 = b0 - c0, cb1 = b1 - c1, cb2 = b3 - c2;
            // n = cross(ca, cb)
            const float n0 = ca1 * cb2 - ca2 * cb1, n1 = ca2 * cb0 - ca0 * cb2, n2 = ca0 * cb1 - ca1 * cb0;
            float rec_det = 1.0f;
            if (n0 * n0 + n1 * n1 + n2 * n2 > 0.0f) rec_det = 1.0f / r_min(n0 * c0 + n1 * c1 + n2 * c2, -1.1920929e-07f);
            // rows: cross(cb, c), cross(c, ca), n — each scaled by rec_det
            float m[3][3];
            m[0][0] = (cb1 * c2 - cb2 * c1) * rec_det; m[0][1] = (cb2 * c0 - cb0 * c2) * rec_det; m[0][2] = (cb0 * c1 - cb1 * c0) * rec_det;
            m[1][0] = (c1 * ca2 - c2 * ca1) * rec_det; m[1][1] = (c2 * ca0 - c0 * ca2) * rec_det; m[1][2] = (c0 * ca1 - c1 * ca0) * rec_det;
            m[2][0] = n0 * rec_det; m[2][1] = n1 * rec_det; m[2][2] = n2 * rec_det;
            const float sx = 2.0f / size_x, sy = -2.0f / size_y;  // to_ndc :44-48
            float cw[3][3];  // matmul(m, to_ndc) :345-355, all nine products kept
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                cw[i][0] = m[i][0] * sx + m[i][1] * 0.0f + m[i][2] * 0.0f;
                cw[i][1] = m[i][0] * 0.0f + m[i][1] * sy + m[i][2] * 0.0f;
                cw[i][2] = m[i][0] * -1.0f + m[i][1] * 1.0f + m[i][2] * 1.0f;
            }
            // :142-145 finite-difference weight deltas
            float o[3], wdx[3], wdy[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                o[i] = cw[i][0] * 0.0f + cw[i][1] * 0.0f + cw[i][2] * 1.0f;
                float atx = cw[i][0] * 1000.0f + cw[i][1] * 0.0f + cw[i][2] * 1.0f;
                float aty = cw[i][0] * 0.0f + cw[i][1] * 1000.0f + cw[i][2] * 1.0f;
                wdx[i] = (atx - o[i]) * (1.0f / 1000.0f);
                wdy[i] = (aty - o[i]) * (1.0f / 1000.0f);
            }
            // :148-171 order by y
            const float min_y = r_min(r_min(scy[0], scy[1]), scy[2]);
            int o0, o1, o2;
            if (scy[0] == min_y) { if (scy[1] < scy[2]) { o0 = 0; o1 = 1; o2 = 2; } else { o0 = 0; o1 = 2; o2 = 1; } }
            else if (scy[1] == min_y) { if (scy[0] < scy[2]) { o0 = 1; o1 = 0; o2 = 2; } else { o0 = 1; o1 = 2; o2 = 0; } }
            else { if (scy[0] < scy[1]) { o0 = 2; o1 = 0; o2 = 1; } else { o0 = 2; o1 = 1; o2 = 0; } }
            // :173 z clip classification
            bool nvc = true;
            if (p.zclip) {
#pragma unroll
                for (int i = 0; i < 3; ++i) nvc = nvc && (p.zmin <= ez[i] && ez[i] <= p.zmax);
            }
            float4* r4 = reinterpret_cast<float4*>(rec);
            r4[0] = make_float4(o[0], o[1], o[2], wdx[0]);
            r4[1] = make_float4(wdx[1], wdx[2], wdy[0], wdy[1]);
            r4[2] = make_float4(wdy[2], hz[0], hz[1], hz[2]);
            const float sxs[3] = {scx[0], scx[1], scx[2]}, sys[3] = {scy[0], scy[1], scy[2]};
            auto pick = [&](const float* a, int k) { return k == 0 ? a[0] : (k == 1 ? a[1] : a[2]); };
            r4[3] = make_float4(pick(sxs, o0), pick(sys, o0), pick(sxs, o1), pick(sys, o1));
            r4[4] = make_float4(pick(sxs, o2), pick(sys, o2), __uint_as_float(bbox.x), __uint_as_float(bbox.y));
            r4[5] = make_float4(__uint_as_float(nvc ? 1u : 0u), __uint_as_float(d), __uint_as_float(tri), 0.0f);
            if constexpr (P::V > 0) {
                float flat[L::VPAD];
#pragma unroll
                for (int k = 0; k < L::VPAD; ++k) flat[k] = 0.0f;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int k = 0; k < P::V; ++k) flat[i * P::V + k] = (i == 1) ? var[1][k] : ((i == 0) != rev ? var[0][k] : var[2][k]);
#pragma unroll
                for (int k = 0; k < L::VPAD / 4; ++k) r4[6 + k] = make_float4(flat[4 * k], flat[4 * k + 1], flat[4 * k + 2], flat[4 * k + 3]);
            }
        }
    }
    if (__any_sync(0xffffffffu, oob) && oob) atomicOr(p.counters + 3, 1ull);
    if (p.sparse_recs) {  // uniform
        if (bbox.x | bbox.y) {
            const float4* src = reinterpret_cast<const float4*>(rec_stage[threadIdx.x]);
            float4* dst = reinterpret_cast<float4*>(p.recs + (size_t)tri * L::WORDS);
#pragma unroll
            for (int k = 0; k < L::WORDS / 4; ++k) dst[k] = src[k];
        }
    } else {   // one bulk store per warp: records of primitives [warp_first, warp_first + nrec)
        const uint32_t warp_first = tri - (threadIdx.x & 31u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes above, async-proxy read below
        __syncwarp();
        if ((threadIdx.x & 31u) == 0 && warp_first < p.n_tris) {
            const uint32_t nrec = min(32u, p.n_tris - warp_first);
            bulk_s2g(p.recs + (size_t)warp_first * L::WORDS, rec_stage[threadIdx.x], nrec * (uint32_t)L::BYTES);
            bulk_commit_wait_read();  // the stage must stay intact until the copy engine has read it
        }
    }

    // binning
    uint32_t npairs = 0;
    if (p.bin_cap) {
        // fast path: a tile that needs more than bin_cap slots flags the render, which is then redone on the exact
        // count -> alloc -> fill path
        bool over = false;
        if (valid && nt <= SETUP_LOCAL_TILES) {
            // chunks of eight tiles: all appends of a chunk, then all its stores (the first chunk's appends were issued above).
            // A warp-cooperative walk (below) serialises its 32 primitives on the append round trip; with teapot-sized
            // triangles (9 .. 32 tiles at 4K) that made the whole setup kernel three times longer.
            uint32_t jj = 0, ii = 0;
            for (uint32_t k0 = 0; k0 < nt; k0 += 8u) {
                uint32_t tl[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (k0 + (uint32_t)k < nt) {
                        tl[k] = r.layer_base + (r.ty0 + jj) * p.tiles_x + r.tx0 + ii;
                        if (k0) sl[k] = atomicAdd(p.tile_count + tl[k], 1u);
                        if (++ii == r.ntx) { ii = 0; ++jj; }
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (k0 + (uint32_t)k < nt) {
                        bin_store(p, tl[k], sl[k], tri, over);
                    }
                }
            }
            npairs += nt;
        }
        bool big = valid && nt > SETUP_LOCAL_TILES;
        if (p.cta_bin) {  // uniform: the host enables it for renders with few primitives (barriers cost ~15 us at 2^20)
            if (bin_huge_cta(p, big && nt > 256u, r, nt, tri, over, npairs)) big = false;
        }
        bin_big_warp(p, big, r, nt, tri, over, npairs);
        if (over) atomicOr(p.counters + 3, p.ovf_cap ? FLAG_OVF_FULL : FLAG_BINS);
    } else {
        for_each_tile(p, valid, r, tri, [&](uint32_t tile, uint32_t) { atomicAdd(p.tile_count + tile, 1u); ++npairs; });
    }
    // total pairs (one atomic per warp)
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) npairs += __shfl_xor_sync(0xffffffffu, npairs, s);
    if ((threadIdx.x & 31u) == 0 && npairs) atomicAdd(p.counters + 0, (unsigned long long)npairs);
}

template <class P, bool LIST = false> __global__ void __launch_bounds__(128, EUC_SETUP_MIN_CTAS) setup_kernel(const __grid_constant__ Params p) {
    using L = RecLayout<P>;
    __shared__ __align__(128) uint32_t rec_stage[128][L::WORDS];
    if constexpr (!LIST) {  // one thread per primitive of the render
        const uint32_t tri = blockIdx.x * blockDim.x + threadIdx.x;
        setup_body<P>(p, tri, tri < p.n_tris, rec_stage);
        return;
    }
    // list mode (row bands): a machine-sized grid strides over the (sliced) list of primitives that meet this rank's rows;
    // the host does not know its length, and a grid sized for all primitives would spend its time launching empty CTAs
    uint32_t pre[EUC_MAX_GROUP + 1];
    pre[0] = 0;
    if (p.surv_counts) {
#pragma unroll
        for (uint32_t k = 0; k < EUC_MAX_GROUP; ++k) pre[k + 1] = pre[k] + (k < p.surv_slices ? *(volatile const uint32_t*)(p.surv_counts + k) : 0u);
    } else {
#pragma unroll
        for (uint32_t k = 0; k < EUC_MAX_GROUP; ++k) pre[k + 1] = (uint32_t)*(volatile unsigned long long*)(p.counters + 9);
    }
    const uint32_t total = pre[EUC_MAX_GROUP];
    for (uint32_t base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const bool live = i < total;
        uint32_t tri = 0;
        if (live) {
            uint32_t sl = 0, start = 0;
#pragma unroll
            for (uint32_t k = 1; k < EUC_MAX_GROUP; ++k) if (p.surv_counts && i >= pre[k]) { sl = k; start = pre[k]; }
            tri = p.survivors[(size_t)sl * p.surv_stride + (i - start)];
        }
        setup_body<P>(p, tri, live, rec_stage);
    }
}

// Group pre-pass: like band_classify_kernel, but this rank looks at its share of the frame's primitives only and routes
// each to the rank(s) whose row band it meets, by appending the id to its own slice of that rank's list (peer store).
// Slices cannot overflow: a slice holds at most this rank's share.  The fills travel with the group barrier that follows.
template <class P> __global__ void __launch_bounds__(256) group_classify_kernel(const __grid_constant__ Params p) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t r_lo = 1, r_hi = 0;  // destination ranks [r_lo, r_hi]
    const uint32_t tri = p.cls_first + k;
    if (k < p.cls_n) {
        const uint32_t d = find_draw(p, tri);
        const DrawDev dr = draw_of(p, d);
        const typename P::Uniforms& u = uniforms_of<P>(p, d);
        const uint32_t s0 = dr.first + 3u * (tri - dr.tri_begin);
        float sy[3];
        bool oob = false;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            long long vi = p.indices ? (long long)__ldg(p.indices + s0 + i) + dr.base_vertex : (long long)(s0 + i) + dr.base_vertex;
            if (vi < 0 || vi >= (long long)p.n_vertices) { oob = true; vi = 0; }
            float4 clip;
            float var[P::V > 0 ? P::V : 1];
            P::vertex(u, p.vertices + (size_t)vi * p.vstride, clip, var);
            const float hy = clip.y * p.flip_y;
            sy[i] = (float)p.h * ((hy / clip.w) * -0.5f + 0.5f);
        }
        const uint32_t eby0 = r_as_usize_clamped(r_min(r_min(sy[0], sy[1]), sy[2]) + 0.0f, 0u, p.h);
        const uint32_t eby1 = r_as_usize_clamped(r_max(r_max(sy[0], sy[1]), sy[2]) + 1.0f, 0u, p.h);
        if (oob) { r_lo = 0; r_hi = p.cls_world - 1u; }  // every rank's set-up kernel sees (and reports) the bad index
        else if (eby1 > eby0) { r_lo = eby0 / p.cls_band_rows; r_hi = min((eby1 - 1u) / p.cls_band_rows, p.cls_world - 1u); }
    }
    // One global atomic per CTA and destination (same-address atomics serialise in L2: one per warp already cost more
    // than everything else here): warps take their offsets from shared-memory counters, the CTA takes one range per
    // destination, and the lanes bound for rank r write consecutive slots.
    __shared__ uint32_t cta_cnt[EUC_MAX_GROUP], cta_base[EUC_MAX_GROUP];
    const uint32_t lane = threadIdx.x & 31u;
    if (threadIdx.x < EUC_MAX_GROUP) cta_cnt[threadIdx.x] = 0u;
    __syncthreads();
    uint32_t woff[EUC_MAX_GROUP];
#pragma unroll
    for (uint32_t r = 0; r < EUC_MAX_GROUP; ++r) {
        const uint32_t m = __ballot_sync(0xffffffffu, r_lo <= r && r <= r_hi);
        uint32_t o = 0;
        if (m && lane == 0) o = atomicAdd(&cta_cnt[r], (uint32_t)__popc(m));
        woff[r] = __shfl_sync(0xffffffffu, o, 0) + __popc(m & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (threadIdx.x < EUC_MAX_GROUP) cta_base[threadIdx.x] = cta_cnt[threadIdx.x] ? atomicAdd(p.cls_counts + threadIdx.x, cta_cnt[threadIdx.x]) : 0u;
    __syncthreads();
#pragma unroll
    for (uint32_t r = 0; r < EUC_MAX_GROUP; ++r)
        if (r_lo <= r && r <= r_hi) p.cls_lists[r][(size_t)p.cls_rank * p.surv_stride + cta_base[r] + woff[r]] = tri;
}

// -------------------------------------------------------------------------------------------------------
// Lines (src/rasterizer/lines.rs:12-120).  The record of a line re-uses the triangle record's size and the word
// positions of bounds / flags / draw / primitive id, so binning and list handling are shared.
// -------------------------------------------------------------------------------------------------------
enum : int {
    LN_SX0 = 0, LN_SY0 = 1,  // verts_screen[0]
    LN_NORM = 2,             // 1 / (major-axis extent in screen space)  :77-82
    LN_EZ0 = 3, LN_EZ1 = 4,  // verts_euc[i][2]
    LN_MINY = 5, LN_MAXY1 = 6,  // min(sy0, sy1) + 0., max(sy0, sy1) + 1.  (clamped to the band per row, :66-73)
    LN_X1 = 8, LN_Y1 = 10, LN_X2 = 12, LN_Y2 = 14  // integer end points as i64 (:60-61)
};
constexpr uint32_t LN_USE_X = 1u, LN_SKIP = 2u;

// `f as isize`
__device__ __forceinline__ long long r_as_isize(float f) { return __float2ll_rz(f); }
// f32::clamp (NaN stays NaN)
__device__ __forceinline__ float r_clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Minor-axis offset of pixel i of the walk that clipline 0.2 performs (restated, unpinned): Bresenham with the error
// recurrence err = 2*dminor - dmajor; if (err > 0) { minor += s; err -= 2*dmajor; } err += 2*dminor, in closed form.
__device__ __forceinline__ long long bres_minor(long long i, long long dminor, long long dmajor) {
    if (i == 0) return 0;
    if ((dminor | dmajor) < (1ll << 30)) return (2 * dminor * i + dmajor - 1) / (2 * dmajor);
    return (long long)(((__int128)2 * dminor * i + dmajor - 1) / ((__int128)2 * dmajor));
}

template <class P> __global__ void __launch_bounds__(128) setup_lines_kernel(const __grid_constant__ Params p) {
    using L = RecLayout<P>;
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = li < p.n_tris;
    uint2 bbox = make_uint2(0u, 0u);
    uint32_t layer = 0;
    bool oob = false;
    if (live) {
        const uint32_t d = find_draw(p, li);
        const DrawDev dr = draw_of(p, d);
        layer = dr.layer;
        const typename P::Uniforms& u = uniforms_of<P>(p, d);
        const uint32_t l0 = li - dr.tri_begin;
        uint32_t s[2];
        if (p.prim_kind == EUC_PRIM_LINE_LIST) {  // primitives.rs:89-103
            s[0] = dr.first + 2u * l0; s[1] = s[0] + 1u;
        } else {                                   // LineTriangleList, primitives.rs:56-76: a b, b c, c a
            const uint32_t t = l0 / 3u, k = l0 - 3u * t;
            s[0] = dr.first + 3u * t + k; s[1] = dr.first + 3u * t + (k + 1u) % 3u;
        }
        float hx[2], hy[2], hz[2], hw[2];
        float var[2][P::V > 0 ? P::V : 1];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            long long vi = p.indices ? (long long)__ldg(p.indices + s[i]) + dr.base_vertex : (long long)s[i] + dr.base_vertex;
            if (vi < 0 || vi >= (long long)p.n_vertices) { oob = true; vi = 0; }
            float4 clip;
            P::vertex(u, p.vertices + (size_t)vi * p.vstride, clip, var[i]);
            hx[i] = clip.x * 1.0f; hy[i] = clip.y * p.flip_y; hz[i] = clip.z; hw[i] = clip.w;  // lines.rs:44
        }
        if (!oob) {
            float ex[2], ey[2], ez[2], sx[2], sy[2];
            const float size_x = (float)p.w, size_y = (float)p.h;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float w = r_max(hw[i], 0.0001f);  // :48
                ex[i] = hx[i] / w; ey[i] = hy[i] / w; ez[i] = hz[i] / w;
                sx[i] = size_x * (ex[i] * 0.5f + 0.5f);   // :53-54
                sy[i] = size_y * (ey[i] * -0.5f + 0.5f);
            }
            const long long x1 = r_as_isize(sx[0]), y1 = r_as_isize(sy[0]), x2 = r_as_isize(sx[1]), y2 = r_as_isize(sy[1]);  // :60-61
            const float minx = r_min(sx[0], sx[1]) + 0.0f, maxx1 = r_max(sx[0], sx[1]) + 1.0f;
            const float miny = r_min(sy[0], sy[1]) + 0.0f, maxy1 = r_max(sy[0], sy[1]) + 1.0f;
            // x window: the band spans the whole width (:63-64, :69-70); y bounds here are the union over all bands
            const long long wx1 = r_as_isize(r_clampf(minx, 0.0f, size_x)), wx2 = r_as_isize(r_clampf(maxx1, 0.0f, size_x));
            const long long wy1 = r_as_isize(r_clampf(miny, 0.0f, size_y)), wy2 = r_as_isize(r_clampf(maxy1, 0.0f, size_y));
            // (x1 - x2).abs() > (y1 - y2).abs() with wrapping arithmetic (overflow checks are off in the reference profile)
            auto wabs = [](long long a, long long b) { long long df = (long long)((unsigned long long)a - (unsigned long long)b); return df < 0 ? (long long)(0ull - (unsigned long long)df) : df; };
            const bool use_x = wabs(x1, x2) > wabs(y1, y2);                          // :76
            const float norm = 1.0f / (use_x ? sx[1] - sx[0] : sy[1] - sy[0]);      // :77-82
            const long long LIM = 1ll << 62;
            const bool skip = x1 <= -LIM || x1 >= LIM || x2 <= -LIM || x2 >= LIM || y1 <= -LIM || y1 >= LIM || y2 <= -LIM || y2 >= LIM;
            const bool empty = skip || wx2 <= wx1 || wy2 <= wy1 || wy2 <= (long long)p.row_begin || wy1 >= (long long)p.row_end;
            if (!empty) {
                bbox = make_uint2((uint32_t)wx1 | ((uint32_t)wx2 << 16), (uint32_t)wy1 | ((uint32_t)wy2 << 16));
                uint32_t* rec = p.recs + (size_t)li * L::WORDS;
                float4* r4 = reinterpret_cast<float4*>(rec);
                r4[0] = make_float4(sx[0], sy[0], norm, ez[0]);
                r4[1] = make_float4(ez[1], miny, maxy1, 0.0f);
                *reinterpret_cast<longlong2*>(rec + LN_X1) = make_longlong2(x1, y1);
                *reinterpret_cast<longlong2*>(rec + LN_X2) = make_longlong2(x2, y2);
                r4[4] = make_float4(0.0f, 0.0f, __uint_as_float(bbox.x), __uint_as_float(bbox.y));
                r4[5] = make_float4(__uint_as_float(use_x ? LN_USE_X : 0u), __uint_as_float(d), __uint_as_float(li), 0.0f);
                if constexpr (P::V > 0) {
                    float flat[L::VPAD];
#pragma unroll
                    for (int k = 0; k < L::VPAD; ++k) flat[k] = 0.0f;
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int k = 0; k < P::V; ++k) flat[i * P::V + k] = var[i][k];
#pragma unroll
                    for (int k = 0; k < L::VPAD / 4; ++k) r4[6 + k] = make_float4(flat[4 * k], flat[4 * k + 1], flat[4 * k + 2], flat[4 * k + 3]);
                }
            }
        }
        p.tri_bbox[li] = bbox;
    }
    if (__any_sync(0xffffffffu, oob) && oob) atomicOr(p.counters + 3, 1ull);
    TileRect r;
    bool valid = live && tile_rect(p, bbox, layer, r);
    uint32_t npairs = 0;
    if (p.bin_cap) {
        bool over = false;
        for_each_tile(p, valid, r, li, [&](uint32_t tile, uint32_t t) {
            bin_store(p, tile, atomicAdd(p.tile_count + tile, 1u), t, over);
            ++npairs;
        });
        if (over) atomicOr(p.counters + 3, p.ovf_cap ? FLAG_OVF_FULL : FLAG_BINS);
    } else {
        for_each_tile(p, valid, r, li, [&](uint32_t tile, uint32_t) { atomicAdd(p.tile_count + tile, 1u); ++npairs; });
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) npairs += __shfl_xor_sync(0xffffffffu, npairs, sft);
    if ((threadIdx.x & 31u) == 0 && npairs) atomicAdd(p.counters + 0, (unsigned long long)npairs);
}

// -------------------------------------------------------------------------------------------------------
// K3: tile list allocation, fill, order restore
// -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) alloc_tiles_kernel(const __grid_constant__ Params p, uint32_t n_tiles) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = t < n_tiles ? p.tile_count[t] : 0u;
    // warp-aggregated allocation: one atomic per warp, tiles of a warp get adjacent slices
    uint32_t lane = threadIdx.x & 31u, incl = n;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, s); if (lane >= (uint32_t)s) incl += v; }
    uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 31 && total) base = atomicAdd(p.counters + 2, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (t < n_tiles) p.tile_range[t] = make_uint2((uint32_t)base + incl - n, n);
    uint32_t mx = n;
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, sft));
    if (lane == 0 && mx) atomicMax(p.counters + 1, (unsigned long long)mx);  // [1] doubles as "longest list" until raster counts fragments
    // the pair total is final here (setup has completed): flag a list that is too small for this render
    if (t == 0 && p.counters[0] > (unsigned long long)p.list_capacity) atomicOr(p.counters + 3, 2ull);
}

// Non-zero when this render must not proceed: a vertex index was out of range (bit 0) or the pair list is too small
// (bit 1).  Written before fill/raster start (stream order); the host re-launches them after growing the list.
__device__ __forceinline__ bool render_aborted(const Params& p) { return (*(volatile unsigned long long*)(p.counters + 3) & (FLAG_OOB | FLAG_BINS)) != 0ull; }

__global__ void __launch_bounds__(128) fill_kernel(const __grid_constant__ Params p) {
    if (render_aborted(p)) return;
    const uint32_t tri = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = tri < p.n_tris;
    uint2 bbox = live ? p.tri_bbox[tri] : make_uint2(0u, 0u);
    const uint32_t layer = live ? draw_of(p, find_draw(p, tri)).layer : 0u;
    TileRect r;
    bool valid = live && tile_rect(p, bbox, layer, r);
    for_each_tile(p, valid, r, tri, [&](uint32_t tile, uint32_t t) {
        uint32_t slot = atomicSub(p.tile_count + tile, 1u) - 1u;  // leaves tile_count zeroed for the next render
        uint32_t off = p.tile_range[tile].x + slot;
        if (off < p.list_capacity) p.tile_list[off] = t;
    });
}

// Tile lists are filled with atomics, i.e. in arbitrary order; the raster kernel restores submission order (ascending
// primitive index, src/pipeline.rs:581 semantics) before it walks a list: in registers for up to 128 entries, in
// shared memory for up to SORT_SMEM, in place in global memory beyond that.
constexpr int SORT_SMEM = 2048;
__device__ __forceinline__ void bitonic_mem(uint32_t* a, uint32_t n_pow2, uint32_t lane) {
    for (uint32_t k = 2; k <= n_pow2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = lane; i < n_pow2; i += 32u) {
                uint32_t ixj = i ^ j;
                if (ixj > i) {
                    uint32_t x = a[i], y = a[ixj];
                    bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncwarp();
        }
    }
}
// Sorts the n <= 128 ids held 4 per lane (element index = r * 32 + lane, padded with 0xffffffff) ascending.
__device__ __forceinline__ void bitonic_regs128(uint32_t (&v)[4], uint32_t lane) {
#pragma unroll
    for (uint32_t k = 2; k <= 128; k <<= 1) {
#pragma unroll
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {  // partner lives in the same lane, register r ^ (j / 32)
                const uint32_t dr = j >> 5;
#pragma unroll
                for (uint32_t r = 0; r < 4; ++r) {
                    if ((r & dr) == 0) {
                        const uint32_t i = r * 32;  // (i & k) only depends on r here because k > 32
                        const bool up = (i & k) == 0;
                        const uint32_t a = v[r], b = v[r | dr];
                        const uint32_t lo = min(a, b), hi = max(a, b);
                        v[r] = up ? lo : hi;
                        v[r | dr] = up ? hi : lo;
                    }
                }
            } else {
#pragma unroll
                for (uint32_t r = 0; r < 4; ++r) {
                    const uint32_t o = __shfl_xor_sync(0xffffffffu, v[r], j);
                    const uint32_t i = r * 32 + lane;
                    const bool up = (i & k) == 0, lower = (lane & j) == 0;
                    v[r] = (lower == up) ? min(v[r], o) : max(v[r], o);
                }
            }
        }
    }
}

// The same for n <= 32 ids, one per lane (most tiles of an icon batch): 15 exchange steps instead of 100.
__device__ __forceinline__ void bitonic_regs32(uint32_t& v, uint32_t lane) {
#pragma unroll
    for (uint32_t k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            const uint32_t o = __shfl_xor_sync(0xffffffffu, v, j);
            const bool up = (lane & k) == 0, lower = (lane & j) == 0;
            v = (lower == up) ? min(v, o) : max(v, o);
        }
    }
}

// -------------------------------------------------------------------------------------------------------
// mbarrier / bulk-copy primitives (PTX ISA 8.x; SASS: SYNCS.*, UBLKCP)
// -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 16-byte asynchronous global -> shared copy with per-thread addresses (SASS: LDGSTS.E.BYPASS.128)
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// -------------------------------------------------------------------------------------------------------
// K4 / K5: tile raster.
//
// One warp per 16x16 tile, lane = (row, half) -> 8 consecutive pixels whose depth and colour stay in registers for
// the whole tile list.  Setup records arrive in shared memory in batches of BATCH through cp.async.bulk (one bulk
// copy per record, issued by BATCH different lanes, completion on an mbarrier; two stages).
//
// Lanes do NOT walk the batch in lock-step: lane t first computes which lanes triangle t's bounds touch, the 32x32
// bit matrix is transposed with ballots, and every lane then iterates only over the triangles that overlap its own
// segment (its own cursor, submission order preserved per lane — which is all euc's ordering semantics need, because
// different lanes own disjoint pixels).  This is what keeps warp execution efficiency up for small triangles.
//
// DEFER (pipelines whose blend ignores the old pixel and whose fragment stage is pure): the loop only resolves, per
// pixel, the LAST triangle whose fragment passed the depth test; fragment + blend run once per pixel afterwards.
// With any depth mode that is the reference's final pixel: blend(_, fragment(last passing)) (pipeline.rs:574-576).
// -------------------------------------------------------------------------------------------------------
constexpr int RASTER_WARPS = 4;
// Resident CTAs per SM the raster kernel is compiled for (register cap 65536 / (128 * n)).  Measured on C4 / C5: 6 CTAs
// (80 registers, no spills) beat the unconstrained 107-register build by 12 % / 14 %: the kernel is issue bound and needs
// the warps; 7 CTAs (72 registers, still no spills in the triangle kernels) gain another 2.4 % / 3.4 %.  The
// immediate-mode MSAA instantiations keep four corner fragments live and would spill.
#ifndef EUC_RASTER_MIN_CTAS
#define EUC_RASTER_MIN_CTAS 7
#endif
#ifndef EUC_SORT32
#define EUC_SORT32 1             // lists of up to 32 ids are sorted by a one-register network
#endif
#ifndef EUC_ROUND64_MAX_REC_BYTES
#define EUC_ROUND64_MAX_REC_BYTES 128u
#endif
constexpr int BATCH = 32;  // setup records per bulk-copy stage (== warp size: one lane-mask per lane)
constexpr uint32_t NO_WINNER = 0xffffffffu;

// Immediate-mode pipelines (blend reads the old pixel) queue passing fragments per lane and shade them with all
// lanes in lock-step: entry = triangle-in-batch << 8 | 8-bit mask of this lane's pixels that passed.
#ifndef EUC_Q_ENTRIES
#define EUC_Q_ENTRIES 12
#endif
#ifndef EUC_Q_FRAGS
#define EUC_Q_FRAGS 32
#endif
// Records above EUC_Q_BIG_REC_BYTES (the voxel icons' 192 bytes) get a shorter FIFO: with 12 entries the warp's shared memory
// is 8240 bytes and six CTAs fit an SM; with 10 it is 7984 and the seventh fits (the register budget allows seven): the icon
// batch's tile kernel 2.59 -> 2.45 ms per 4096 icons.
#ifndef EUC_Q_ENTRIES_BIG_REC
#define EUC_Q_ENTRIES_BIG_REC 10
#endif
#ifndef EUC_Q_BIG_REC_BYTES
#define EUC_Q_BIG_REC_BYTES 160u
#endif
template <class P> struct QGeom {
    static constexpr int ENTRIES = RecLayout<P>::WORDS * 4u > EUC_Q_BIG_REC_BYTES ? EUC_Q_ENTRIES_BIG_REC : EUC_Q_ENTRIES;  // entries per lane (u16)
    static constexpr int STRIDE_WORDS = (ENTRIES / 2) | 1;  // lane stride is odd in words -> conflict-free banks
    static_assert(ENTRIES >= 2 && ENTRIES <= 2 * STRIDE_WORDS, "fragment FIFO geometry");
};
constexpr int Q_FRAGS = EUC_Q_FRAGS;      // stop generating once a lane holds this many fragments
constexpr int COL_STRIDE = 9;     // colour row of a lane: 8 words + 1 pad

constexpr int IDS_REGS = 128;      // lists up to this length are sorted, and then kept, in registers (4 per lane)

// Shared-memory stage of the raster kernel.  Deferred pipelines only need the 24-word base of a record in the tile loop
// (varyings are read by resolve_kernel), so only that part is copied.  A round holds 64 records when they are small
// and 32 when they are large: the stage size decides how many CTAs fit an SM (ncu on C5: 3 CTAs/SM, issue-active 49 %).
template <class P, bool DEFER> struct StageGeom {
    static constexpr uint32_t REC_WORDS = DEFER ? (uint32_t)REC_BASE_WORDS : (uint32_t)RecLayout<P>::WORDS;  // words of a record kept in the stage
    static constexpr uint32_t BATCHES = REC_WORDS * 4u <= EUC_ROUND64_MAX_REC_BYTES ? 2u : 1u;                                    // batches of 32 records per round
    static constexpr uint32_t WORDS = BATCHES * BATCH * REC_WORDS;                                           // stage words per warp
};

// per warp: record stage | 16 B (mbarriers) | 32 B (tiles noted for the slow pass) | per-lane fragment queues | per-lane colour rows
constexpr uint32_t WARP_CTRL_BYTES = 48u;
template <class P, bool DEFER> struct WarpSmem {
    static constexpr uint32_t BYTES = StageGeom<P, DEFER>::WORDS * 4u + WARP_CTRL_BYTES + ((!DEFER && P::HAS_FRAGMENT) ? 32u * (QGeom<P>::STRIDE_WORDS + COL_STRIDE) * 4u : 0u);
};
template <class P, bool DEFER> constexpr size_t raster_smem_bytes() { return (size_t)RASTER_WARPS * WarpSmem<P, DEFER>::BYTES; }

// interpolate() reading the setup record from shared memory with 128-bit loads (lanes address different records)
template <class P> __device__ __forceinline__ void interpolate_smem(const float4* __restrict__ rec4, float xf, float yf, float* var) {
    const float4 q0 = rec4[0], q1 = rec4[1], q2 = rec4[2];  // o0 o1 o2 dx0 | dx1 dx2 dy0 dy1 | dy2 z0 z1 z2
    const float wh0 = (q0.x + q1.z * yf) + q0.w * xf;
    const float wh1 = (q0.y + q1.w * yf) + q1.x * xf;
    const float wh2 = (q0.z + q2.x * yf) + q1.y * xf;
    const float wu2 = wh2 - wh0 - wh1;
    const float r = 1.0f / wh2;
    const float w0 = wh0 * r, w1 = wh1 * r, w2 = wu2 * r;
    constexpr int NV4 = RecLayout<P>::VPAD / 4;
    float vv[NV4 * 4 > 0 ? NV4 * 4 : 1];
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
        const float4 t = rec4[R_VAR / 4 + k];
        vv[4 * k] = t.x; vv[4 * k + 1] = t.y; vv[4 * k + 2] = t.z; vv[4 * k + 3] = t.w;
    }
#pragma unroll
    for (int k = 0; k < P::V; ++k) var[k] = vv[k] * w0 + vv[P::V + k] * w1 + vv[2 * P::V + k] * w2;
}

// get_v_data (triangles.rs:274-294): closed-form weights at (x, y), perspective divide, weighted_sum3 (math.rs:38-40)
template <class P> __device__ __forceinline__ void interpolate(const float* __restrict__ rec, float xf, float yf, float* var) {
    const float wh0 = (rec[R_O + 0] + rec[R_DY + 0] * yf) + rec[R_DX + 0] * xf;
    const float wh1 = (rec[R_O + 1] + rec[R_DY + 1] * yf) + rec[R_DX + 1] * xf;
    const float wh2 = (rec[R_O + 2] + rec[R_DY + 2] * yf) + rec[R_DX + 2] * xf;
    const float wu2 = wh2 - wh0 - wh1;
    const float r = 1.0f / wh2;
    const float w0 = wh0 * r, w1 = wh1 * r, w2 = wu2 * r;
    const float* v0 = rec + R_VAR;
    const float* v1 = v0 + P::V;
    const float* v2 = v1 + P::V;
#pragma unroll
    for (int k = 0; k < P::V; ++k) var[k] = v0[k] * w0 + v1[k] * w1 + v2[k] * w2;
}

// get_v_data of a line (lines.rs:100-113): weighted_sum2(v0, v1, 1 - frac, frac) with frac along the major axis
template <class P> __device__ __forceinline__ void interpolate_line(const float* __restrict__ rec, float xf, float yf, float* var) {
    const bool use_x = (__float_as_uint(rec[R_FLAGS]) & LN_USE_X) != 0u;
    const float frac = (use_x ? xf - rec[LN_SX0] : yf - rec[LN_SY0]) * rec[LN_NORM];
    const float om = 1.0f - frac;
    const float* v0 = rec + R_VAR;
    const float* v1 = v0 + P::V;
#pragma unroll
    for (int k = 0; k < P::V; ++k) var[k] = v0[k] * om + v1[k] * frac;
}

template <class P, bool LINES = false>
__device__ __forceinline__ void shade_at(const typename P::Uniforms& u, const SamplerDev* samp, const float* rec, float xf, float yf, float* frag) {
    float var[P::V > 0 ? P::V : 1];
    if (LINES) interpolate_line<P>(rec, xf, yf, var);
    else interpolate<P>(rec, xf, yf, var);
    P::fragment(u, samp, var, frag);
}

// One out-of-line copy of interpolate + fragment for the MSAA corner shades: four inlined copies of a large shader
// (Phong with powf) per pixel made the resolve kernel instruction-cache bound (ncu: stall_no_instruction dominant).
template <class P, bool LINES>
__device__ __noinline__ void shade_corner(const typename P::Uniforms* u, const SamplerDev* samp, const float* rec, float xf, float yf, float* frag) {
    shade_at<P, LINES>(*u, samp, rec, xf, yf, frag);
}

// euc's coarse-shading "MSAA" (pipeline.rs:544-570) for one pixel; corner fragments are memoised in the reference
// (pure function of corner and primitive), so recomputing them is exact.  `cache` carries the corner pair of the
// previous pixel group: cx0 of this group may equal cx1 of the previous one.
struct CornerCache {
    uint32_t tri, cx, cy0;  // owner triangle, corner x, corner-row base of the cached column
    float t0[4], t1[4];     // fragments at (cx, cy0) and (cx, cy1)
};
template <class P, bool LINES = false>
__device__ __forceinline__ void msaa_fragment(const typename P::Uniforms& u, const SamplerDev* samp, const float* rec, uint32_t tri, uint32_t x,
                                               uint32_t y, uint32_t band_lo, uint32_t Lv, CornerCache& left, CornerCache& right, float* frag) {
    const float msaa_div = 1.0f / (float)(1u << Lv);
    const uint32_t rx = x, ry = y - band_lo;  // x - tgt_min[0], y - tgt_min[1]
    const float fractx = r_fract((float)rx * msaa_div), fracty = r_fract((float)ry * msaa_div);
    const uint32_t posix = rx >> Lv, posiy = ry >> Lv;
    const uint32_t cx0 = (posix + 0u) << Lv, cx1 = (posix + 1u) << Lv;
    const uint32_t cy0 = band_lo + ((posiy + 0u) << Lv), cy1 = band_lo + ((posiy + 1u) << Lv);
    if (!(left.tri == tri && left.cx == cx0 && left.cy0 == cy0)) {
        if (right.tri == tri && right.cx == cx0 && right.cy0 == cy0) {
            left = right;
        } else {
            shade_corner<P, LINES>(&u, samp, rec, (float)cx0, (float)cy0, left.t0);
            shade_corner<P, LINES>(&u, samp, rec, (float)cx0, (float)cy1, left.t1);
            left.tri = tri; left.cx = cx0; left.cy0 = cy0;
        }
    }
    if (!(right.tri == tri && right.cx == cx1 && right.cy0 == cy0)) {
        shade_corner<P, LINES>(&u, samp, rec, (float)cx1, (float)cy0, right.t0);
        shade_corner<P, LINES>(&u, samp, rec, (float)cx1, (float)cy1, right.t1);
        right.tri = tri; right.cx = cx1; right.cy0 = cy0;
    }
    const float omy = 1.0f - fracty, omx = 1.0f - fractx;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float t0 = left.t0[c] * omy + left.t1[c] * fracty;    // weighted_sum2(t00, t01, 1-fy, fy)
        const float t1 = right.t0[c] * omy + right.t1[c] * fracty;  // weighted_sum2(t10, t11, 1-fy, fy)
        frag[c] = t0 * omx + t1 * fractx;                           // weighted_sum2(t0, t1, 1-fx, fx)
    }
}

// One pixel of the coverage + depth loop on the fast path (LESS / GREATER with depth write, no per-fragment z clip),
// triangles.rs:262-271, :301 and pipeline.rs:519-538.  Written in PTX so that the whole test stays ONE predicate chain
// (range, three weights through a NaN-propagating 3-input min, depth) and the chain advance is three predicated adds:
// the C++ form compiled to a select per condition and per weight, and the kernel is bound by the ALU pipe (ncu).
// Every f32 operation carries .rn and is therefore never contracted; operation order is the reference's.
//   d     sgn*depth of pixel J (register)        mask  bit J set when the fragment passed
//   w0..2 chain values at pixel J, advanced to J+1 when J >= jlo       [jlo, jhi) pixels of the segment inside row_range
template <int J, bool DEFER, bool GREATER>
__device__ __forceinline__ void px_step(float& d, uint32_t& cwj, float& w0, float& w1, float& w2, uint32_t& mask, const float dx0,
                                        const float dx1, const float dx2, const float z0, const float z1, const float z2, const float dsgn,
                                        const uint32_t jlo, const uint32_t jhi, const uint32_t tri) {
// the 3-input min (FMNMX3) needs PTX ISA 8.8 (CUDA 12.9); older NVRTC builds (e.g. the one bundled with torch, which
// wins the dlopen when torch is loaded first) get two 2-input ones
#if defined(__CUDACC_VER_MAJOR__) && (__CUDACC_VER_MAJOR__ > 12 || (__CUDACC_VER_MAJOR__ == 12 && __CUDACC_VER_MINOR__ >= 9))
#define EUC_PX_MIN3 "min.NaN.f32 m, %1, %2, wu;\n"
#else
#define EUC_PX_MIN3 "min.NaN.f32 m, %1, %2;\n min.NaN.f32 m, m, wu;\n"
#endif
#define EUC_PX_BODY                                                                                                     \
    "{\n"                                                                                                               \
    ".reg .pred pa, pp;\n"                                                                                              \
    ".reg .f32 wu, z, t, m;\n"                                                                                          \
    "setp.le.u32 pa, %13, %15;\n"          /* jlo <= J: the chain advances from here on (:301) */                        \
    "setp.gt.and.u32 pp, %14, %15, pa;\n"  /* J < jhi: inside row_range (:262) */                                        \
    "sub.rn.f32 wu, %3, %1;\n"                                                                                          \
    "sub.rn.f32 wu, wu, %2;\n"             /* :264 w_unbalanced[2] = w2 - w0 - w1 */                                     \
    "mul.rn.f32 z, %9, %1;\n"                                                                                           \
    "mul.rn.f32 t, %10, %2;\n"                                                                                          \
    "add.rn.f32 z, z, t;\n"                                                                                             \
    "mul.rn.f32 t, %11, wu;\n"                                                                                          \
    "add.rn.f32 z, z, t;\n"                /* :269 z = z0*w0 + z1*w1 + z2*wu */                                          \
    EUC_PX_SGN                             /* GREATER keeps -depth in the registers: negate z (exact) */                 \
    EUC_PX_MIN3                            /* all three >= 0 (:267); NaN in any weight fails like the three compares */  \
    "setp.ge.and.f32 pp, m, 0f00000000, pp;\n"                                                                          \
    "setp.lt.and.f32 pp, z, %0, pp;\n"     /* pipeline.rs:519-526 */                                                     \
    "@pp mov.f32 %0, z;\n"                 /* pipeline.rs:536-538 */                                                     \
    "@pp or.b32 %4, %4, %16;\n"
#define EUC_PX_TAIL                                                                                                     \
    "@pa add.rn.f32 %1, %1, %6;\n"                                                                                      \
    "@pa add.rn.f32 %2, %2, %7;\n"                                                                                      \
    "@pa add.rn.f32 %3, %3, %8;\n"                                                                                      \
    "}\n"
#define EUC_PX_OPERANDS                                                                                                 \
    : "+f"(d), "+f"(w0), "+f"(w1), "+f"(w2), "+r"(mask), "+r"(cwj)                                                      \
    : "f"(dx0), "f"(dx1), "f"(dx2), "f"(z0), "f"(z1), "f"(z2), "f"(dsgn), "r"(jlo), "r"(jhi), "n"(J), "n"(1 << J), "r"(tri)
    if constexpr (GREATER) {
#define EUC_PX_SGN "neg.f32 z, z;\n"
        if constexpr (DEFER) asm(EUC_PX_BODY "@pp mov.b32 %5, %17;\n" EUC_PX_TAIL EUC_PX_OPERANDS);
        else asm(EUC_PX_BODY EUC_PX_TAIL EUC_PX_OPERANDS);
#undef EUC_PX_SGN
    } else {
#define EUC_PX_SGN
        if constexpr (DEFER) asm(EUC_PX_BODY "@pp mov.b32 %5, %17;\n" EUC_PX_TAIL EUC_PX_OPERANDS);
        else asm(EUC_PX_BODY EUC_PX_TAIL EUC_PX_OPERANDS);
#undef EUC_PX_SGN
    }
#undef EUC_PX_OPERANDS
#undef EUC_PX_BODY
#undef EUC_PX_MIN3
#undef EUC_PX_TAIL
}

// Packed f32x2 arithmetic (sm_100 FADD2 / FMUL2) was tried for the chain and rejected: ptxas contracts a mul.f32x2 that
// feeds an add.f32x2 into FFMA2 even with .rn and -fmad=false (fatal for bit-exactness), a predicated FADD2 compiles to
// FADD2 + two selects, and the pack / unpack moves of the prefix replay cost more than the saved additions (+2 % time).

// One 16x16 tile, walked by one warp.
// Not inlined on purpose: inside the persistent loop the register allocation of the (large) tile body got worse.
// Returns (mbarrier phase after the tile, fragments emitted).
template <class P, bool MSAA, bool DEFER, bool LINES, bool SLOW>
__device__ __forceinline__ uint2 raster_tile(const Params& p, const uint32_t tile, const uint32_t lane, uint32_t* const recs_sm, uint64_t* const bar,
                                          uint32_t phase, uint16_t* const queue, uint32_t* const col_sm, const uint32_t cnt_raw) {
    using L = RecLayout<P>;
    uint32_t nfrag = 0;
    constexpr uint32_t SW = StageGeom<P, DEFER>::REC_WORDS;
    constexpr uint32_t NB = StageGeom<P, DEFER>::BATCHES;
    constexpr bool QUEUE = !DEFER && P::HAS_FRAGMENT;
    uint2 rg;
    uint32_t n_bin;  // entries of the list that live in the tile's own bin / slice; the rest (bin overflow) in `ext`
    // SLOW: the tile's bin overflowed.  Either the rest of its list sits in the overflow buffer (collected into `ext`
    // below), or pairs were dropped (TILE_LOST) and every primitive of the render is tested against the tile.  A separate
    // instantiation, run after the ordinary tiles, so that the ordinary tile loop carries none of this.
    const bool scan_all = SLOW && (cnt_raw & TILE_LOST) != 0u;
    if (p.bin_cap) {
        const uint32_t cnt_t = cnt_raw & ~TILE_LOST;  // cnt_raw: tile_count[tile], read by the caller
        rg = make_uint2(tile * p.bin_cap, scan_all ? 1u : cnt_t);
        __syncwarp();
        if (lane == 0 && cnt_raw) p.tile_count[tile] = 0u;  // leave the counters zeroed for the next render
        n_bin = min(cnt_t, p.bin_cap);
    } else {
        rg = p.tile_range[tile];
        n_bin = rg.y;
    }
    const uint32_t n = rg.y;
    // lists that come close to the bin size are reported (the host then enlarges the bins of later renders); no running maximum
    // is carried through the tile loop
    if (p.summary && n * 4u > p.bin_cap * 3u && lane == 0) atomicMax(p.counters + 7, (unsigned long long)n);
    if (n == 0) {
        // A tile without primitives still owes its rows to the mirrors (fused gather; immediate-mode pipelines, deferred
        // ones forward from resolve_kernel) and, under a fused clear, the clear values to its own targets.
        const bool fwd = p.n_mirrors && !DEFER && P::HAS_FRAGMENT && p.pixel_write;
        const bool clr_px = (p.clear_mask & 1u) && !DEFER, clr_z = (p.clear_mask & 2u) != 0u;
        if (fwd || clr_px || clr_z) {
            const uint32_t tpl = p.tiles_x * p.tiles_y, lay = tile / tpl, tl0 = tile - lay * tpl;
            const uint32_t ty0 = tl0 / p.tiles_x, tx0 = tl0 - ty0 * p.tiles_x;
            const uint32_t yy = ty0 * TILE + (lane >> 1), sx = tx0 * TILE + (lane & 1u) * 8u;
            if (yy < p.h && yy >= p.row_begin && yy < p.row_end && sx < p.w) {
                const size_t bs = (size_t)lay * p.w * p.h + (size_t)yy * p.w + sx;
                if (sx + 8u <= p.w && (p.w & 3u) == 0) {
                    if (clr_z) {
                        const float4 z4 = make_float4(p.clear_z, p.clear_z, p.clear_z, p.clear_z);
                        *reinterpret_cast<float4*>(p.depth + bs) = z4;
                        *reinterpret_cast<float4*>(p.depth + bs + 4) = z4;
                    }
                    if (fwd || clr_px) {
                        uint4 a, b2;
                        if (clr_px) {
                            a = b2 = make_uint4(p.clear_px, p.clear_px, p.clear_px, p.clear_px);
                            *reinterpret_cast<uint4*>(p.pixel + bs) = a;
                            *reinterpret_cast<uint4*>(p.pixel + bs + 4) = b2;
                        } else {
                            a = *reinterpret_cast<const uint4*>(p.pixel + bs); b2 = *reinterpret_cast<const uint4*>(p.pixel + bs + 4);
                        }
                        if (fwd) {
                            for (uint32_t mi = 0; mi < p.n_mirrors; ++mi) {
                                *reinterpret_cast<uint4*>(p.mirrors[mi] + bs) = a;
                                *reinterpret_cast<uint4*>(p.mirrors[mi] + bs + 4) = b2;
                            }
                        }
                    }
                } else {
                    for (uint32_t j = 0; j < 8u && sx + j < p.w; ++j) {
                        if (clr_z) p.depth[bs + j] = p.clear_z;
                        if (fwd || clr_px) {
                            uint32_t c;
                            if (clr_px) { c = p.clear_px; p.pixel[bs + j] = c; } else c = p.pixel[bs + j];
                            if (fwd) for (uint32_t mi = 0; mi < p.n_mirrors; ++mi) p.mirrors[mi][bs + j] = c;
                        }
                    }
                }
            }
        }
        return make_uint2(phase, 0u);
    }
    // restore submission order inside this tile's list (the fill pass appended with atomics)
    // Short lists: sorted element r*32+lane ends up in register v[r] of this lane, which is exactly the id this lane
    // needs when it issues the bulk copy of batch r.  Long lists are sorted in place and read back from global memory.
    const bool short_list = n <= (uint32_t)IDS_REGS;
    uint32_t* const bin = p.tile_list + rg.x;
    uint32_t* ext = nullptr;
    if (SLOW && !scan_all && n > n_bin) {
        // The bin overflowed (warp-uniform, rare: the host sizes the bins from the lists of earlier renders).  The pairs
        // that did not fit are somewhere in the overflow buffer: take a slice of `ext` and collect this tile's pairs.
        uint32_t base_e = 0;
        if (lane == 0) base_e = (uint32_t)atomicAdd(p.counters + 6, (unsigned long long)(n - n_bin));
        ext = p.ext + __shfl_sync(0xffffffffu, base_e, 0);
        const uint32_t K = (uint32_t)min(*(volatile unsigned long long*)(p.counters + 5), (unsigned long long)p.ovf_cap);
        uint32_t off = 0;
        for (uint32_t i0 = 0; i0 < K; i0 += 32u) {
            const uint32_t i = i0 + lane;
            const uint2 e = i < K ? p.ovf[i] : make_uint2(0xffffffffu, 0u);
            const bool hit = e.x == tile;
            const uint32_t b = __ballot_sync(0xffffffffu, hit);
            if (hit) ext[off + __popc(b & ((1u << lane) - 1u))] = e.y;
            off += __popc(b);
        }
        __syncwarp();
    }
    auto lst = [&](uint32_t i) -> uint32_t& { return (!SLOW || i < n_bin) ? bin[i] : ext[i - n_bin]; };
    uint32_t v[4];
    if (SLOW && scan_all) {
        v[0] = v[1] = v[2] = v[3] = 0u;  // ids come from the scan below, already in submission order
    } else if (short_list) {
#pragma unroll
        for (uint32_t r = 0; r < 4; ++r) v[r] = r * 32 + lane < n ? lst(r * 32 + lane) : 0xffffffffu;
#if EUC_SORT32
        if (n <= 32u) {  // v[1 .. 3] are all padding
            if (n > 1) bitonic_regs32(v[0], lane);
        } else
#endif
        if (n > 1) bitonic_regs128(v, lane);
    } else {
        uint32_t np2 = 1;
        while (np2 < n) np2 <<= 1;
        if (np2 <= (uint32_t)SORT_SMEM && np2 <= StageGeom<P, DEFER>::WORDS) {
            uint32_t* a = recs_sm;  // the record stages are not in use yet
            for (uint32_t i = lane; i < np2; i += 32u) a[i] = i < n ? lst(i) : 0xffffffffu;
            __syncwarp();
            bitonic_mem(a, np2, lane);
            for (uint32_t i = lane; i < n; i += 32u) lst(i) = a[i];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes above, bulk-copy writes below
        } else {
            // In place in global memory, any n: the all-ascending bitonic network (first step of every merge pairs i
            // with i ^ (k-1), the rest with i ^ j; the lower index always receives the minimum).  Elements past n act as
            // +inf, never move below n, so pairs that reach past n are simply skipped.
            auto cmpx = [&](uint32_t i, uint32_t partner) {
                if (partner > i && partner < n) {
                    const uint32_t x = lst(i), y2 = lst(partner);
                    if (x > y2) { lst(i) = y2; lst(partner) = x; }
                }
            };
            for (uint32_t k = 2; k <= np2; k <<= 1) {
                for (uint32_t i = lane; i < n; i += 32u) cmpx(i, i ^ (k - 1u));
                __syncwarp();
                for (uint32_t j = k >> 2; j > 0; j >>= 1) {
                    for (uint32_t i = lane; i < n; i += 32u) cmpx(i, i ^ j);
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    }
    auto batch_id = [&](uint32_t b) -> uint32_t {  // id of element b*32+lane of the sorted list (0 past the end)
        if (short_list) return b == 0 ? v[0] : (b == 1 ? v[1] : (b == 2 ? v[2] : v[3]));
        const uint32_t pos = b * BATCH + lane;
        return pos < n ? lst(pos) : 0u;
    };

    // tile / lane geometry
    const uint32_t tiles_per_layer = p.tiles_x * p.tiles_y;
    const uint32_t layer = tile / tiles_per_layer;
    const uint32_t tl = tile - layer * tiles_per_layer;
    const uint32_t ty = tl / p.tiles_x, tx = tl - ty * p.tiles_x;
    const uint32_t tile_x0 = tx * TILE, tile_y0 = ty * TILE;
    const uint32_t y = tile_y0 + (lane >> 1);
    const uint32_t segx0 = tile_x0 + (lane & 1u) * 8u;
    const bool row_ok = y < p.h && y >= p.row_begin && y < p.row_end && segx0 < p.w;
    const float yf = (float)y;
    // euc band of this row (pipeline.rs:341-349): tgt_min.y = band_lo, tgt_max.y = band_hi
    const uint32_t band_lo = (y / p.group_rows) * p.group_rows;
    const uint32_t band_hi = min(band_lo + p.group_rows, p.h);

    const size_t layer_off = (size_t)layer * p.w * p.h;
    const size_t base = layer_off + (size_t)y * p.w + segx0;
    const bool vec_ok = segx0 + 8u <= p.w && (p.w & 3u) == 0;
    const bool shade_px = P::HAS_FRAGMENT && p.pixel_write;
    float depth[8];
    uint32_t cw[8];  // colour (immediate mode) or winning triangle id (deferred mode)
#pragma unroll
    for (int j = 0; j < 8; ++j) { depth[j] = 0.0f; cw[j] = DEFER ? NO_WINNER : 0u; }
    // Fast depth modes: LESS and GREATER share one comparison by keeping sgn*depth in the registers (sgn = -1 for
    // GREATER; negation is exact and NaN stays NaN, so `sgn*z < sgn*old` has exactly partial_cmp's outcome).
    const bool fast_depth = p.depth_test == EUC_DEPTH_LESS || p.depth_test == EUC_DEPTH_GREATER;
    const float dsgn = p.depth_test == EUC_DEPTH_GREATER ? -1.0f : 1.0f;
    if (row_ok && p.uses_depth && (p.clear_mask & 2u)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) depth[j] = p.clear_z;
    } else if (row_ok && p.uses_depth) {
        if (vec_ok) {
            const float4 a = *reinterpret_cast<const float4*>(p.depth + base), b = *reinterpret_cast<const float4*>(p.depth + base + 4);
            depth[0] = a.x; depth[1] = a.y; depth[2] = a.z; depth[3] = a.w; depth[4] = b.x; depth[5] = b.y; depth[6] = b.z; depth[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) if (segx0 + j < p.w) depth[j] = p.depth[base + j];
        }
    }
    if (fast_depth) {
#pragma unroll
        for (int j = 0; j < 8; ++j) depth[j] = depth[j] * dsgn;
    }
    if (QUEUE && row_ok && shade_px) {
        if (p.clear_mask & 1u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) cw[j] = p.clear_px;
        } else if (vec_ok) {
            const uint4 a = *reinterpret_cast<const uint4*>(p.pixel + base), b = *reinterpret_cast<const uint4*>(p.pixel + base + 4);
            cw[0] = a.x; cw[1] = a.y; cw[2] = a.z; cw[3] = a.w; cw[4] = b.x; cw[5] = b.y; cw[6] = b.z; cw[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) if (segx0 + j < p.w) cw[j] = p.pixel[base + j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) col_sm[j] = cw[j];
    }

    // Records are processed in rounds of ROUND = 2*BATCH: both record stages are filled, then every lane walks its own
    // primitives of the round.  Longer rounds bring the busiest lane closer to the mean (the round ends when the last
    // lane is done); the load of the next round is hidden by the other warps of the SM.
    constexpr uint32_t ROUND = NB * BATCH;
    const uint32_t n_rounds = (SLOW && scan_all) ? 0xffffffffu : (n + ROUND - 1) / ROUND;
    uint32_t scan_pos = 0;  // scan mode: next primitive id to test
    for (uint32_t rd = 0; rd < n_rounds; ++rd) {
        __syncwarp();  // every lane is done with the records of the previous round
        uint32_t cnt, id0, id1 = 0u;
        if (!SLOW || !scan_all) {
            cnt = min(ROUND, n - rd * ROUND);
            id0 = batch_id(NB * rd);
            if (cnt > (uint32_t)BATCH) id1 = batch_id(NB * rd + 1u);
        } else {
            // Scan mode: collect the next (up to) 32 primitives whose bounds meet this tile's rows and columns, in id order.
            // Lane t of the round must hold the t-th hit: hits of a group of 32 candidates are picked with find-nth-set.
            cnt = 0u; id0 = 0u;
            const uint32_t tpl = p.tiles_x * p.tiles_y, lay = tile / tpl, tl0 = tile - lay * tpl;
            const uint32_t sty = (tl0 / p.tiles_x) * TILE, stx = (tl0 % p.tiles_x) * TILE;
            while (cnt < (uint32_t)BATCH && scan_pos < p.n_tris) {
                const uint32_t cand = scan_pos + lane;
                bool hit = false;
                if (cand < p.n_tris) {
                    const uint2 bb = p.tri_bbox[cand];
                    const uint32_t bx0 = bb.x & 0xffffu, bx1 = bb.x >> 16, by0 = max(bb.y & 0xffffu, p.row_begin), by1 = min(bb.y >> 16, p.row_end);
                    hit = bx1 > bx0 && by1 > by0 && bx0 < stx + TILE && bx1 > stx && by0 < sty + TILE && by1 > sty &&
                          (p.layers == 1u || draw_of(p, find_draw(p, cand)).layer == lay);
                }
                const uint32_t m = __ballot_sync(0xffffffffu, hit);
                const uint32_t take = min((uint32_t)__popc(m), (uint32_t)BATCH - cnt);
                if (lane >= cnt && lane < cnt + take) id0 = scan_pos + __fns(m, 0u, (int)(lane - cnt) + 1);
                if (take < (uint32_t)__popc(m)) {  // the round is full: resume after the last hit taken
                    scan_pos += __fns(m, 0u, (int)take) + 1u;
                    cnt += take;
                    break;
                }
                cnt += take;
                scan_pos += 32u;
            }
            if (cnt == 0u) break;
        }
        const uint32_t cnt0 = min(cnt, (uint32_t)BATCH), cnt1 = cnt - cnt0;
#ifdef EUC_RECS_TMA
        if (lane == 0) mbar_expect_tx(&bar[0], cnt * SW * 4u);
        __syncwarp();
        if (lane < cnt0) bulk_g2s(recs_sm + lane * SW, p.recs + (size_t)id0 * L::WORDS, SW * 4u, &bar[0]);
        if (NB > 1 && lane < cnt1) bulk_g2s(recs_sm + (BATCH + lane) * SW, p.recs + (size_t)id1 * L::WORDS, SW * 4u, &bar[0]);
        mbar_wait(&bar[0], phase);
        phase ^= 1u;
#else
        // Lane t copies record t with 16-byte cp.async (SASS: LDGSTS, no register staging).  One bulk (TMA) copy per
        // record was measured first: UBLKCP takes uniform-register addresses, so 32 per-lane copies compile to a
        // 9-instruction elect / broadcast / issue loop per record (5.3 % of the kernel's instructions on C4).
        if (lane < cnt0) {
            const uint32_t* src = p.recs + (size_t)id0 * L::WORDS;
            uint32_t* dst = recs_sm + lane * SW;
#pragma unroll
            for (uint32_t k = 0; k < SW / 4u; ++k) cp_async16(dst + 4u * k, src + 4u * k);
        }
        if (NB > 1 && lane < cnt1) {
            const uint32_t* src = p.recs + (size_t)id1 * L::WORDS;
            uint32_t* dst = recs_sm + (BATCH + lane) * SW;
#pragma unroll
            for (uint32_t k = 0; k < SW / 4u; ++k) cp_async16(dst + 4u * k, src + 4u * k);
        }
        cp_async_wait_all();
        __syncwarp();
#endif
        const uint32_t* stage = recs_sm;
        // the PTX pixel loop (px_step) covers the common case; a round in which any record needs the per-fragment z clip
        // (:271: some vertex failed the clip test), and every other depth mode, takes the general loops
        bool fast_px = false;
        if (!LINES && fast_depth && p.depth_write) {
            bool need_zc = false;
            if (p.zclip) {
                if (lane < cnt0) need_zc = (stage[lane * SW + R_FLAGS] & 1u) == 0u;
                if (NB > 1 && lane < cnt1) need_zc = need_zc || (stage[(BATCH + lane) * SW + R_FLAGS] & 1u) == 0u;
            }
            fast_px = !__any_sync(0xffffffffu, need_zc);
        }

        // lane t: which lanes (row, half) can triangle t cover?  Bounding box first, then per tile row the range of
        // integer x on which all three weights can be non-negative.  Each weight is linear in x (w = A + d*x); euc
        // evaluates it by an accumulated chain whose value after k additions differs from the closed form by at most
        // (k + 4) * 2^-24 * M with M <= |A| + |d| * x1.  We solve A + d*x >= -m with m at least twice that bound (so the
        // range is a superset of every pixel the chain can accept), widen by a slack for the rounding of the solve
        // itself (reciprocal + product: relative 2^-22 of |x| <= x1), and intersect the three ranges.  NaN compares
        // false everywhere and therefore never constrains or rejects.
        auto lane_mask = [&](uint32_t ri, bool valid) -> uint32_t {
        uint32_t m = 0;
        if (valid) {
            const float4* rec4 = reinterpret_cast<const float4*>(stage + ri * SW);
            const float4 q4 = rec4[4];
            const uint32_t bbx = __float_as_uint(q4.z), bby = __float_as_uint(q4.w);
            const uint32_t x0 = bbx & 0xffffu, x1 = bbx >> 16, y0 = bby & 0xffffu, y1 = bby >> 16;
            const uint32_t ra = max(y0, tile_y0) - tile_y0, rb = min(y1, tile_y0 + TILE) - tile_y0;  // rows [ra, rb) of the tile
            if (LINES) {
                if (y1 > tile_y0 && y0 < tile_y0 + TILE && rb > ra && x1 > x0) {
                    const uint32_t hi = rb >= 16u ? 0xffffffffu : ((1u << (2u * rb)) - 1u);
                    const uint32_t lo = (1u << (2u * ra)) - 1u;
                    uint32_t seg = 0;
                    if (x0 < tile_x0 + 8u && x1 > tile_x0) seg |= 0x55555555u;
                    if (x0 < tile_x0 + 16u && x1 > tile_x0 + 8u) seg |= 0xaaaaaaaau;
                    m = hi & ~lo & seg;
                }
            } else
            if (y1 > tile_y0 && y0 < tile_y0 + TILE && rb > ra && x1 > x0) {
                const bool seg0 = x0 < tile_x0 + 8u && x1 > tile_x0, seg1 = x0 < tile_x0 + 16u && x1 > tile_x0 + 8u;
                const float4 q0 = rec4[0], q1 = rec4[1], q2 = rec4[2];  // o0 o1 o2 dx0 | dx1 dx2 dy0 dy1 | dy2 z0 z1 z2
                // weight e at pixel (x, y): a + b*y + d*x in exact arithmetic (e = 0, 1 and the unbalanced third, :264)
                const float a0 = q0.x, a1 = q0.y, a2 = q0.z, d0 = q0.w, d1 = q1.x, d2 = q1.y, b0 = q1.z, b1 = q1.w, b2 = q2.x;
                const float au = a2 - a0 - a1, bu = b2 - b0 - b1, du = d2 - d0 - d1;
                const float x1f = (float)x1, ymaxf = (float)(tile_y0 + (uint32_t)TILE);
                // Error budget in weight units, one value per edge for the whole tile: the chain of k <= x1 - x0 additions,
                // its start, the rounding of a + b*y and of the fused evaluation below are each bounded by a few 2^-24 of
                // S = |a| + |b|*ymax + |d|*x1 (k + 2 for the chain, 2 for a + b*y, 3 for the fused evaluation and its two
                // coefficients); (k + 12) * 2^-23 * S is at least twice their sum.
                const float kerr = (float)(x1 - x0 + 12u) * 1.1920929e-07f;
                const float s0 = (fabsf(a0) + fabsf(b0) * ymaxf) + fabsf(d0) * x1f, s1 = (fabsf(a1) + fabsf(b1) * ymaxf) + fabsf(d1) * x1f,
                            s2 = (fabsf(a2) + fabsf(b2) * ymaxf) + fabsf(d2) * x1f;
                const float m0 = kerr * s0, m1 = kerr * s1, mu = 2.0f * (kerr * ((s0 + s1) + s2));
                const float slack = 0.01f + x1f * 1e-5f;  // pixels: rounding of the reciprocal and of the products (2^-22 of |x| <= x1)
                const float BIG = 3.0e38f;
                // Per edge the admissible x of row y is x >= t(y) (d > 0) or x <= t(y) (d < 0), t(y) = (-m - a - b*y) / d, which is
                // linear in y: t = y*P + Q, one FMA per row.  An edge that is not a bound of that kind contributes -/+BIG.  A
                // slope too small to invert (zero, denormal: the weight is constant along the row) is replaced by +-1e-30, which
                // turns t into -/+huge according to the sign of a + b*y + m, i.e. the whole row passes or fails that edge.
                float PL0, QL0, PH0, QH0, PL1, QL1, PH1, QH1, PLu, QLu, PHu, QHu;
                auto edge = [&](float a, float b, float d, float m, float& PL, float& QL, float& PH, float& QH) {
                    float i = 1.0f / d;
                    if (!(fabsf(i) < 1.0e30f)) i = copysignf(1.0e30f, d);
                    const float tp = -b * i, tq = (-m - a) * i;
                    const bool lower = i > 0.0f;
                    PL = lower ? tp : 0.0f; QL = lower ? tq - slack : -BIG;
                    PH = lower ? 0.0f : tp; QH = lower ? BIG : tq + slack;
                };
                edge(a0, b0, d0, m0, PL0, QL0, PH0, QH0);
                edge(a1, b1, d1, m1, PL1, QL1, PH1, QH1);
                edge(au, bu, du, mu, PLu, QLu, PHu, QHu);
                // segment end points clamped to the bounds: [xa, xb] inclusive pixel coordinates
                const float xa0 = (float)max(tile_x0, x0), xb0 = (float)min(tile_x0 + 7u, x1 - 1u);
                const float xa1 = (float)max(tile_x0 + 8u, x0), xb1 = (float)min(tile_x0 + 15u, x1 - 1u);
                if (!(fminf(fminf(s0, s1), s2) > 1.0e-18f)) {
                    // degenerate weights (or NaN): the budget above is meaningless, every segment inside the bounds is visited
                    m = (((seg0 ? 0x55555555u : 0u) | (seg1 ? 0xaaaaaaaau : 0u)) >> (2u * ra)) << (2u * ra);
                    if (rb < 16u) m &= (1u << (2u * rb)) - 1u;
                } else {
                    float yr = (float)(tile_y0 + ra);
                    for (uint32_t r = ra; r < rb; ++r) {
                        // NaN (from Inf - Inf or NaN vertices) is ignored by fmaxf / fminf and by the comparisons: never rejects
                        const float lo = fmaxf(fmaxf(__fmaf_rn(yr, PL0, QL0), __fmaf_rn(yr, PL1, QL1)), __fmaf_rn(yr, PLu, QLu));
                        const float hi = fminf(fminf(__fmaf_rn(yr, PH0, QH0), __fmaf_rn(yr, PH1, QH1)), __fmaf_rn(yr, PHu, QHu));
                        const float ilo = ceilf(lo), ihi = floorf(hi);  // integer pixel range that can pass
                        const bool ok0 = seg0 && !(fmaxf(xa0, ilo) > fminf(xb0, ihi));
                        const bool ok1 = seg1 && !(fmaxf(xa1, ilo) > fminf(xb1, ihi));
                        m |= ((ok0 ? 1u : 0u) | (ok1 ? 2u : 0u)) << (2u * r);
                        yr += 1.0f;
                    }
                }
            }
        }
        return m;
        };
        // transpose: own bit t <=> primitive t of the round touches this lane
        // 32x32 bit-matrix transpose across the warp (row = lane) in five butterfly stages: lanes l and l^j exchange the
        // off-diagonal j x j blocks (32 ballots cost ~6 instructions each)
        auto transpose = [&](uint32_t x) -> uint32_t {
#pragma unroll
            for (uint32_t j = 16; j >= 1; j >>= 1) {
                const uint32_t k = 0xffffffffu / ((1u << j) + 1u);  // bits whose index has bit j clear
                const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
                x = (lane & j) ? (((y >> j) & k) | (x & ~k)) : ((x & k) | ((y & k) << j));
            }
            return row_ok ? x : 0u;
        };
        uint32_t own0 = transpose(lane_mask(lane, lane < cnt0));
        uint32_t own1 = (NB > 1 && cnt1) ? transpose(lane_mask(BATCH + lane, lane < cnt1)) : 0u;

        uint32_t qn = 0, qf = 0;  // queued entries / fragments of this lane
        for (;;) {
            // ---- generate: coverage + depth for this lane's own triangles, until its queue is nearly full ----
            while ((own0 | own1) && qn < (uint32_t)QGeom<P>::ENTRIES && qf < (uint32_t)Q_FRAGS) {
                uint32_t t;
                if (own0) { t = (uint32_t)__ffs((int)own0) - 1u; own0 &= own0 - 1u; }
                else { t = (uint32_t)BATCH + (uint32_t)__ffs((int)own1) - 1u; own1 &= own1 - 1u; }
                const float4* rec4 = reinterpret_cast<const float4*>(stage + t * SW);
                if constexpr (LINES) {
                    const float* rec = reinterpret_cast<const float*>(rec4);
                    const uint32_t* recu = stage + t * SW;
                    const uint32_t wxa = recu[R_BBX] & 0xffffu, wxb = recu[R_BBX] >> 16;  // x window [wxa, wxb - 1]
                    // y window of this row's band (lines.rs:66-67, :72-73, :86)
                    const float blo = (float)band_lo, bhi = (float)band_hi;
                    const long long wy1 = r_as_isize(r_clampf(rec[LN_MINY], blo, bhi)), wy2 = r_as_isize(r_clampf(rec[LN_MAXY1], blo, bhi));
                    const long long yl = (long long)y;
                    uint32_t passmask = 0;
                    if (yl >= wy1 && yl <= wy2 - 1 && segx0 < wxb && segx0 + 8u > wxa) {
                        const longlong2 pa = *reinterpret_cast<const longlong2*>(recu + LN_X1), pb = *reinterpret_cast<const longlong2*>(recu + LN_X2);
                        const long long lx1 = pa.x, ly1 = pa.y, lx2 = pb.x, ly2 = pb.y;
                        const long long ldx = lx2 > lx1 ? lx2 - lx1 : lx1 - lx2, ldy = ly2 > ly1 ? ly2 - ly1 : ly1 - ly2;
                        const long long lsx = lx1 < lx2 ? 1 : -1, lsy = ly1 < ly2 ? 1 : -1;
                        const bool xmajor = ldx >= ldy;
                        long long xhit = 0;  // y-major: the single x of this row
                        bool row_on = true;
                        if (!xmajor) {
                            const long long i = (yl - ly1) * lsy;
                            row_on = i >= 0 && i <= ldy;
                            if (row_on) xhit = lx1 + lsx * bres_minor(i, ldx, ldy);
                        }
                        const uint32_t flags = recu[R_FLAGS];
                        const bool use_x = (flags & LN_USE_X) != 0u;
                        const float sx0 = rec[LN_SX0], sy0 = rec[LN_SY0], norm = rec[LN_NORM], ez0 = rec[LN_EZ0], ez1 = rec[LN_EZ1];
                        const uint32_t tri_id = recu[R_TRI];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t x = segx0 + (uint32_t)j;
                            bool on = row_on && x >= wxa && x < wxb;
                            if (on) {
                                if (xmajor) {
                                    const long long i = ((long long)x - lx1) * lsx;
                                    on = i >= 0 && i <= ldx && ly1 + lsy * bres_minor(i, ldy, ldx) == yl;
                                } else {
                                    on = (long long)x == xhit;
                                }
                            }
                            if (on) {
                                const float frac = (use_x ? (float)x - sx0 : yf - sy0) * norm;   // lines.rs:90-94
                                const float z = ez0 + frac * (ez1 - ez0);                       // :97
                                bool pass = !p.zclip || (p.zmin <= z && z <= p.zmax);           // :99
                                if (pass && p.depth_test != EUC_DEPTH_NONE) {
                                    const float old_z = fast_depth ? depth[j] * dsgn : depth[j];
                                    pass = p.depth_test == EUC_DEPTH_LESS ? (z < old_z) : (p.depth_test == EUC_DEPTH_EQUAL ? (z == old_z) : (z > old_z));
                                }
                                if (pass) {
                                    passmask |= 1u << j;
                                    if (p.depth_write) depth[j] = fast_depth ? z * dsgn : z;
                                    if (DEFER) cw[j] = tri_id;
                                }
                            }
                        }
                    }
                    nfrag += __popc(passmask);
                    if (QUEUE && passmask && shade_px) {
                        queue[qn++] = (uint16_t)((t << 8) | passmask);
                        qf += __popc(passmask);
                    }
                    continue;
                }
                const float4 q4 = rec4[4];  // c.x c.y bbx bby
                const uint32_t bbx = __float_as_uint(q4.z), bby = __float_as_uint(q4.w);
                const uint32_t x0 = bbx & 0xffffu, x1 = bbx >> 16, y0 = bby & 0xffffu, y1 = bby >> 16;
                // band-clamped vertical bounds (triangles.rs:114-139 with tgt_min/max of this row's band)
                const uint32_t bymin = min(max(y0, band_lo), band_hi), bymax = min(max(y1, band_lo), band_hi);
                const uint32_t extent = (x1 - x0) * (bymax - bymin);
                uint32_t r0, r1;
                if (extent < 128u) {  // :224-226
                    r0 = x0; r1 = x1;
                } else {  // :228-253
                    const float4 q3 = rec4[3];  // a.x a.y b.x b.y
                    const float a_x = q3.x, a_y = q3.y, b_x = q3.z, b_y = q3.w, c_x = q4.x, c_y = q4.y;
                    const float ac = a_x + ((yf - a_y) / (c_y - a_y)) * (c_x - a_x);
                    float lo, hi;
                    if (yf < b_y) {
                        const float ab = a_x + ((yf - a_y) / (b_y - a_y)) * (b_x - a_x);
                        lo = r_min(ab, ac); hi = r_max(ab, ac);
                    } else {
                        const float bc = b_x + ((yf - b_y) / (c_y - b_y)) * (c_x - b_x);
                        lo = r_min(bc, ac); hi = r_max(bc, ac);
                    }
                    const float e0 = floorf(lo), e1 = ceilf(hi);
                    const float fx0 = (float)x0, fx1 = (float)x1;
                    r0 = (e0 >= fx0 && e0 < fx1) ? __float2uint_rz(e0) : x0;
                    r1 = (e1 >= fx0 && e1 < fx1) ? __float2uint_rz(e1) : x1;
                }
                if (!(segx0 >= r1 || segx0 + 8u <= r0 || r1 <= r0)) {
                // pixels of this segment inside [r0, r1), and from which pixel on the chain advances (:262, :301)
                const uint32_t jlo = r0 > segx0 ? r0 - segx0 : 0u, jhi = min(r1 - segx0, 8u);
                const uint32_t inmask = ((1u << jhi) - 1u) & ~((1u << jlo) - 1u);
                // chain start (:257-260) and replay up to this lane's segment (:301)
                const float4 q0 = rec4[0], q1 = rec4[1], q2 = rec4[2];  // o0 o1 o2 dx0 | dx1 dx2 dy0 dy1 | dy2 z0 z1 z2
                const float dx0 = q0.w, dx1 = q1.x, dx2 = q1.y;
                const float r0f = (float)r0;
                float w0 = (q0.x + q1.z * yf) + dx0 * r0f;
                float w1 = (q0.y + q1.w * yf) + dx1 * r0f;
                float w2 = (q0.z + q2.x * yf) + dx2 * r0f;
                if (segx0 > r0) {
                    const uint32_t npre = segx0 - r0;
#pragma unroll 4
                    for (uint32_t i = 0; i < npre; ++i) { w0 = w0 + dx0; w1 = w1 + dx1; w2 = w2 + dx2; }
                }
                const float z0 = q2.y, z1 = q2.z, z2 = q2.w;
                const float4 q5 = rec4[5];  // flags draw tri -
                const bool zc = p.zclip && (__float_as_uint(q5.x) & 1u) == 0u;  // per-fragment z clip needed (:271)
                const uint32_t tri_id = __float_as_uint(q5.z);
                uint32_t passmask = 0;
#ifdef EUC_DBG_COUNT  // development builds only: the fragment counter reports visits (1), covered pixels (2), visits without coverage (3)
                {
                    float a0 = w0, a1 = w1, a2 = w2;
                    uint32_t cov = 0;
                    for (uint32_t j = 0; j < 8u; ++j) {
                        const float au = a2 - a0 - a1;
                        if (((inmask >> j) & 1u) && a0 >= 0.0f && a1 >= 0.0f && au >= 0.0f) ++cov;
                        if (j >= jlo) { a0 = a0 + dx0; a1 = a1 + dx1; a2 = a2 + dx2; }
                    }
                    nfrag += EUC_DBG_COUNT == 1 ? 1u : (EUC_DBG_COUNT == 2 ? cov : (cov == 0u ? 1u : 0u));
                }
#endif
                if (fast_px) {  // warp-uniform: LESS / GREATER with depth write, no record of this round needs the z clip
                    if (dsgn < 0.0f) {
                        px_step<0, DEFER, true>(depth[0], cw[0], w0, w1, w2, passmask, dx0, dx1, dx2, z0, z1, z2, dsgn, jlo, jhi, tri_id);
                        px_step<1, DEFER, true>(depth[1], cw[1], w0, w1, w2, passmask, dx0, dx1, dx2, z0, z1, z2, dsgn, jlo, jhi, tri_id);
                        px_step<2, DEFER, true>(depth[2], cw[2], w0, w1, w2, passmask, dx0, dx1, dx2, z0, z1, z2, dsgn, jlo, jhi, tri_id);
                        px_step<3, DEFER, true>(depth[3], cw[3], w0, w1, w2, passmask, dx0, dx1, dx2, z0, z1, z2, dsgn, jlo, jhi, tri_id);
                        px_step<4, DEFER, true>(depth[4], cw[4], w0, w1, w2, passmask, dx0, dx1, dx2, z0, z1, z2, dsgn, jlo, jhi, tri_id);
                        px_step<5, DEFER, true>(depth[5], cw[5], w0, w1, w2, passmask, dx0, dx1, dx2, z0, z1, z2, dsgn, jlo, jhi, tri_id);
                        px_step<6, DEFER, true>(depth[6], cw[6], w0, w1, w2, passmask, dx0, dx1, dx2, z0, z1, z2, dsgn, jlo, jhi, tri_id);
                        px_step<7, DEFER, true>(depth[7], cw[7], w0, w1, w2, passmask, dx0, dx1, dx2, z0, z1, z2, dsgn, jlo, jhi, tri_id);
                    } else {
                        px_step<0, DEFER, false>(depth[0], cw[0], w0, w1, w2, passmask, dx0, dx1, dx2, z0, z1, z2, dsgn, jlo, jhi, tri_id);
                        px_step<1, DEFER, false>(depth[1], cw[1], w0, w1, w2, passmask, dx0, dx1, dx2, z0, z1, z2, dsgn, jlo, jhi, tri_id);
                        px_step<2, DEFER, false>(depth[2], cw[2], w0, w1, w2, passmask, dx0, dx1, dx2, z0, z1, z2, dsgn, jlo, jhi, tri_id);
                        px_step<3, DEFER, false>(depth[3], cw[3], w0, w1, w2, passmask, dx0, dx1, dx2, z0, z1, z2, dsgn, jlo, jhi, tri_id);
                        px_step<4, DEFER, false>(depth[4], cw[4], w0, w1, w2, passmask, dx0, dx1, dx2, z0, z1, z2, dsgn, jlo, jhi, tri_id);
                        px_step<5, DEFER, false>(depth[5], cw[5], w0, w1, w2, passmask, dx0, dx1, dx2, z0, z1, z2, dsgn, jlo, jhi, tri_id);
                        px_step<6, DEFER, false>(depth[6], cw[6], w0, w1, w2, passmask, dx0, dx1, dx2, z0, z1, z2, dsgn, jlo, jhi, tri_id);
                        px_step<7, DEFER, false>(depth[7], cw[7], w0, w1, w2, passmask, dx0, dx1, dx2, z0, z1, z2, dsgn, jlo, jhi, tri_id);
                    }
                } else if (fast_depth) {
                    // z clip as bounds: when every vertex passed the clip the per-fragment test is skipped (:271), i.e.
                    // the bounds are infinite.  A NaN z fails here but would fail the depth comparison anyway.
                    const float zlo = zc ? p.zmin : -__int_as_float(0x7f800000), zhi = zc ? p.zmax : __int_as_float(0x7f800000);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float wu2 = w2 - w0 - w1;                                          // :264
                        const float z = z0 * w0 + z1 * w1 + z2 * wu2;                            // :269
                        const float zs = z * dsgn;
                        const bool pass = ((inmask >> j) & 1u) && w0 >= 0.0f && w1 >= 0.0f && wu2 >= 0.0f &&  // :262, :267
                                          zlo <= z && z <= zhi && zs < depth[j];                 // :271, pipeline.rs:519-526
                        if (pass) {
                            passmask |= 1u << j;
                            if (p.depth_write) depth[j] = zs;                                    // pipeline.rs:536-538
                            if (DEFER) cw[j] = tri_id;
                        }
                        if ((uint32_t)j >= jlo) { w0 = w0 + dx0; w1 = w1 + dx1; w2 = w2 + dx2; } // :301
                    }
                } else
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float wu2 = w2 - w0 - w1;                                              // :264
                    if (((inmask >> j) & 1u) && w0 >= 0.0f && w1 >= 0.0f && wu2 >= 0.0f) {      // :262, :267
                        const float z = z0 * w0 + z1 * w1 + z2 * wu2;                            // :269
                        bool pass = !zc || (p.zmin <= z && z <= p.zmax);
                        if (p.depth_test != EUC_DEPTH_NONE) {                                    // pipeline.rs:519-526
                            const float old_z = depth[j];
                            pass = pass && (p.depth_test == EUC_DEPTH_LESS ? (z < old_z) : (p.depth_test == EUC_DEPTH_EQUAL ? (z == old_z) : (z > old_z)));
                        }
                        if (pass) {
                            passmask |= 1u << j;
                            if (p.depth_write) depth[j] = z;                                     // pipeline.rs:536-538
                            if (DEFER) cw[j] = tri_id;
                        }
                    }
                    if ((uint32_t)j >= jlo) { w0 = w0 + dx0; w1 = w1 + dx1; w2 = w2 + dx2; }     // :301
                }
#ifndef EUC_DBG_COUNT
                nfrag += __popc(passmask);
#endif
                if (QUEUE && passmask && shade_px) {
                    queue[qn++] = (uint16_t)((t << 8) | passmask);
                    qf += __popc(passmask);
                }
                }
            }
            if (!QUEUE) break;  // the generate loop above ran to completion (its FIFO limits never trigger)
            // ---- drain: fragment + blend (pipeline.rs:540-577) for the queued fragments, all lanes in lock-step ----
            __syncwarp();  // reconverge: lanes leave the generate loop at different times
            {
                uint32_t k = 0, cur = 0, ct = 0;
                for (;;) {
                    if (cur == 0) {
                        if (k == qn) break;
                        const uint32_t e = queue[k++];
                        cur = e & 0xffu;
                        ct = e >> 8;
                    }
                    const uint32_t j = (uint32_t)__ffs((int)cur) - 1u;
                    cur &= cur - 1u;
                    const float4* rec4 = reinterpret_cast<const float4*>(stage + ct * SW);
                    const float4 q5 = rec4[5];
                    const typename P::Uniforms& u = uniforms_of<P>(p, __float_as_uint(q5.y));
                    float frag[4];
                    if (!MSAA) {
                        float var[P::V > 0 ? P::V : 1];
                        if (LINES) interpolate_line<P>(reinterpret_cast<const float*>(rec4), (float)(segx0 + j), yf, var);
                        else interpolate_smem<P>(rec4, (float)(segx0 + j), yf, var);
                        P::fragment(u, p.samp, var, frag);
                    } else {
                        CornerCache lc, rc;
                        lc.tri = rc.tri = NO_WINNER;
                        msaa_fragment<P, LINES>(u, p.samp, reinterpret_cast<const float*>(rec4), __float_as_uint(q5.z), segx0 + j, y, band_lo, p.msaa_level, lc, rc, frag);
                    }
                    col_sm[j] = P::blend(col_sm[j], frag);
                }
                qn = 0; qf = 0;
            }
            __syncwarp();
            if (!__any_sync(0xffffffffu, (own0 | own1) != 0u)) break;
        }
    }

    // deferred pipelines: hand the per-pixel winners to resolve_kernel (the buffer is pre-filled with NO_WINNER)
    if (DEFER && row_ok && shade_px) {
        bool any = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) any = any || cw[j] != NO_WINNER;
        if (any) {
            if (vec_ok) {
                *reinterpret_cast<uint4*>(p.winner + base) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
                *reinterpret_cast<uint4*>(p.winner + base + 4) = make_uint4(cw[4], cw[5], cw[6], cw[7]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (segx0 + j < p.w) p.winner[base + j] = cw[j];
            }
        }
    }

    if (QUEUE && row_ok && shade_px) {
#pragma unroll
        for (int j = 0; j < 8; ++j) cw[j] = col_sm[j];
    }
    // write back
    if (fast_depth) {
#pragma unroll
        for (int j = 0; j < 8; ++j) depth[j] = depth[j] * dsgn;
    }
    if (row_ok) {
        if (p.depth_write || (p.clear_mask & 2u)) {
            if (vec_ok) {
                *reinterpret_cast<float4*>(p.depth + base) = make_float4(depth[0], depth[1], depth[2], depth[3]);
                *reinterpret_cast<float4*>(p.depth + base + 4) = make_float4(depth[4], depth[5], depth[6], depth[7]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (segx0 + j < p.w) p.depth[base + j] = depth[j];
            }
        }
        if (shade_px && !DEFER) {
            if (vec_ok) {
                const uint4 a = make_uint4(cw[0], cw[1], cw[2], cw[3]), b2 = make_uint4(cw[4], cw[5], cw[6], cw[7]);
                *reinterpret_cast<uint4*>(p.pixel + base) = a;
                *reinterpret_cast<uint4*>(p.pixel + base + 4) = b2;
                for (uint32_t mi = 0; mi < p.n_mirrors; ++mi) {  // fused gather: peer stores over NVLink
                    *reinterpret_cast<uint4*>(p.mirrors[mi] + base) = a;
                    *reinterpret_cast<uint4*>(p.mirrors[mi] + base + 4) = b2;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (segx0 + j < p.w) {
                        p.pixel[base + j] = cw[j];
                        for (uint32_t mi = 0; mi < p.n_mirrors; ++mi) p.mirrors[mi][base + j] = cw[j];
                    }
                }
            }
        }
    }
    return make_uint2(phase, nfrag);
}

// Persistent kernel: every warp takes tiles from a ticket counter until none are left, so the grid is sized by the
// machine (SMs x resident CTAs), not by the frame, and there is no partial last wave.
// Tiles whose bin overflowed take another instantiation of the tile loop (SLOW).  They are rare (the host sizes the bins from
// the lists of earlier renders), and the ordinary loop must not pay for them with registers or instruction-cache space: a warp
// only notes such a tile and comes back to it after its ordinary tiles.
constexpr uint32_t SLOW_SLOTS = 7;
template <class P, bool MSAA, bool DEFER, bool LINES>
__device__ __noinline__ uint2 raster_tile_slow(const Params& p, const uint32_t tile, const uint32_t lane, uint32_t* const recs_sm, uint64_t* const bar,
                                               uint32_t phase, uint16_t* const queue, uint32_t* const col_sm, const uint32_t cnt_raw) {
    return raster_tile<P, MSAA, DEFER, LINES, true>(p, tile, lane, recs_sm, bar, phase, queue, col_sm, cnt_raw);
}
// Taking several tiles per ticket (one atomic and one round trip for the tile counters of a chunk) was measured on the icon
// batch, where a warp walks hundreds of small or empty tiles, and lost twice: fixed chunks of eight -17 %, chunks that shrink
// to single tiles towards the end -12.6 % (profiles/README.md); one tile per ticket stays.
template <class P, bool MSAA, bool DEFER, bool LINES>
__global__ void __launch_bounds__(RASTER_WARPS * 32, (MSAA && !DEFER) ? 4 : EUC_RASTER_MIN_CTAS) raster_kernel(const __grid_constant__ Params p, uint32_t n_tiles) {
    using L = RecLayout<P>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    constexpr uint32_t STW = StageGeom<P, DEFER>::WORDS;
    uint8_t* const warp_sm = smem_raw + (size_t)warp * WarpSmem<P, DEFER>::BYTES;
    uint32_t* const recs_sm = reinterpret_cast<uint32_t*>(warp_sm);
    uint64_t* const bar = reinterpret_cast<uint64_t*>(warp_sm + STW * 4);
    uint32_t* const slow = reinterpret_cast<uint32_t*>(warp_sm + STW * 4 + 16);  // [0]: noted tiles, [1 ..]: their indices
    uint32_t* const lane_sm = reinterpret_cast<uint32_t*>(warp_sm + STW * 4 + WARP_CTRL_BYTES);
    uint16_t* const queue = reinterpret_cast<uint16_t*>(lane_sm + lane * QGeom<P>::STRIDE_WORDS);   // this lane's fragment FIFO
    uint32_t* const col_sm = lane_sm + 32 * QGeom<P>::STRIDE_WORDS + lane * COL_STRIDE;             // this lane's 8 colours
    if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); slow[0] = 0u; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t phase = 0, nfrag = 0;
    if (!render_aborted(p)) {
        unsigned int* const ticket = reinterpret_cast<unsigned int*>(p.counters + 4);
        // tickets enumerate only the tile rows that intersect the rendered rows [row_begin, row_end)
        const uint32_t ty_lo = p.row_begin / TILE, ty_hi = (min(p.row_end, p.h) + TILE - 1) / TILE;
        const uint32_t per_layer = (ty_hi - ty_lo) * p.tiles_x, n_active = per_layer * p.layers;
        for (bool more = true; more;) {
            more = false;
            for (;;) {
                uint32_t tk = 0;
                if (lane == 0) tk = atomicAdd(ticket, 1u);
                tk = __shfl_sync(0xffffffffu, tk, 0);
                if (tk >= n_active) break;
                const uint32_t lay = tk / per_layer;
                const uint32_t tile = lay * p.tiles_x * p.tiles_y + ty_lo * p.tiles_x + (tk - lay * per_layer);
                if (tile >= n_tiles) break;
                const uint32_t cnt_raw = p.bin_cap ? p.tile_count[tile] : 0u;
                if (cnt_raw > p.bin_cap) {  // (only with bins) overflowed, or marked TILE_LOST: later
                    const uint32_t k = slow[0];
                    __syncwarp();
                    if (lane == 0) { slow[1u + k] = tile; slow[0] = k + 1u; }
                    __syncwarp();
                    if (k + 1u == SLOW_SLOTS) { more = true; break; }
                    continue;
                }
                const uint2 res = raster_tile<P, MSAA, DEFER, LINES, false>(p, tile, lane, recs_sm, bar, phase, queue, col_sm, cnt_raw);
                phase = res.x;
                nfrag += res.y;
                __syncwarp();
            }
            const uint32_t n_slow = slow[0];
            for (uint32_t i = 0; i < n_slow; ++i) {
                const uint32_t tile = slow[1u + i];
                const uint2 res = raster_tile_slow<P, MSAA, DEFER, LINES>(p, tile, lane, recs_sm, bar, phase, queue, col_sm, p.tile_count[tile]);
                phase = res.x;
                nfrag += res.y;
                __syncwarp();
            }
            __syncwarp();
            if (lane == 0) slow[0] = 0u;
            __syncwarp();
        }
        if (p.stats) {
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) nfrag += __shfl_xor_sync(0xffffffffu, nfrag, s);
            if (lane == 0 && nfrag) atomicAdd(p.counters + 1, (unsigned long long)nfrag);
        }
    } else if (p.summary) {
        // an aborted asynchronous render has no host behind it to restore the all-zero tile counters
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_tiles; i += gridDim.x * blockDim.x) p.tile_count[i] = 0u;
    }
    // Summary for the host: the last warp of the grid to get here publishes flags / longest list / overflow volume into
    // mapped pinned memory, so that the host can size the bins of later renders without ever waiting for this one.
    if (p.summary && lane == 0) {
        __threadfence();
        const unsigned long long done = atomicAdd(p.counters + 8, 1ull) + 1ull;
        if (done == (unsigned long long)gridDim.x * RASTER_WARPS) {
            volatile unsigned long long* c = p.counters;
            if (c[3]) atomicOr(p.counters + 15, c[3]);  // sticky until the host has seen it (several renders may finish between two looks)
            p.summary[1] = c[15];
            p.summary[2] = c[7];
            p.summary[3] = c[5];
            p.summary[4] = c[0];
            p.summary[5] = n_tiles;
            __threadfence_system();
            p.summary[0] = p.seq;
        }
    }
}

// -------------------------------------------------------------------------------------------------------
// Resolve (deferred pipelines): fragment + blend once per pixel for the winning primitive (pipeline.rs:540-577).
// One thread per pixel, 32x4-pixel CTAs: rows of a warp are contiguous, so winner loads and colour stores coalesce,
// and the whole GPU shades in parallel instead of one warp per tile.
// -------------------------------------------------------------------------------------------------------
// Without MSAA: one thread per pixel.  With MSAA a thread owns one cell of euc's shading grid (2^L x 2^L pixels whose four
// corners coincide: tgt_min + (pos << L), anchored at x = 0 and at the first row of the euc band, pipeline.rs:544-551);
// while the winning primitive stays the same the four corner fragments are shaded once for the whole cell (the
// reference memoises them per primitive and corner in `fragment_cache`, so this is the same value, not an approximation).
// Cells never straddle a band: the last cell row of a band is cut at the band's end.
template <class P, bool MSAA, bool LINES> __global__ void __launch_bounds__(128) resolve_kernel(const __grid_constant__ Params p) {
    using L = RecLayout<P>;
    if (render_aborted(p)) return;
    const uint32_t row_end = min(p.row_end, p.h);
    if (row_end <= p.row_begin) return;
    if constexpr (!MSAA) {
        // grid-stride over blocks of 32 x 4 pixels: most blocks of a frame hold no winner at all, so the grid is sized by
        // the machine and a CTA just moves on
        const uint32_t nbx = (p.w + 31u) / 32u;
        const uint32_t rows = row_end - p.row_begin, nby = (rows + 3u) / 4u;
        const uint32_t nblocks = nbx * nby * p.layers;
        for (uint32_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
            const uint32_t layer = blk / (nbx * nby), b2 = blk - layer * nbx * nby, by = b2 / nbx, bx = b2 - by * nbx;
            const uint32_t x = bx * 32u + (threadIdx.x & 31u);
            const uint32_t y = p.row_begin + by * 4u + (threadIdx.x >> 5);
            if (x >= p.w || y >= row_end) continue;
            const size_t idx = (size_t)layer * p.w * p.h + (size_t)y * p.w + x;
            const uint32_t win = p.winner[idx];
            if (win == NO_WINNER) {
                if (p.clear_mask & 1u) {  // fused clear: untouched pixels take the clear value
                    p.pixel[idx] = p.clear_px;
                    for (uint32_t mi = 0; mi < p.n_mirrors; ++mi) p.mirrors[mi][idx] = p.clear_px;
                } else if (p.n_mirrors) {  // fused gather: untouched pixels of this rank's rows are forwarded as they are
                    const uint32_t c = p.pixel[idx];
                    for (uint32_t mi = 0; mi < p.n_mirrors; ++mi) p.mirrors[mi][idx] = c;
                }
                continue;
            }
            p.winner[idx] = NO_WINNER;  // leave the buffer clean for the next render
            const float* rec = reinterpret_cast<const float*>(p.recs + (size_t)win * L::WORDS);
            const typename P::Uniforms& u = uniforms_of<P>(p, __float_as_uint(rec[R_DRAW]));
            float frag[4];
            shade_at<P, LINES>(u, p.samp, rec, (float)x, (float)y, frag);
            const uint32_t out = P::blend(p.pixel[idx], frag);
            p.pixel[idx] = out;
            for (uint32_t mi = 0; mi < p.n_mirrors; ++mi) p.mirrors[mi][idx] = out;
        }
    } else {
        // one CTA per block of 32 x 4 cells (the host sizes the grid): the hardware balances the expensive blocks
        const uint32_t Lv = p.msaa_level, cs = 1u << Lv;
        const uint32_t cells_x = (p.w + cs - 1u) >> Lv;
        const uint32_t cells_per_band = (p.group_rows + cs - 1u) >> Lv;
        const uint32_t band_first = p.row_begin / p.group_rows, band_last = (row_end - 1u) / p.group_rows;  // bands touching the rendered rows
        const uint32_t cell_rows = (band_last - band_first + 1u) * cells_per_band;
        const uint32_t nbx = (cells_x + 31u) / 32u, nby = (cell_rows + 3u) / 4u;
        const uint32_t nblocks = nbx * nby * p.layers;
        for (uint32_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
            const uint32_t layer = blk / (nbx * nby), b2 = blk - layer * nbx * nby, by = b2 / nbx, bx = b2 - by * nbx;
            const uint32_t cxi = bx * 32u + (threadIdx.x & 31u), cr = by * 4u + (threadIdx.x >> 5);
            if (cxi >= cells_x || cr >= cell_rows) continue;
            const uint32_t band = band_first + cr / cells_per_band, kc = cr % cells_per_band;
            const uint32_t band_lo = band * p.group_rows, band_hi = min(band_lo + p.group_rows, p.h);
            const uint32_t ya = max(band_lo + (kc << Lv), p.row_begin), yb = min(min(band_lo + ((kc + 1u) << Lv), band_hi), row_end);
            const uint32_t xa = cxi << Lv, xb = min(xa + cs, p.w);
            const size_t lay = (size_t)layer * p.w * p.h;
            if (p.n_mirrors || (p.clear_mask & 1u)) {
                // untouched pixels: the clear value under a fused clear; forwarded to the mirrors under the fused gather
                for (uint32_t y = ya; y < yb; ++y)
                    for (uint32_t x = xa; x < xb; ++x) {
                        const size_t idx = lay + (size_t)y * p.w + x;
                        if (p.winner[idx] == NO_WINNER) {
                            uint32_t c;
                            if (p.clear_mask & 1u) { c = p.clear_px; p.pixel[idx] = c; } else c = p.pixel[idx];
                            for (uint32_t mi = 0; mi < p.n_mirrors; ++mi) p.mirrors[mi][idx] = c;
                        }
                    }
            }
            // One round per distinct winning primitive of the cell (almost always one).  The four corner shades of a
            // round are issued before the pixel loop, so the threads of a warp run them together: shading a second
            // primitive at whatever pixel it first appears left 12 of 32 lanes active in the shader (ncu).
            const float cx0 = (float)xa, cx1 = (float)(xa + cs);
            const float cy0 = (float)(band_lo + (kc << Lv)), cy1 = (float)(band_lo + ((kc + 1u) << Lv));
            const float msaa_div = 1.0f / (float)cs;
            for (;;) {
                uint32_t cur = NO_WINNER;
                for (uint32_t y = ya; y < yb && cur == NO_WINNER; ++y)
                    for (uint32_t x = xa; x < xb && cur == NO_WINNER; ++x) cur = p.winner[lay + (size_t)y * p.w + x];
                if (cur == NO_WINNER) break;
                const float* rec = reinterpret_cast<const float*>(p.recs + (size_t)cur * L::WORDS);
                const typename P::Uniforms& u = uniforms_of<P>(p, __float_as_uint(rec[R_DRAW]));
                float t00[4], t01[4], t10[4], t11[4];  // fragments at (cx0, cy0), (cx0, cy1), (cx1, cy0), (cx1, cy1): pipeline.rs:552-561
                shade_corner<P, LINES>(&u, p.samp, rec, cx0, cy0, t00);
                shade_corner<P, LINES>(&u, p.samp, rec, cx0, cy1, t01);
                shade_corner<P, LINES>(&u, p.samp, rec, cx1, cy0, t10);
                shade_corner<P, LINES>(&u, p.samp, rec, cx1, cy1, t11);
                for (uint32_t y = ya; y < yb; ++y) {
                    const float fracty = r_fract((float)(y - band_lo) * msaa_div), omy = 1.0f - fracty;
                    for (uint32_t x = xa; x < xb; ++x) {
                        const size_t idx = lay + (size_t)y * p.w + x;
                        if (p.winner[idx] != cur) continue;
                        p.winner[idx] = NO_WINNER;  // leave the buffer clean for the next render
                        const float fractx = r_fract((float)x * msaa_div), omx = 1.0f - fractx;
                        float frag[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float t0 = t00[c] * omy + t01[c] * fracty;  // weighted_sum2(t00, t01, 1-fy, fy)  :562-570
                            const float t1 = t10[c] * omy + t11[c] * fracty;  // weighted_sum2(t10, t11, 1-fy, fy)
                            frag[c] = t0 * omx + t1 * fractx;                 // weighted_sum2(t0, t1, 1-fx, fx)
                        }
                        const uint32_t out = P::blend(p.pixel[idx], frag);
                        p.pixel[idx] = out;
                        for (uint32_t mi = 0; mi < p.n_mirrors; ++mi) p.mirrors[mi][idx] = out;
                    }
                }
            }
        }
    }
}

}  // namespace eucb
