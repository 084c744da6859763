// raster2.cuh — the tile kernel of immediate-mode pipelines (blend reads the old pixel: BASELINE configs 4 and 5), second form.
//
// raster_kernel (kernels.cuh) gives every lane one 8-pixel row segment of the tile and lets it walk "its" primitives.  That
// keeps depth in registers, but a round lasts as long as its busiest lane, and every visit issues eight pixel steps whatever
// the primitive covers there: on the 2^20-triangle scene 16 of 32 lanes and 3 of 8 pixel slots were live (ncu,
// profiles/README.md).  The cause is structural: whole visits are ordered per segment, so the longest per-segment chain
// bounds the round.
//
// Here only what euc's semantics really order is kept in order, and everything else is spread evenly over the lanes:
//
//   A  coverage + z   work item = (primitive, tile row).  Items are independent of each other (the weight chain of a row
//                     starts at row_range[0] of that row, triangles.rs:257-260) and of the depth buffer, so the items of a
//                     round are dealt to the lanes 32 at a time whatever tile row they belong to.  A lane replays the chain
//                     up to the first pixel that can be covered, then walks only the pixels the conservative per-row
//                     interval allows, and leaves (x, z) of every covered pixel in its scratch row, in x order.
//   B  depth          the candidates of the 32 items, taken 32 at a time in (item, x) order = submission order, are tested
//                     against the tile's depth in shared memory by 32 lanes at once (pipeline.rs:519-538).  Two candidates
//                     of one pixel inside a group are resolved in order (match_any).
//   C  shade + blend  fragments that passed are queued in order; whenever 32 are there they are interpolated, shaded and
//                     blended by 32 lanes at once (pipeline.rs:540-577), same-pixel fragments of a group in order.
//
// Arithmetic is the same as in raster_tile (same expressions, same order): bit-exact against the oracle.
#pragma once

namespace eucb {

#ifndef EUC_R2_MIN_CTAS
#define EUC_R2_MIN_CTAS 5
#endif
constexpr uint32_t R2_R = 32;         // records per round
constexpr uint32_t R2_ZS = 17;        // stride of a lane's scratch row (16 candidates + 1: conflict-free banks)

template <class P> struct R2Smem {
    static constexpr uint32_t REC_BYTES = R2_R * RecLayout<P>::WORDS * 4u;
    static constexpr uint32_t OFF_BAR = REC_BYTES;                        // 16 B mbarriers (slow pass) + 32 B noted tiles
    static constexpr uint32_t OFF_ITEM_ROW = OFF_BAR + 48u;               // u8 [32][16]  tile row of item k of record t
    static constexpr uint32_t OFF_ITEM_X = OFF_ITEM_ROW + 512u;           // u8 [32][16]  xs << 4 | xe (pixels of the tile, inclusive)
    static constexpr uint32_t OFF_ZS = OFF_ITEM_X + 512u;                 // f32[32][17]  z of a lane's candidates
    static constexpr uint32_t OFF_XS = OFF_ZS + 32u * R2_ZS * 4u;         // u8 [32][17]  x of a lane's candidates (padded to 560 B)
    static constexpr uint32_t OFF_DEPTH = OFF_XS + 560u;                  // f32[256]     depth of the tile, pixel = row * 16 + x
    static constexpr uint32_t OFF_COL = OFF_DEPTH + 1024u;                // u32[256]     colour of the tile
    static constexpr uint32_t OFF_FRAG = OFF_COL + 1024u;                 // u16[64]      fragments waiting to be shaded: record << 8 | pixel
    static constexpr uint32_t OFF_BAND = OFF_FRAG + 128u;                 // u32[16] band_lo, u32[16] band_hi per tile row
    static constexpr uint32_t BYTES = OFF_BAND + 128u;
    static_assert(BYTES % 16u == 0, "per-warp block must keep 16-byte alignment");
    static_assert(StageGeom<P, false>::BATCHES == 1, "the slow pass reuses the record stage: 32-record rounds only");
    static_assert(512u + 512u + 32u * R2_ZS * 4u >= 32u * (Q_STRIDE_WORDS + COL_STRIDE) * 4u, "the slow pass's queues and colour rows live in the item / scratch area");
};
template <class P> constexpr size_t raster2_smem_bytes() { return (size_t)RASTER_WARPS * R2Smem<P>::BYTES; }

// Number of lanes whose inclusive prefix is <= g: the owner of dense index g (32 when g is past the end).
__device__ __forceinline__ uint32_t r2_owner(uint32_t incl, uint32_t g) {
    uint32_t j = 0;
#pragma unroll
    for (uint32_t st = 16; st > 0; st >>= 1) { if (__shfl_sync(0xffffffffu, incl, (j + st - 1u) & 31u) <= g) j += st; }
    return j;
}

template <class P>
__device__ __forceinline__ uint32_t raster2_tile(const Params& p, const uint32_t tile, const uint32_t lane, uint8_t* const sm, const uint32_t n) {
    using L = RecLayout<P>;
    using S = R2Smem<P>;
    constexpr uint32_t SW = L::WORDS;
    uint32_t* const recs_sm = reinterpret_cast<uint32_t*>(sm);
    uint8_t* const item_row = sm + S::OFF_ITEM_ROW;
    uint8_t* const item_x = sm + S::OFF_ITEM_X;
    float* const zs = reinterpret_cast<float*>(sm + S::OFF_ZS);
    uint8_t* const xs = sm + S::OFF_XS;
    float* const depth_sm = reinterpret_cast<float*>(sm + S::OFF_DEPTH);
    uint32_t* const col_sm = reinterpret_cast<uint32_t*>(sm + S::OFF_COL);
    uint16_t* const frag_sm = reinterpret_cast<uint16_t*>(sm + S::OFF_FRAG);
    uint32_t* const band_sm = reinterpret_cast<uint32_t*>(sm + S::OFF_BAND);
    const uint32_t full = 0xffffffffu, lt = (1u << lane) - 1u;
    uint32_t nfrag = 0;

    // tile geometry; lane = (row, half) only for loading and storing the tile
    const uint32_t tiles_per_layer = p.tiles_x * p.tiles_y;
    const uint32_t layer = tile / tiles_per_layer;
    const uint32_t tl = tile - layer * tiles_per_layer;
    const uint32_t ty = tl / p.tiles_x, tx = tl - ty * p.tiles_x;
    const uint32_t tile_x0 = tx * TILE, tile_y0 = ty * TILE;
    const uint32_t y_own = tile_y0 + (lane >> 1), segx0 = tile_x0 + (lane & 1u) * 8u;
    const bool row_ok = y_own < p.h && y_own >= p.row_begin && y_own < p.row_end && segx0 < p.w;
    const size_t base = (size_t)layer * p.w * p.h + (size_t)y_own * p.w + segx0;
    const bool vec_ok = segx0 + 8u <= p.w && (p.w & 3u) == 0;
    const bool shade_px = p.pixel_write != 0;
    const uint32_t pix_own = (lane >> 1) * 16u + (lane & 1u) * 8u;

    if (n == 0) {
        // a tile without primitives still owes its rows to the mirrors (fused gather) and, under a fused clear, the clear values
        const bool fwd = p.n_mirrors && shade_px;
        const bool clr_px = (p.clear_mask & 1u) != 0u, clr_z = (p.clear_mask & 2u) != 0u;
        if ((fwd || clr_px || clr_z) && row_ok) {
#pragma unroll
            for (uint32_t j = 0; j < 8u; ++j) {
                if (segx0 + j >= p.w) break;
                if (clr_z) p.depth[base + j] = p.clear_z;
                if (fwd || clr_px) {
                    uint32_t c;
                    if (clr_px) { c = p.clear_px; p.pixel[base + j] = c; } else c = p.pixel[base + j];
                    if (fwd) for (uint32_t mi = 0; mi < p.n_mirrors; ++mi) p.mirrors[mi][base + j] = c;
                }
            }
        }
        return 0u;
    }

    // ---- the tile's list in submission order (ascending primitive id), as in raster_tile ----
    uint32_t* const list = p.tile_list + (p.bin_cap ? tile * p.bin_cap : p.tile_range[tile].x);
    const bool short_list = n <= (uint32_t)IDS_REGS;
    uint32_t v[4];
    if (short_list) {
#pragma unroll
        for (uint32_t r = 0; r < 4; ++r) v[r] = r * 32 + lane < n ? list[r * 32 + lane] : 0xffffffffu;
        if (n > 1) bitonic_regs128(v, lane);
    } else {
        uint32_t np2 = 1;
        while (np2 < n) np2 <<= 1;
        if (np2 <= (uint32_t)SORT_SMEM && np2 <= R2_R * SW) {
            uint32_t* a = recs_sm;  // the record stage is not in use yet
            for (uint32_t i = lane; i < np2; i += 32u) a[i] = i < n ? list[i] : 0xffffffffu;
            __syncwarp();
            bitonic_mem(a, np2, lane);
            for (uint32_t i = lane; i < n; i += 32u) list[i] = a[i];
        } else {
            auto cmpx = [&](uint32_t i, uint32_t partner) {
                if (partner > i && partner < n) {
                    const uint32_t x = list[i], y2 = list[partner];
                    if (x > y2) { list[i] = y2; list[partner] = x; }
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

    // ---- tile state into shared memory: depth and colour of the 256 pixels, the euc band of every row ----
    {
        float d8[8];
        uint32_t c8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { d8[j] = 0.0f; c8[j] = 0u; }
        if (row_ok && p.uses_depth) {
            if (p.clear_mask & 2u) {
#pragma unroll
                for (int j = 0; j < 8; ++j) d8[j] = p.clear_z;
            } else if (vec_ok) {
                const float4 a = *reinterpret_cast<const float4*>(p.depth + base), b = *reinterpret_cast<const float4*>(p.depth + base + 4);
                d8[0] = a.x; d8[1] = a.y; d8[2] = a.z; d8[3] = a.w; d8[4] = b.x; d8[5] = b.y; d8[6] = b.z; d8[7] = b.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (segx0 + j < p.w) d8[j] = p.depth[base + j];
            }
        }
        if (row_ok && shade_px) {
            if (p.clear_mask & 1u) {
#pragma unroll
                for (int j = 0; j < 8; ++j) c8[j] = p.clear_px;
            } else if (vec_ok) {
                const uint4 a = *reinterpret_cast<const uint4*>(p.pixel + base), b = *reinterpret_cast<const uint4*>(p.pixel + base + 4);
                c8[0] = a.x; c8[1] = a.y; c8[2] = a.z; c8[3] = a.w; c8[4] = b.x; c8[5] = b.y; c8[6] = b.z; c8[7] = b.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (segx0 + j < p.w) c8[j] = p.pixel[base + j];
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { depth_sm[pix_own + j] = d8[j]; col_sm[pix_own + j] = c8[j]; }
        if (lane < 16u) {
            const uint32_t yy = tile_y0 + lane;
            const uint32_t blo = (yy / p.group_rows) * p.group_rows;  // pipeline.rs:341-349
            band_sm[lane] = blo;
            band_sm[16u + lane] = min(blo + p.group_rows, p.h);
        }
    }
    // rows of the tile that this render produces
    const uint32_t row_lo = p.row_begin > tile_y0 ? p.row_begin - tile_y0 : 0u;
    const uint32_t row_hi = min(min(p.row_end, p.h) - tile_y0, (uint32_t)TILE);  // exclusive; the tile meets the rendered rows
    const int32_t dtest = p.depth_test;
    const bool dwrite = p.depth_write != 0;
    uint32_t fcount = 0;  // fragments waiting in frag_sm (warp-uniform)

    // shade + blend the first `cntf` (<= 32) queued fragments, one per lane
    auto shade_chunk = [&](uint32_t cntf) {
        const bool live = lane < cntf;
        const uint32_t e = live ? (uint32_t)frag_sm[lane] : 0u;
        const uint32_t ct = e >> 8, pix = e & 0xffu;
        float frag[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (live) {
            const float4* rec4 = reinterpret_cast<const float4*>(recs_sm + ct * SW);
            const typename P::Uniforms& u = uniforms_of<P>(p, __float_as_uint(rec4[5].y));
            float var[P::V > 0 ? P::V : 1];
            interpolate_smem<P>(rec4, (float)(tile_x0 + (pix & 15u)), (float)(tile_y0 + (pix >> 4)), var);
            P::fragment(u, p.samp, var, frag);
        }
        // blend in queue order per pixel (pipeline.rs:574-576): fragments of one pixel inside this group take turns
        const uint32_t same = __match_any_sync(full, live ? pix : 0x100u + lane);
        const uint32_t turn = __popc(same & lt);
        const uint32_t turns = __reduce_max_sync(full, turn);
#pragma unroll 1
        for (uint32_t r = 0; r <= turns; ++r) {
            if (live && turn == r) col_sm[pix] = P::blend(col_sm[pix], frag);
            if (turns) __syncwarp();
        }
    };

    const uint32_t n_rounds = (n + R2_R - 1) / R2_R;
    for (uint32_t rd = 0; rd < n_rounds; ++rd) {
        __syncwarp();  // the previous round's records, items and scratch are no longer read
        const uint32_t cnt = min(R2_R, n - rd * R2_R);
        uint32_t id0;
        if (short_list) id0 = rd == 0 ? v[0] : (rd == 1 ? v[1] : (rd == 2 ? v[2] : v[3]));
        else id0 = rd * R2_R + lane < n ? list[rd * R2_R + lane] : 0u;
        if (lane < cnt) {
            const uint32_t* src = p.recs + (size_t)id0 * SW;
            uint32_t* dst = recs_sm + lane * SW;
#pragma unroll
            for (uint32_t k = 0; k < SW / 4u; ++k) cp_async16(dst + 4u * k, src + 4u * k);
        }
        cp_async_wait_all();
        __syncwarp();

        // ---- items: lane t lists, for record t, the tile rows it can cover and the pixel range of each (the per-row
        // conservative interval of raster_tile's lane masks: a superset of what the chain can accept) ----
        uint32_t n_items = 0;
        if (lane < cnt) {
            const float4* rec4 = reinterpret_cast<const float4*>(recs_sm + lane * SW);
            const float4 q4 = rec4[4];
            const uint32_t bbx = __float_as_uint(q4.z), bby = __float_as_uint(q4.w);
            const uint32_t x0 = bbx & 0xffffu, x1 = bbx >> 16, y0 = bby & 0xffffu, y1 = bby >> 16;
            const uint32_t ra = max(max(y0, tile_y0) - tile_y0, row_lo), rb = min(min(y1, tile_y0 + TILE) - tile_y0, row_hi);
            if (y1 > tile_y0 && y0 < tile_y0 + TILE && rb > ra && x1 > x0 && x0 < tile_x0 + TILE && x1 > tile_x0) {
                const uint32_t xa_i = max(tile_x0, x0) - tile_x0, xb_i = min(tile_x0 + 15u, x1 - 1u) - tile_x0;  // bounds inside the tile, inclusive
                const float4 q0 = rec4[0], q1 = rec4[1], q2 = rec4[2];
                const float a0 = q0.x, a1 = q0.y, a2 = q0.z, d0 = q0.w, d1 = q1.x, d2 = q1.y, b0 = q1.z, b1 = q1.w, b2 = q2.x;
                const float au = a2 - a0 - a1, bu = b2 - b0 - b1, du = d2 - d0 - d1;
                const float x1f = (float)x1, ymaxf = (float)(tile_y0 + (uint32_t)TILE);
                const float kerr = (float)(x1 - x0 + 12u) * 1.1920929e-07f;
                const float s0 = (fabsf(a0) + fabsf(b0) * ymaxf) + fabsf(d0) * x1f, s1 = (fabsf(a1) + fabsf(b1) * ymaxf) + fabsf(d1) * x1f,
                            s2 = (fabsf(a2) + fabsf(b2) * ymaxf) + fabsf(d2) * x1f;
                if (!(fminf(fminf(s0, s1), s2) > 1.0e-18f)) {
                    // degenerate weights (or NaN): no usable bound, every row inside the bounds is walked in full
                    for (uint32_t r = ra; r < rb; ++r) { item_row[lane * 16u + n_items] = (uint8_t)r; item_x[lane * 16u + n_items] = (uint8_t)((xa_i << 4) | xb_i); ++n_items; }
                } else {
                    const float m0 = kerr * s0, m1 = kerr * s1, mu = 2.0f * (kerr * ((s0 + s1) + s2));
                    const float slack = 0.01f + x1f * 1e-5f;
                    const float BIG = 3.0e38f;
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
                    const float xaf = (float)(tile_x0 + xa_i), xbf = (float)(tile_x0 + xb_i);
                    float yr = (float)(tile_y0 + ra);
                    for (uint32_t r = ra; r < rb; ++r) {
                        const float lo = fmaxf(fmaxf(__fmaf_rn(yr, PL0, QL0), __fmaf_rn(yr, PL1, QL1)), __fmaf_rn(yr, PLu, QLu));
                        const float hi = fminf(fminf(__fmaf_rn(yr, PH0, QH0), __fmaf_rn(yr, PH1, QH1)), __fmaf_rn(yr, PHu, QHu));
                        const float fa = fmaxf(xaf, ceilf(lo)), fb = fminf(xbf, floorf(hi));  // NaN bounds are ignored: never rejects
                        if (!(fa > fb)) {
                            item_row[lane * 16u + n_items] = (uint8_t)r;
                            item_x[lane * 16u + n_items] = (uint8_t)(((__float2uint_rz(fa) - tile_x0) << 4) | (__float2uint_rz(fb) - tile_x0));
                            ++n_items;
                        }
                        yr += 1.0f;
                    }
                }
            }
        }
        uint32_t incl_i = n_items;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) { const uint32_t up = __shfl_up_sync(full, incl_i, sft); if (lane >= (uint32_t)sft) incl_i += up; }
        const uint32_t excl_i = incl_i - n_items, total_items = __shfl_sync(full, incl_i, 31);
        __syncwarp();

        for (uint32_t w0 = 0; w0 < total_items; w0 += 32u) {
            // ---- A: one item per lane ----
            const uint32_t gi = w0 + lane;
            const bool have = gi < total_items;
            const uint32_t ot = r2_owner(incl_i, gi) & 31u;             // record (lane) that listed item gi
            const uint32_t ok_ = gi - __shfl_sync(full, excl_i, ot);     // its k-th item
            uint32_t ncand = 0, it_row = 0;
            if (have) {
                it_row = item_row[ot * 16u + ok_];
                const uint32_t ix = item_x[ot * 16u + ok_];
                const uint32_t y = tile_y0 + it_row;
                const float yf = (float)y;
                const uint32_t band_lo = band_sm[it_row], band_hi = band_sm[16u + it_row];
                const float4* rec4 = reinterpret_cast<const float4*>(recs_sm + ot * SW);
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
                // pixels tested: the item's range inside row_range [r0, r1) (:262)
                const uint32_t ta = max(tile_x0 + (ix >> 4), r0), tb = min(tile_x0 + (ix & 15u) + 1u, r1);
                if (ta < tb && r1 > r0) {
                    const float4 q0 = rec4[0], q1 = rec4[1], q2 = rec4[2];  // o0 o1 o2 dx0 | dx1 dx2 dy0 dy1 | dy2 z0 z1 z2
                    const float dx0 = q0.w, dx1 = q1.x, dx2 = q1.y;
                    const float r0f = (float)r0;
                    // chain start (:257-260) and replay up to the first tested pixel (:301)
                    float w0_ = (q0.x + q1.z * yf) + dx0 * r0f;
                    float w1_ = (q0.y + q1.w * yf) + dx1 * r0f;
                    float w2_ = (q0.z + q2.x * yf) + dx2 * r0f;
                    for (uint32_t i = r0; i < ta; ++i) { w0_ = w0_ + dx0; w1_ = w1_ + dx1; w2_ = w2_ + dx2; }
                    const float z0 = q2.y, z1 = q2.z, z2 = q2.w;
                    const bool zc = p.zclip && (__float_as_uint(rec4[5].x) & 1u) == 0u;  // per-fragment z clip needed (:271)
                    const float zlo = zc ? p.zmin : -__int_as_float(0x7f800000), zhi = zc ? p.zmax : __int_as_float(0x7f800000);
                    float* const zrow = zs + lane * R2_ZS;
                    uint8_t* const xrow = xs + lane * R2_ZS;
                    for (uint32_t x = ta; x < tb; ++x) {
                        const float wu2 = w2_ - w0_ - w1_;                                          // :264
                        const float z = z0 * w0_ + z1 * w1_ + z2 * wu2;                             // :269
                        // :267 coverage, :271 z clip (inclusive; with infinite bounds a NaN z is kept and fails the depth comparison,
                        // except under DepthMode::NONE, where the reference emits it: handled below)
                        const bool cov = w0_ >= 0.0f && w1_ >= 0.0f && wu2 >= 0.0f && (!zc || (zlo <= z && z <= zhi));
                        if (cov) { zrow[ncand] = z; xrow[ncand] = (uint8_t)(x - tile_x0); ++ncand; }
                        w0_ = w0_ + dx0; w1_ = w1_ + dx1; w2_ = w2_ + dx2;                          // :301
                    }
                }
            }
            uint32_t incl_c = ncand;
#pragma unroll
            for (int sft = 1; sft < 32; sft <<= 1) { const uint32_t up = __shfl_up_sync(full, incl_c, sft); if (lane >= (uint32_t)sft) incl_c += up; }
            const uint32_t excl_c = incl_c - ncand, total_c = __shfl_sync(full, incl_c, 31);
            __syncwarp();  // scratch rows are read by other lanes below

            // ---- B: depth test of the candidates, 32 at a time in (item, x) order ----
            for (uint32_t c0 = 0; c0 < total_c; c0 += 32u) {
                const uint32_t g = c0 + lane;
                const bool live = g < total_c;
                const uint32_t ow = r2_owner(incl_c, g) & 31u;
                const uint32_t k = g - __shfl_sync(full, excl_c, ow);
                const uint32_t row_o = __shfl_sync(full, it_row, ow), rec_o = __shfl_sync(full, ot, ow);
                float z = 0.0f;
                uint32_t pix = 0x100u + lane;
                if (live) { z = zs[ow * R2_ZS + k]; pix = row_o * 16u + xs[ow * R2_ZS + k]; }
                const uint32_t same = __match_any_sync(full, pix);
                const uint32_t turn = __popc(same & lt);
                const uint32_t turns = __reduce_max_sync(full, turn);
                bool pass = false;
#pragma unroll 1
                for (uint32_t r = 0; r <= turns; ++r) {
                    if (live && turn == r) {
                        pass = true;
                        if (dtest != EUC_DEPTH_NONE) {                                              // pipeline.rs:519-526
                            const float old_z = depth_sm[pix];
                            pass = dtest == EUC_DEPTH_LESS ? (z < old_z) : (dtest == EUC_DEPTH_EQUAL ? (z == old_z) : (z > old_z));
                        }
                        if (pass && dwrite) depth_sm[pix] = z;                                      // pipeline.rs:536-538
                    }
                    if (turns) __syncwarp();
                }
                const uint32_t pm = __ballot_sync(full, pass);
                nfrag += __popc(pm);
                if (shade_px && pm) {
                    if (pass) frag_sm[fcount + __popc(pm & lt)] = (uint16_t)((rec_o << 8) | pix);
                    fcount += __popc(pm);
                    __syncwarp();
                    if (fcount >= 32u) {
                        // ---- C: 32 fragments are waiting ----
                        shade_chunk(32u);
                        __syncwarp();
                        const uint32_t rest = fcount - 32u;
                        const uint16_t mv = lane < rest ? frag_sm[32u + lane] : (uint16_t)0;
                        __syncwarp();
                        if (lane < rest) frag_sm[lane] = mv;
                        fcount = rest;
                        __syncwarp();
                    }
                }
            }
            __syncwarp();
        }
        // the round's records are about to be replaced: shade what is still queued
        if (fcount) { __syncwarp(); shade_chunk(fcount); fcount = 0; }
    }
    __syncwarp();

    // ---- write back (one coalesced 32-byte piece per lane and target, as in raster_tile) ----
    if (row_ok) {
        float d8[8];
        uint32_t c8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { d8[j] = depth_sm[pix_own + j]; c8[j] = col_sm[pix_own + j]; }
        if (p.uses_depth && (p.depth_write || (p.clear_mask & 2u))) {
            if (vec_ok) {
                *reinterpret_cast<float4*>(p.depth + base) = make_float4(d8[0], d8[1], d8[2], d8[3]);
                *reinterpret_cast<float4*>(p.depth + base + 4) = make_float4(d8[4], d8[5], d8[6], d8[7]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (segx0 + j < p.w) p.depth[base + j] = d8[j];
            }
        }
        if (shade_px) {
            if (vec_ok) {
                const uint4 a = make_uint4(c8[0], c8[1], c8[2], c8[3]), b2 = make_uint4(c8[4], c8[5], c8[6], c8[7]);
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
                        p.pixel[base + j] = c8[j];
                        for (uint32_t mi = 0; mi < p.n_mirrors; ++mi) p.mirrors[mi][base + j] = c8[j];
                    }
                }
            }
        }
    }
    return nfrag;
}

// Tiles whose bin overflowed (rare) are rendered by raster_tile's SLOW instantiation, whose per-lane queues and colour rows
// are carved out of this kernel's item / scratch area.
template <class P>
__device__ __noinline__ uint2 raster2_tile_slow(const Params& p, const uint32_t tile, const uint32_t lane, uint8_t* const sm, uint32_t phase, const uint32_t cnt_raw) {
    using S = R2Smem<P>;
    uint32_t* const lane_sm = reinterpret_cast<uint32_t*>(sm + S::OFF_ITEM_ROW);
    uint16_t* const queue = reinterpret_cast<uint16_t*>(lane_sm + lane * Q_STRIDE_WORDS);
    uint32_t* const col = lane_sm + 32 * Q_STRIDE_WORDS + lane * COL_STRIDE;
    return raster_tile<P, false, false, false, true>(p, tile, lane, reinterpret_cast<uint32_t*>(sm), reinterpret_cast<uint64_t*>(sm + S::OFF_BAR), phase, queue, col, cnt_raw);
}

template <class P>
__global__ void __launch_bounds__(RASTER_WARPS * 32, EUC_R2_MIN_CTAS) raster2_kernel(const __grid_constant__ Params p, uint32_t n_tiles) {
    using S = R2Smem<P>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint8_t* const sm = smem_raw + (size_t)warp * S::BYTES;
    uint64_t* const bar = reinterpret_cast<uint64_t*>(sm + S::OFF_BAR);
    uint32_t* const slow = reinterpret_cast<uint32_t*>(sm + S::OFF_BAR + 16);
    if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); slow[0] = 0u; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t phase = 0, nfrag = 0;
    if (!render_aborted(p)) {
        unsigned int* const ticket = reinterpret_cast<unsigned int*>(p.counters + 4);
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
                uint32_t n;
                if (p.bin_cap) {
                    n = p.tile_count[tile];
                    if (n > p.bin_cap) {  // overflowed, or marked TILE_LOST: later
                        const uint32_t k = slow[0];
                        __syncwarp();
                        if (lane == 0) { slow[1u + k] = tile; slow[0] = k + 1u; }
                        __syncwarp();
                        if (k + 1u == SLOW_SLOTS) { more = true; break; }
                        continue;
                    }
                    __syncwarp();
                    if (lane == 0 && n) p.tile_count[tile] = 0u;  // leave the counters zeroed for the next render
                    if (p.summary && n * 4u > p.bin_cap * 3u && lane == 0) atomicMax(p.counters + 7, (unsigned long long)n);
                } else {
                    n = p.tile_range[tile].y;
                }
                { const uint32_t f = raster2_tile<P>(p, tile, lane, sm, n); if (lane == 0) nfrag += f; }  // every lane holds the same count
                __syncwarp();
            }
            const uint32_t n_slow = slow[0];
            for (uint32_t i = 0; i < n_slow; ++i) {
                const uint32_t tile = slow[1u + i];
                const uint2 res = raster2_tile_slow<P>(p, tile, lane, sm, phase, p.tile_count[tile]);
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
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_tiles; i += gridDim.x * blockDim.x) p.tile_count[i] = 0u;
    }
    if (p.summary && lane == 0) {  // summary for the host, as in raster_kernel
        __threadfence();
        const unsigned long long done = atomicAdd(p.counters + 8, 1ull) + 1ull;
        if (done == (unsigned long long)gridDim.x * RASTER_WARPS) {
            volatile unsigned long long* c = p.counters;
            if (c[3]) atomicOr(p.counters + 15, c[3]);
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

}  // namespace eucb
