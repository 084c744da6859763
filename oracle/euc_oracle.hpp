// euc_oracle.hpp — CPU restatement of euc's `Pipeline::render` hot path (TEST INFRASTRUCTURE ONLY).
//
// This file is the parity oracle and the CPU baseline for euc_b200.  It is NOT part of the product: only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or
// call it.  The product (euc_b200/) never includes or links anything from this directory.
//
// PARITY UNPINNED: the reference (zesterer/euc v0.6.0) ships no tests, golden vectors or fixtures, and no
// Rust toolchain exists in this environment, so the reference itself cannot be run.  This restatement is
// pinned only by (a) the Rust source text it follows line by line (citations below, paths relative to the
// euc crate root), (b) hand-derivable known answers (README triangle, band table, sampler KATs) in
// tests/test_oracle_kat.py, and (c) an independently written numpy restatement (oracle/np_oracle.py).
// Third-party arithmetic that is not in the reference tree (vek 0.17 Mat4*Vec4 / normalized / reflected,
// wavefront 0.2 vertex order, clipline 0.2 line walking) is restated from its published behaviour; see
// DESIGN.md "Unpinned beliefs".
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math  (contraction MUST be off: rustc never fuses).
#pragma once
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include <array>
#include <algorithm>

namespace euc {

using usize = uint64_t;
using f32x3 = std::array<float, 3>;
using f32x4 = std::array<float, 4>;

// ---------------------------------------------------------------------------------------------------------
// Rust language semantics for the leaf operations the path relies on (SURVEY §8c "semantics checklist").
// ---------------------------------------------------------------------------------------------------------

// `f as usize`: truncate toward zero, saturate, NaN -> 0.
inline usize f32_as_usize(float f) {
    if (!(f == f)) return 0;
    if (f <= 0.0f) return 0;
    if (f >= 18446744073709551616.0f) return UINT64_MAX;
    return (usize)f;
}
// `f as isize`
inline int64_t f32_as_isize(float f) {
    if (!(f == f)) return 0;
    if (f >= 9223372036854775808.0f) return INT64_MAX;
    if (f <= -9223372036854775808.0f) return INT64_MIN;
    return (int64_t)f;
}
// `f as u8`
inline uint8_t f32_as_u8(float f) {
    if (!(f == f)) return 0;
    if (f <= 0.0f) return 0;
    if (f >= 255.0f) return 255;
    return (uint8_t)f;
}
// f32::min / f32::max: IEEE minNum/maxNum — if exactly one operand is NaN the other is returned.
inline float f32_min(float a, float b) { return a < b ? a : (b != b ? a : b); }
inline float f32_max(float a, float b) { return a > b ? a : (b != b ? a : b); }
inline float f32_fract(float x) { return x - std::trunc(x); }  // negative stays negative
inline float f32_rem_euclid(float x, float rhs) {
    float r = std::fmod(x, rhs);
    return r < 0.0f ? r + std::fabs(rhs) : r;
}
inline usize usize_clamp(usize v, usize lo, usize hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline float f32_clamp(float v, float lo, float hi) {  // f32::clamp: NaN stays NaN
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}

// ---------------------------------------------------------------------------------------------------------
// src/math.rs:3-41 — WeightedSum.  Every V below provides operator*(V,float) and operator+(V,V), and the
// blanket impl computes v0*w0 + v1*w1 (+ v2*w2), unfused, left-associated.
// ---------------------------------------------------------------------------------------------------------
struct Unit {};  // src/math.rs:13-21
inline Unit operator*(Unit, float) { return {}; }
inline Unit operator+(Unit, Unit) { return {}; }

template <class V> inline V weighted_sum2(const V& v0, const V& v1, float w0, float w1) { return v0 * w0 + v1 * w1; }
template <class V> inline V weighted_sum3(const V& v0, const V& v1, const V& v2, float w0, float w1, float w2) {
    return v0 * w0 + v1 * w1 + v2 * w2;
}

template <int N> struct VecN {
    float e[N];
    float& operator[](int i) { return e[i]; }
    const float& operator[](int i) const { return e[i]; }
};
template <int N> inline VecN<N> operator*(const VecN<N>& a, float s) {
    VecN<N> r;
    for (int i = 0; i < N; ++i) r.e[i] = a.e[i] * s;
    return r;
}
template <int N> inline VecN<N> operator+(const VecN<N>& a, const VecN<N>& b) {
    VecN<N> r;
    for (int i = 0; i < N; ++i) r.e[i] = a.e[i] + b.e[i];
    return r;
}
using Rgba = VecN<4>;

// ---------------------------------------------------------------------------------------------------------
// src/buffer.rs:19-22, 87-100, 147-149, 183-219 — Buffer2d as a borrowed view over host memory, and
// src/texture.rs:285-319 — Empty.
// ---------------------------------------------------------------------------------------------------------
template <class T> struct Buffer2d {
    T* items = nullptr;
    usize size[2] = {0, 0};
    usize linear_index2(usize x, usize y) const { return y * size[0] + x; }                  // buffer.rs:147-149
    T read_unchecked(usize x, usize y) const { return items[x + size[0] * y]; }              // buffer.rs:176-180, :90
    T read_exclusive_unchecked(usize x, usize y) const { return items[linear_index2(x, y)]; }  // buffer.rs:185-189
    void write_exclusive_unchecked(usize x, usize y, T t) const { items[linear_index2(x, y)] = t; }  // :192-199
    void clear(T t) { for (usize i = 0; i < size[0] * size[1]; ++i) items[i] = t; }          // buffer.rs:213-218
    bool is_empty_target() const { return false; }
};
template <class T> struct Empty {
    usize size[2] = {0, 0};                                                   // texture.rs:300-302
    T read_exclusive_unchecked(usize, usize) const { return T(); }            // texture.rs:312-314
    void write_exclusive_unchecked(usize, usize, T) const {}                  // texture.rs:316-317
};

// ---------------------------------------------------------------------------------------------------------
// Samplers.  src/sampler/linear.rs:30-65, src/sampler/nearest.rs:27-32 + src/math.rs:51-53,
// src/sampler/mod.rs:100-179, src/texture.rs:140-176 (Map).
// ---------------------------------------------------------------------------------------------------------
// Texture adaptor: Buffer2d<f32> read as f32.
struct TexF32 {
    const float* data; usize w, h;
    using Texel = float;
    float read_unchecked(usize x, usize y) const { return data[x + w * y]; }
};
// Texture adaptor: Buffer2d<[u8;4]>.map(|p| Rgba::from(p).map(|e| e as f32))  (examples/texture_mapping.rs:119-121)
struct TexRgba8AsF32 {
    const uint8_t* data; usize w, h;
    using Texel = Rgba;
    Rgba read_unchecked(usize x, usize y) const {
        const uint8_t* p = data + 4 * (x + w * y);
        return Rgba{{(float)p[0], (float)p[1], (float)p[2], (float)p[3]}};
    }
};
inline float texel_mul(float t, float s) { return t * s; }

template <class Tex> struct Linear {  // linear.rs:30-65
    Tex tex;
    using Sample = typename Tex::Texel;
    Sample sample(float x, float y) const {
        usize w = tex.w, h = tex.h;
        float index_tex_x = f32_fract(x) * (float)w;
        float index_tex_y = f32_fract(y) * (float)h;
        usize posi_x = f32_as_usize(std::trunc(index_tex_x));
        usize posi_y = f32_as_usize(std::trunc(index_tex_y));
        float fract_x = f32_fract(index_tex_x);
        float fract_y = f32_fract(index_tex_y);
        usize p0x = std::min(posi_x + 0, w - 1);
        usize p0y = std::min(posi_y + 0, h - 1);
        usize p1x = std::min(posi_x + 1, w - 1);
        usize p1y = std::min(posi_y + 1, h - 1);
        Sample t00 = tex.read_unchecked(p0x, p0y);
        Sample t10 = tex.read_unchecked(p1x, p0y);
        Sample t01 = tex.read_unchecked(p0x, p1y);
        Sample t11 = tex.read_unchecked(p1x, p1y);
        Sample t0 = t00 * (1.0f - fract_y) + t01 * fract_y;
        Sample t1 = t10 * (1.0f - fract_y) + t11 * fract_y;
        Sample t = t0 * (1.0f - fract_x) + t1 * fract_x;
        return t;
    }
};
template <class Tex> struct Nearest {  // nearest.rs:27-32; math.rs:51-53
    Tex tex;
    using Sample = typename Tex::Texel;
    static usize denormalize_to(float self, usize scale) {
        return std::min(f32_as_usize(f32_max(self * (float)scale, 0.0f)), scale - 1);
    }
    Sample sample(float x, float y) const { return tex.read_unchecked(denormalize_to(x, tex.w), denormalize_to(y, tex.h)); }
};
template <class S> struct Clamped {  // sampler/mod.rs:110-113
    S s;
    using Sample = typename S::Sample;
    static float map(float e) { return f32_min(f32_max(e, 0.0f), 1.0f); }
    Sample sample(float x, float y) const { return s.sample(map(x), map(y)); }
};
template <class S> struct Tiled {  // sampler/mod.rs:134-137
    S s;
    using Sample = typename S::Sample;
    static float map(float e) { return f32_rem_euclid(e, 1.0f); }
    Sample sample(float x, float y) const { return s.sample(map(x), map(y)); }
};
template <class S> struct Mirrored {  // sampler/mod.rs:159-168
    S s;
    using Sample = typename S::Sample;
    static float map(float e) {
        if (f32_rem_euclid(e, 2.0f) >= 1.0f) return 1.0f - f32_rem_euclid(e, 1.0f);
        return f32_rem_euclid(e, 1.0f);
    }
    Sample sample(float x, float y) const { return s.sample(map(x), map(y)); }
};
// Run-time composition of the above (the C entry point receives filter/wrap as enums).
template <class Tex> struct DynSampler {
    Tex tex; int filter; int wrap;  // euc_filter, euc_wrap
    using Sample = typename Tex::Texel;
    Sample sample(float x, float y) const {
        switch (wrap) {
            case 1: x = Clamped<Linear<Tex>>::map(x); y = Clamped<Linear<Tex>>::map(y); break;
            case 2: x = Tiled<Linear<Tex>>::map(x); y = Tiled<Linear<Tex>>::map(y); break;
            case 3: x = Mirrored<Linear<Tex>>::map(x); y = Mirrored<Linear<Tex>>::map(y); break;
            default: break;
        }
        if (filter == 1) return Linear<Tex>{tex}.sample(x, y);
        return Nearest<Tex>{tex}.sample(x, y);
    }
};

// ---------------------------------------------------------------------------------------------------------
// src/pipeline.rs:14-163 — modes.
// ---------------------------------------------------------------------------------------------------------
enum class Ordering { Less, Equal, Greater };
struct DepthMode {
    bool has_test; Ordering test; bool write;
    bool uses_depth() const { return has_test || write; }  // pipeline.rs:50-52
};
struct PixelMode { bool write; };
enum class YAxisDirection { Down, Up };
struct CoordinateMode {
    YAxisDirection y_axis_direction; bool has_z_clip; float z_start, z_end;
    bool passes_z_clip(float z) const { return !has_z_clip || (z_start <= z && z <= z_end); }  // pipeline.rs:151-156
};
enum class CullMode { None, Back, Front };

struct RenderStats {
    std::atomic<uint64_t> fragments{0};   // emit_fragment calls
    std::atomic<uint64_t> primitives{0};  // begin_primitive calls of ONE band (= assembled primitives)
};

// ---------------------------------------------------------------------------------------------------------
// src/rasterizer/triangles.rs:309-367 — helpers.
// ---------------------------------------------------------------------------------------------------------
inline f32x3 cross(f32x3 a, f32x3 b) { return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}; }
inline f32x3 sub(f32x3 a, f32x3 b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline f32x3 add(f32x3 a, f32x3 b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
inline float dot(f32x3 a, f32x3 b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline float magnitude_squared(f32x3 v) { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; }
using mat3 = std::array<f32x3, 3>;
inline mat3 matmul(const mat3& a, const mat3& b) {
    mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j];
    return r;
}
inline f32x3 mat3_mul_vec3(const mat3& m, f32x3 v) {
    return {m[0][0] * v[0] + m[0][1] * v[1] + m[0][2] * v[2], m[1][0] * v[0] + m[1][1] * v[1] + m[1][2] * v[2],
            m[2][0] * v[0] + m[2][1] * v[1] + m[2][2] * v[2]};
}
inline float lerp(float a, float b, float t) { return a + t * (b - a); }
inline f32x3 scale3(f32x3 v, float s) { return {v[0] * s, v[1] * s, v[2] * s}; }

// Optional per-stage dump of the band-independent part of triangle setup (used by stage-by-stage parity tests).
struct SetupDump {
    uint32_t culled;
    float w_hom_origin[3], w_hom_dx[3], w_hom_dy[3];
    float z_hom[3];
    float verts_by_y[6];
    uint32_t bounds_min[2], bounds_max[2];  // clamped to the whole target ([0,w] x [0,h])
    uint32_t no_verts_clipped;
};

// ---------------------------------------------------------------------------------------------------------
// src/rasterizer/triangles.rs:15-306 — Triangles::rasterize.
// Blitter B provides: target_size/min/max, begin_primitive, test_fragment, emit_fragment (mod.rs:22-49).
// ---------------------------------------------------------------------------------------------------------
template <class V, class B>
void triangles_rasterize(const std::pair<f32x4, V>* vertices, usize n_vertices, const CoordinateMode& coords,
                         CullMode cull_mode, B& blitter, SetupDump* dump = nullptr) {
    const usize tgt_size[2] = {blitter.tgt_size[0], blitter.tgt_size[1]};
    const usize tgt_min[2] = {blitter.tgt_min[0], blitter.tgt_min[1]};
    const usize tgt_max[2] = {blitter.tgt_max[0], blitter.tgt_max[1]};

    bool has_cull = cull_mode != CullMode::None;                       // :31-35
    float cull_dir = cull_mode == CullMode::Back ? 1.0f : -1.0f;
    float flip[2] = {1.0f, coords.y_axis_direction == YAxisDirection::Down ? 1.0f : -1.0f};  // :37-40
    float size_x = (float)tgt_size[0], size_y = (float)tgt_size[1];   // :42
    mat3 to_ndc = {{{2.0f / size_x, 0.0f, -1.0f}, {0.0f, -2.0f / size_y, 1.0f}, {0.0f, 0.0f, 1.0f}}};  // :44-48

    for (usize t = 0; t + 3 <= n_vertices; t += 3) {  // :50-54
        blitter.begin_primitive();                    // :55
        SetupDump* d = dump ? dump + t / 3 : nullptr;
        if (d) std::memset(d, 0, sizeof(*d));

        f32x4 verts_hom[3] = {vertices[t].first, vertices[t + 1].first, vertices[t + 2].first};  // :58
        const V* verts_out[3] = {&vertices[t].second, &vertices[t + 1].second, &vertices[t + 2].second};
        for (auto& v : verts_hom) v = {v[0] * flip[0], v[1] * flip[1], v[2], v[3]};              // :61
        f32x3 verts_euc[3];
        for (int i = 0; i < 3; ++i)
            verts_euc[i] = {verts_hom[i][0] / verts_hom[i][3], verts_hom[i][1] / verts_hom[i][3], verts_hom[i][2] / verts_hom[i][3]};  // :64

        float winding = cross(sub(verts_euc[1], verts_euc[0]), sub(verts_euc[2], verts_euc[0]))[2];  // :67-70
        if (has_cull && winding * cull_dir < 0.0f) {  // :73-77
            if (d) d->culled = 1;
            continue;
        } else if (winding >= 0.0f) {                 // :78-80 reverse vertex order
            std::swap(verts_hom[0], verts_hom[2]);
            std::swap(verts_euc[0], verts_euc[2]);
            std::swap(verts_out[0], verts_out[2]);
        }

        mat3 coords_to_weights;  // :86-102
        {
            f32x4 a = verts_hom[0], b = verts_hom[1], c4 = verts_hom[2];
            f32x3 c = {c4[0], c4[1], c4[3]};
            f32x3 ca = sub({a[0], a[1], a[3]}, c);
            f32x3 cb = sub({b[0], b[1], b[3]}, c);
            f32x3 n = cross(ca, cb);
            float rec_det = magnitude_squared(n) > 0.0f ? 1.0f / f32_min(dot(n, c), -1.1920929e-07f) : 1.0f;  // EPSILON
            mat3 m = {{scale3(cross(cb, c), rec_det), scale3(cross(c, ca), rec_det), scale3(n, rec_det)}};
            coords_to_weights = matmul(m, to_ndc);
        }

        float verts_screen[3][2];  // :110-111
        for (int i = 0; i < 3; ++i) {
            verts_screen[i][0] = size_x * (verts_euc[i][0] * 0.5f + 0.5f);
            verts_screen[i][1] = size_y * (verts_euc[i][1] * -0.5f + 0.5f);
        }

        usize bounds_clamped_min[2], bounds_clamped_max[2];  // :114-139
        for (int k = 0; k < 2; ++k) {
            float mn = f32_min(f32_min(verts_screen[0][k], verts_screen[1][k]), verts_screen[2][k]) + 0.0f;
            float mx = f32_max(f32_max(verts_screen[0][k], verts_screen[1][k]), verts_screen[2][k]) + 1.0f;
            bounds_clamped_min[k] = usize_clamp(f32_as_usize(mn), tgt_min[k], tgt_max[k]);
            bounds_clamped_max[k] = usize_clamp(f32_as_usize(mx), tgt_min[k], tgt_max[k]);
        }

        auto weights_at = [&](float p0, float p1) { return mat3_mul_vec3(coords_to_weights, {p0, p1, 1.0f}); };  // :142
        f32x3 w_hom_origin = weights_at(0.0f, 0.0f);
        f32x3 w_hom_dx = scale3(sub(weights_at(1000.0f, 0.0f), w_hom_origin), 1.0f / 1000.0f);  // :144
        f32x3 w_hom_dy = scale3(sub(weights_at(0.0f, 1000.0f), w_hom_origin), 1.0f / 1000.0f);  // :145

        float min_y = f32_min(f32_min(verts_screen[0][1], verts_screen[1][1]), verts_screen[2][1]);  // :148-151
        int ord[3];
        if (verts_screen[0][1] == min_y) {                       // :152-171
            if (verts_screen[1][1] < verts_screen[2][1]) { ord[0] = 0; ord[1] = 1; ord[2] = 2; }
            else { ord[0] = 0; ord[1] = 2; ord[2] = 1; }
        } else if (verts_screen[1][1] == min_y) {
            if (verts_screen[0][1] < verts_screen[2][1]) { ord[0] = 1; ord[1] = 0; ord[2] = 2; }
            else { ord[0] = 1; ord[1] = 2; ord[2] = 0; }
        } else {
            if (verts_screen[0][1] < verts_screen[1][1]) { ord[0] = 2; ord[1] = 0; ord[2] = 1; }
            else { ord[0] = 2; ord[1] = 1; ord[2] = 0; }
        }
        float verts_by_y[3][2];
        for (int i = 0; i < 3; ++i) { verts_by_y[i][0] = verts_screen[ord[i]][0]; verts_by_y[i][1] = verts_screen[ord[i]][1]; }

        bool no_verts_clipped = coords.passes_z_clip(verts_euc[0][2]) && coords.passes_z_clip(verts_euc[1][2]) &&
                                coords.passes_z_clip(verts_euc[2][2]);  // :173

        if (d) {
            for (int i = 0; i < 3; ++i) { d->w_hom_origin[i] = w_hom_origin[i]; d->w_hom_dx[i] = w_hom_dx[i]; d->w_hom_dy[i] = w_hom_dy[i]; d->z_hom[i] = verts_hom[i][2]; }
            for (int i = 0; i < 3; ++i) { d->verts_by_y[2 * i] = verts_by_y[i][0]; d->verts_by_y[2 * i + 1] = verts_by_y[i][1]; }
            for (int k = 0; k < 2; ++k) {  // band-independent form: clamp to the whole target instead of the band
                float mn = f32_min(f32_min(verts_screen[0][k], verts_screen[1][k]), verts_screen[2][k]) + 0.0f;
                float mx = f32_max(f32_max(verts_screen[0][k], verts_screen[1][k]), verts_screen[2][k]) + 1.0f;
                d->bounds_min[k] = (uint32_t)usize_clamp(f32_as_usize(mn), 0, tgt_size[k]);
                d->bounds_max[k] = (uint32_t)usize_clamp(f32_as_usize(mx), 0, tgt_size[k]);
            }
            d->no_verts_clipped = no_verts_clipped;
        }

        // inner rasterize::<_, _, NO_VERTS_CLIPPED>  :203-304
        for (usize y = bounds_clamped_min[1]; y < bounds_clamped_max[1]; ++y) {  // :219
            usize extent[2] = {bounds_clamped_max[0] - bounds_clamped_min[0], bounds_clamped_max[1] - bounds_clamped_min[1]};
            usize row_range[2];
            if (extent[0] * extent[1] < 128) {  // :224-226
                row_range[0] = bounds_clamped_min[0];
                row_range[1] = bounds_clamped_max[0];
            } else {
                const float* a = verts_by_y[0]; const float* b = verts_by_y[1]; const float* c = verts_by_y[2];
                float yf = (float)y;
                float ac = lerp(a[0], c[0], (yf - a[1]) / (c[1] - a[1]));  // :231
                float row_bounds[2];
                if (yf < b[1]) {                                              // :233-239
                    float ab = lerp(a[0], b[0], (yf - a[1]) / (b[1] - a[1]));
                    row_bounds[0] = f32_min(ab, ac); row_bounds[1] = f32_max(ab, ac);
                } else {
                    float bc = lerp(b[0], c[0], (yf - b[1]) / (c[1] - b[1]));
                    row_bounds[0] = f32_min(bc, ac); row_bounds[1] = f32_max(bc, ac);
                }
                auto screen_clamp = [&](float e, usize bnd) {               // :242-249
                    if (e >= (float)bounds_clamped_min[0] && e < (float)bounds_clamped_max[0]) return f32_as_usize(e);
                    return bnd;
                };
                row_range[0] = screen_clamp(std::floor(row_bounds[0]), bounds_clamped_min[0]);  // :251
                row_range[1] = screen_clamp(std::ceil(row_bounds[1]), bounds_clamped_max[0]);   // :252
            }

            f32x3 w_hom = add(add(w_hom_origin, scale3(w_hom_dy, (float)y)), scale3(w_hom_dx, (float)row_range[0]));  // :257-260

            for (usize x = row_range[0]; x < row_range[1]; ++x) {  // :262
                f32x3 w_unbalanced = {w_hom[0], w_hom[1], w_hom[2] - w_hom[0] - w_hom[1]};  // :264
                if (w_unbalanced[0] >= 0.0f && w_unbalanced[1] >= 0.0f && w_unbalanced[2] >= 0.0f) {  // :267
                    float z = dot({verts_hom[0][2], verts_hom[1][2], verts_hom[2][2]}, w_unbalanced);  // :269
                    if ((no_verts_clipped || coords.passes_z_clip(z)) && blitter.test_fragment(x, y, z)) {  // :271-272
                        auto get_v_data = [&](float fx, float fy) {  // :274-294
                            f32x3 wh = add(add(w_hom_origin, scale3(w_hom_dy, fy)), scale3(w_hom_dx, fx));
                            f32x3 wu = {wh[0], wh[1], wh[2] - wh[0] - wh[1]};
                            float r = 1.0f / wh[2];  // recip()
                            return weighted_sum3(*verts_out[0], *verts_out[1], *verts_out[2], wu[0] * r, wu[1] * r, wu[2] * r);
                        };
                        blitter.emit_fragment(x, y, get_v_data, z);  // :296
                    }
                }
                w_hom = add(w_hom, w_hom_dx);  // :301
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// src/rasterizer/lines.rs:12-120 — Lines::rasterize.  clipline 0.2 is not in the reference tree; the walk
// below restates its documented behaviour: Bresenham from (x1,y1) to (x2,y2) INCLUSIVE of both endpoints,
// visiting only points inside the inclusive window, in order from the first endpoint.  (PARITY UNPINNED.)
// ---------------------------------------------------------------------------------------------------------
// Closed form of the walk: pixel i along the major axis (0 <= i <= dmajor) has minor offset
// k_i = 0 for i == 0, else floor((2*dminor*i + dmajor - 1) / (2*dmajor)), i.e. the Bresenham error recurrence
// `err = 2*dminor - dmajor; if (err > 0) { minor += s; err -= 2*dmajor; } err += 2*dminor;` (ties stay).
// Only the pixels inside the inclusive window are visited (in walk order), so far-away end points cost nothing.
// Segments with a coordinate beyond +-2^62 (saturated casts of garbage vertices) are skipped: their differences
// overflow i64 and the behaviour of the real crate is unknowable here.
inline int64_t bres_minor(int64_t i, int64_t dminor, int64_t dmajor) {
    if (i == 0) return 0;
    return (int64_t)(((__int128)2 * dminor * i + dmajor - 1) / ((__int128)2 * dmajor));
}
template <class F> inline void clipline_walk(int64_t x1, int64_t y1, int64_t x2, int64_t y2, int64_t wx1, int64_t wy1,
                                             int64_t wx2, int64_t wy2, F&& f) {
    if (wx1 > wx2 || wy1 > wy2) return;
    const int64_t LIM = (int64_t)1 << 62;
    if (x1 <= -LIM || x1 >= LIM || x2 <= -LIM || x2 >= LIM || y1 <= -LIM || y1 >= LIM || y2 <= -LIM || y2 >= LIM) return;
    const __int128 dx128 = x2 > x1 ? (__int128)x2 - x1 : (__int128)x1 - x2, dy128 = y2 > y1 ? (__int128)y2 - y1 : (__int128)y1 - y2;
    if (dx128 >= ((__int128)1 << 62) || dy128 >= ((__int128)1 << 62)) return;
    const int64_t dx = (int64_t)dx128, dy = (int64_t)dy128;
    const int64_t sx = x1 < x2 ? 1 : -1, sy = y1 < y2 ? 1 : -1;
    if (dx >= dy) {
        // window columns in walk order
        const int64_t lo = std::max(wx1, std::min(x1, x2)), hi = std::min(wx2, std::max(x1, x2));
        for (int64_t c = 0; c <= hi - lo; ++c) {
            const int64_t x = sx > 0 ? lo + c : hi - c;
            const int64_t i = (x - x1) * sx;
            const int64_t y = y1 + sy * bres_minor(i, dy, dx);
            if (y >= wy1 && y <= wy2) f(x, y);
        }
    } else {
        const int64_t lo = std::max(wy1, std::min(y1, y2)), hi = std::min(wy2, std::max(y1, y2));
        for (int64_t c = 0; c <= hi - lo; ++c) {
            const int64_t y = sy > 0 ? lo + c : hi - c;
            const int64_t i = (y - y1) * sy;
            const int64_t x = x1 + sx * bres_minor(i, dx, dy);
            if (x >= wx1 && x <= wx2) f(x, y);
        }
    }
}

template <class V, class B>
void lines_rasterize(const std::pair<f32x4, V>* vertices, usize n_vertices, const CoordinateMode& coords, B& blitter) {
    const usize* tgt_size = blitter.tgt_size; const usize* tgt_min = blitter.tgt_min; const usize* tgt_max = blitter.tgt_max;
    float flip[2] = {1.0f, coords.y_axis_direction == YAxisDirection::Down ? 1.0f : -1.0f};  // lines.rs:28-31
    float size[2] = {(float)tgt_size[0], (float)tgt_size[1]};
    for (usize t = 0; t + 2 <= n_vertices; t += 2) {  // :35-37
        blitter.begin_primitive();                    // :38
        f32x4 verts_hom[2] = {vertices[t].first, vertices[t + 1].first};
        const V* verts_out[2] = {&vertices[t].second, &vertices[t + 1].second};
        for (auto& v : verts_hom) v = {v[0] * flip[0], v[1] * flip[1], v[2], v[3]};  // :44
        f32x3 verts_euc[2];
        for (int i = 0; i < 2; ++i) {  // :47-50
            float w = f32_max(verts_hom[i][3], 0.0001f);
            verts_euc[i] = {verts_hom[i][0] / w, verts_hom[i][1] / w, verts_hom[i][2] / w};
        }
        float verts_screen[2][2];  // :53-54
        for (int i = 0; i < 2; ++i) {
            verts_screen[i][0] = size[0] * (verts_euc[i][0] * 0.5f + 0.5f);
            verts_screen[i][1] = size[1] * (verts_euc[i][1] * -0.5f + 0.5f);
        }
        float screen_min[2] = {(float)tgt_min[0], (float)tgt_min[1]};  // :57-58
        float screen_max[2] = {(float)tgt_max[0], (float)tgt_max[1]};
        int64_t x1 = f32_as_isize(verts_screen[0][0]), y1 = f32_as_isize(verts_screen[0][1]);  // :60-61
        int64_t x2 = f32_as_isize(verts_screen[1][0]), y2 = f32_as_isize(verts_screen[1][1]);
        int64_t wx1 = f32_as_isize(f32_clamp(f32_min(verts_screen[0][0], verts_screen[1][0]) + 0.0f, screen_min[0], screen_max[0]));  // :63-68
        int64_t wy1 = f32_as_isize(f32_clamp(f32_min(verts_screen[0][1], verts_screen[1][1]) + 0.0f, screen_min[1], screen_max[1]));
        int64_t wx2 = f32_as_isize(f32_clamp(f32_max(verts_screen[0][0], verts_screen[1][0]) + 1.0f, screen_min[0], screen_max[0]));  // :69-74
        int64_t wy2 = f32_as_isize(f32_clamp(f32_max(verts_screen[0][1], verts_screen[1][1]) + 1.0f, screen_min[1], screen_max[1]));
        // (x1 - x2).abs() > (y1 - y2).abs(); overflow-checks are off in the reference profile -> wrapping
        auto wabs = [](int64_t a, int64_t b) { int64_t d = (int64_t)((uint64_t)a - (uint64_t)b); return d < 0 ? (int64_t)(0 - (uint64_t)d) : d; };
        bool use_x = wabs(x1, x2) > wabs(y1, y2);  // :76
        float norm = 1.0f / (use_x ? verts_screen[1][0] - verts_screen[0][0] : verts_screen[1][1] - verts_screen[0][1]);  // :77-82
        clipline_walk(x1, y1, x2, y2, wx1, wy1, wx2 - 1, wy2 - 1, [&](int64_t xi, int64_t yi) {  // :84-87
            usize x = (usize)xi, y = (usize)yi;
            float frac = (use_x ? (float)x - verts_screen[0][0] : (float)y - verts_screen[0][1]) * norm;  // :90-94
            float z = verts_euc[0][2] + frac * (verts_euc[1][2] - verts_euc[0][2]);                       // :97
            if (coords.passes_z_clip(z) && blitter.test_fragment(x, y, z)) {                              // :99
                auto get_v_data = [&](float fx, float fy) {                                               // :100-113
                    float fr = (use_x ? fx - verts_screen[0][0] : fy - verts_screen[0][1]) * norm;
                    return weighted_sum2(*verts_out[0], *verts_out[1], 1.0f - fr, fr);
                };
                blitter.emit_fragment(x, y, get_v_data, z);  // :115
            }
        });
    }
}

// ---------------------------------------------------------------------------------------------------------
// src/pipeline.rs:396-614 — render_inner with BlitterImpl.
// Pipe provides: using VertexData, Fragment, Pixel; fragment(VertexData)->Fragment; blend(Pixel, Fragment)->Pixel.
// ---------------------------------------------------------------------------------------------------------
enum class PrimKind { TriangleList, LineList, LineTriangleList };

template <class Pipe, class P, class D> struct BlitterImpl {  // pipeline.rs:452-468
    using VD = typename Pipe::VertexData;
    using Frag = typename Pipe::Fragment;
    bool write_pixels; DepthMode depth_mode;
    usize tgt_min[2], tgt_max[2], tgt_size[2];
    const Pipe* pipeline; const P* pixel; const D* depth;
    uint64_t primitive_count = 0;
    usize msaa_level; float msaa_div;
    struct Cell { uint64_t tag; Frag frag; };
    std::vector<Cell> msaa_buf; usize msaa_w = 0;
    uint64_t fragments = 0;

    void begin_primitive() { primitive_count += 1; }  // :514-516 (wrapping_add)

    bool test_fragment(usize x, usize y, float z) {   // :519-526
        if (depth_mode.has_test) {
            float old_z = depth->read_exclusive_unchecked(x, y);
            switch (depth_mode.test) {  // z.partial_cmp(&old_z) == Some(test); NaN compares as None
                case Ordering::Less: return z < old_z;
                case Ordering::Equal: return z == old_z;
                default: return z > old_z;
            }
        }
        return true;
    }

    template <class G> Frag msaa_fragment(usize x, usize y, G& get_v_data) {  // :477-494
        Cell& texel = msaa_buf[(x + 1) + msaa_w * (y + 1)];
        if (texel.tag != primitive_count) {
            texel.tag = primitive_count;
            texel.frag = pipeline->fragment(get_v_data(x, y));
        }
        return texel.frag;
    }

    template <class G> void emit_fragment(usize x, usize y, G&& get_v_data_f, float z) {  // :529-578
        ++fragments;
        if (depth_mode.write) depth->write_exclusive_unchecked(x, y, z);  // :536-538
        if (write_pixels) {                                                // :540
            Frag frag;
            if (msaa_level == 0) {
                frag = pipeline->fragment(get_v_data_f((float)x, (float)y));  // :542
            } else {
                float fractx = f32_fract((float)(x - tgt_min[0]) * msaa_div);  // :544-547
                float fracty = f32_fract((float)(y - tgt_min[1]) * msaa_div);
                usize posix = (x - tgt_min[0]) >> msaa_level;                  // :549-550
                usize posiy = (y - tgt_min[1]) >> msaa_level;
                auto get_v_data = [&](usize cx, usize cy) {                    // :554-559
                    return get_v_data_f((float)(tgt_min[0] + (cx << msaa_level)), (float)(tgt_min[1] + (cy << msaa_level)));
                };
                Frag t00 = msaa_fragment(posix + 0, posiy + 0, get_v_data);   // :561-564
                Frag t10 = msaa_fragment(posix + 1, posiy + 0, get_v_data);
                Frag t01 = msaa_fragment(posix + 0, posiy + 1, get_v_data);
                Frag t11 = msaa_fragment(posix + 1, posiy + 1, get_v_data);
                Frag t0 = weighted_sum2(t00, t01, 1.0f - fracty, fracty);     // :566-567
                Frag t1 = weighted_sum2(t10, t11, 1.0f - fracty, fracty);
                frag = weighted_sum2(t0, t1, 1.0f - fractx, fractx);          // :569
            }
            auto old_px = pixel->read_exclusive_unchecked(x, y);              // :574
            auto blended_px = pipeline->blend(old_px, frag);                  // :575
            pixel->write_exclusive_unchecked(x, y, blended_px);               // :576
        }
    }
};

template <class Pipe, class P, class D>
uint64_t render_inner(const Pipe& pipeline, const std::vector<std::pair<f32x4, typename Pipe::VertexData>>& verts,
                      PrimKind kind, const usize tgt_min[2], const usize tgt_max[2], const usize tgt_size[2], const P& pixel,
                      const D& depth, usize msaa_level, PixelMode pixel_mode, DepthMode depth_mode, const CoordinateMode& coords,
                      CullMode cull, SetupDump* dump, uint64_t* prims) {
    BlitterImpl<Pipe, P, D> b;
    b.write_pixels = pixel_mode.write; b.depth_mode = depth_mode;
    for (int i = 0; i < 2; ++i) { b.tgt_min[i] = tgt_min[i]; b.tgt_max[i] = tgt_max[i]; b.tgt_size[i] = tgt_size[i]; }
    b.pipeline = &pipeline; b.pixel = &pixel; b.depth = &depth;
    b.msaa_level = msaa_level; b.msaa_div = 1.0f / (float)(1u << msaa_level);  // :611
    if (msaa_level > 0) {                                                        // :600-610
        b.msaa_w = ((tgt_max[0] - tgt_min[0]) >> msaa_level) + 3;
        usize mh = ((tgt_max[1] - tgt_min[1]) >> msaa_level) + 3;
        b.msaa_buf.assign(b.msaa_w * mh, {UINT64_MAX, typename Pipe::Fragment{}});
    }
    if (kind == PrimKind::TriangleList)
        triangles_rasterize<typename Pipe::VertexData>(verts.data(), verts.size(), coords, cull, b, dump);  // :581
    else
        lines_rasterize<typename Pipe::VertexData>(verts.data(), verts.size(), coords, b);
    if (prims) *prims = b.primitive_count;
    return b.fragments;
}

// ---------------------------------------------------------------------------------------------------------
// src/pipeline.rs:248-366 — render + render_par.  `shaded` = the collected vertex-stage output
// (`fetch_vertex.collect()` :322), already primitive-assembled (src/primitives.rs:28-43, :56-76, :89-103).
// n_threads == 0 means `available_parallelism()`.
// ---------------------------------------------------------------------------------------------------------
template <class Pipe, class P, class D>
void render_par(const Pipe& pipeline, const std::vector<std::pair<f32x4, typename Pipe::VertexData>>& vertices, PrimKind kind,
                const usize tgt_size[2], const P& pixel, const D& depth, usize msaa_level, PixelMode pm, DepthMode dm,
                const CoordinateMode& coords, CullMode cull, unsigned n_threads, RenderStats* stats, SetupDump* dump,
                usize row_begin = 0, usize row_end = ~(usize)0) {
    usize threads = n_threads ? n_threads : std::max(1u, std::thread::hardware_concurrency());  // :323-325
    const usize FRAGMENTS_PER_GROUP = 20000;                                                     // :328
    usize group_rows = FRAGMENTS_PER_GROUP * ((usize)1 << msaa_level) / std::max<usize>(tgt_size[0], 1);  // :329
    if (group_rows == 0) return;  // reference: division by zero panic at :330 (width > 20000·2^msaa)
    usize needed_threads = std::min(tgt_size[1] / group_rows, threads);                         // :330
    std::atomic<usize> row{0};                                                                   // :326
    std::atomic<bool> dumped{false};
    auto worker = [&]() {
        for (;;) {  // :340-362
            usize row_start = row.fetch_add(group_rows, std::memory_order_relaxed);
            if (row_start >= tgt_size[1]) break;
            usize row_stop = std::min(row_start + group_rows, tgt_size[1]);
            // Row-restricted variant (not in the reference): skip bands outside [row_begin,row_end). Bands are
            // independent, so the rows produced are identical to those of a full render.
            if (row_stop <= row_begin || row_start >= row_end) continue;
            usize tgt_min[2] = {0, row_start}, tgt_max[2] = {tgt_size[0], row_stop};
            bool want_dump = dump && !dumped.exchange(true);
            uint64_t prims = 0;
            uint64_t frags = render_inner(pipeline, vertices, kind, tgt_min, tgt_max, tgt_size, pixel, depth, msaa_level, pm, dm,
                                          coords, cull, want_dump ? dump : nullptr, &prims);
            if (stats) { stats->fragments.fetch_add(frags, std::memory_order_relaxed); stats->primitives.store(prims, std::memory_order_relaxed); }
        }
    };
    if (needed_threads <= 1) {
        if (needed_threads == 1) worker();  // needed_threads == 0  =>  nothing is rendered (:337)
        return;
    }
    std::vector<std::thread> pool;  // :336-339 "Respawning them each time is dumb"
    for (usize i = 0; i < needed_threads; ++i) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
}

}  // namespace euc
