// shaders.cuh — sampler device functions (K6) and the shader stages of the benchmarked pipelines.
// f32 operation order follows the Rust shaders of the reference (file:line cited per function); nothing here
// may be contracted (TU is built with --fmad=false), the only fused operations are the explicit __fmaf_rn in
// mat4_mul_vec4.
#pragma once
#include "../../include/euc_b200.h"
#include "rust_f32.cuh"

namespace eucb {

struct SamplerDev {
    const void* data;
    uint32_t w, h;
    int32_t format, filter, wrap;
};

// src/sampler/mod.rs:110-113 (Clamped), :134-137 (Tiled), :159-168 (Mirrored)
__device__ __forceinline__ float wrap_coord(int wrap, float e) {
    switch (wrap) {
        case EUC_WRAP_CLAMP: return r_min(r_max(e, 0.0f), 1.0f);
        case EUC_WRAP_TILE: return r_rem_euclid(e, 1.0f);
        case EUC_WRAP_MIRROR:
            if (r_rem_euclid(e, 2.0f) >= 1.0f) return 1.0f - r_rem_euclid(e, 1.0f);
            return r_rem_euclid(e, 1.0f);
        default: return e;
    }
}

// Texel fetch: Buffer::read_unchecked (src/buffer.rs:176-180, index x + w*y :90), with `Map` applied on read
// (src/texture.rs:169-176).
template <int C> struct Texel;
template <> struct Texel<1> {
    float v[1];
    static __device__ __forceinline__ Texel fetch(const SamplerDev& s, unsigned long long x, unsigned long long y) {
        Texel t;
        t.v[0] = __ldg((const float*)s.data + (x + (unsigned long long)s.w * y));
        return t;
    }
};
template <> struct Texel<4> {
    float v[4];
    static __device__ __forceinline__ Texel fetch(const SamplerDev& s, unsigned long long x, unsigned long long y) {
        uint32_t p = __ldg((const uint32_t*)s.data + (x + (unsigned long long)s.w * y));
        Texel t;  // examples/texture_mapping.rs:119-121: Rgba::from(pixel.0).map(|e: u8| e as f32)
        t.v[0] = (float)(p & 0xffu);
        t.v[1] = (float)((p >> 8) & 0xffu);
        t.v[2] = (float)((p >> 16) & 0xffu);
        t.v[3] = (float)(p >> 24);
        return t;
    }
};

// src/sampler/linear.rs:30-65 and src/sampler/nearest.rs:27-32 (+ Denormalize, src/math.rs:51-53).
template <int C> __device__ __forceinline__ void sample2d(const SamplerDev& s, float x, float y, float* out) {
    x = wrap_coord(s.wrap, x);
    y = wrap_coord(s.wrap, y);
    const unsigned long long w = s.w, h = s.h;
    if (s.filter == EUC_FILTER_LINEAR) {
        float index_tex_x = r_fract(x) * (float)s.w;
        float index_tex_y = r_fract(y) * (float)s.h;
        unsigned long long posi_x = r_as_usize(truncf(index_tex_x));
        unsigned long long posi_y = r_as_usize(truncf(index_tex_y));
        float fract_x = r_fract(index_tex_x);
        float fract_y = r_fract(index_tex_y);
        unsigned long long p0x = min(posi_x + 0ull, w - 1ull), p0y = min(posi_y + 0ull, h - 1ull);
        unsigned long long p1x = min(posi_x + 1ull, w - 1ull), p1y = min(posi_y + 1ull, h - 1ull);
        Texel<C> t00 = Texel<C>::fetch(s, p0x, p0y);
        Texel<C> t10 = Texel<C>::fetch(s, p1x, p0y);
        Texel<C> t01 = Texel<C>::fetch(s, p0x, p1y);
        Texel<C> t11 = Texel<C>::fetch(s, p1x, p1y);
        float omy = 1.0f - fract_y, omx = 1.0f - fract_x;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float t0 = t00.v[c] * omy + t01.v[c] * fract_y;
            float t1 = t10.v[c] * omy + t11.v[c] * fract_y;
            out[c] = t0 * omx + t1 * fract_x;
        }
    } else {
        unsigned long long ix = min(r_as_usize(r_max(x * (float)s.w, 0.0f)), w - 1ull);
        unsigned long long iy = min(r_as_usize(r_max(y * (float)s.h, 0.0f)), h - 1ull);
        Texel<C> t = Texel<C>::fetch(s, ix, iy);
#pragma unroll
        for (int c = 0; c < C; ++c) out[c] = t.v[c];
    }
}

// -------------------------------------------------------------------------------------------------------
// Pipelines.  Each provides:
//   V            number of interpolated f32 varyings (VertexData)
//   HAS_FRAGMENT false when Fragment = Unit (depth-only pipelines never shade)
//   vertex(u, vptr, clip, var)       Pipeline::vertex
//   fragment(u, samp, var, frag)     Pipeline::fragment   (frag = Rgba<f32>)
//   blend(old, frag)                 Pipeline::blend      (Pixel = u32)
//   BLEND_IGNORES_OLD  blend does not read `old` (and fragment is pure): shading may be deferred to once per pixel
// -------------------------------------------------------------------------------------------------------

// benches/teapot.rs:10-51
struct PipeTeapotShadow {
    static constexpr int V = 0;
    static constexpr bool HAS_FRAGMENT = false;
    static constexpr bool BLEND_IGNORES_OLD = false;
    using Uniforms = euc_uniforms_teapot_shadow;
    static constexpr uint32_t VERTEX_BYTES = sizeof(euc_vertex_pn);
    static __device__ __forceinline__ void vertex(const Uniforms& u, const uint8_t* vp, float4& clip, float* var) {
        const float* f = (const float*)vp;  // :37-42
        clip = mat4_mul_vec4(u.mvp, f[0], f[1], f[2], 1.0f);
    }
    static __device__ __forceinline__ void fragment(const Uniforms&, const SamplerDev*, const float*, float* frag) {}
    static __device__ __forceinline__ uint32_t blend(uint32_t old, const float*) { return old; }
};

// benches/teapot.rs:53-142
struct PipeTeapotPhong {
    static constexpr int V = 9;
    static constexpr bool HAS_FRAGMENT = true;
    static constexpr bool BLEND_IGNORES_OLD = true;
    using Uniforms = euc_uniforms_teapot_phong;
    static constexpr uint32_t VERTEX_BYTES = sizeof(euc_vertex_pn);
    static __device__ __forceinline__ void vertex(const Uniforms& u, const uint8_t* vp, float4& clip, float* var) {
        const float* f = (const float*)vp;  // :83-97
        float4 wpos = mat4_mul_vec4(u.m, f[0], f[1], f[2], 1.0f);
        float4 wnorm = mat4_mul_vec4(u.m, -f[3], -f[4], -f[5], 0.0f);
        float4 lvp = mat4_mul_vec4(u.light_vp, wpos.x, wpos.y, wpos.z, 1.0f);
        float4 vw = mat4_mul_vec4(u.v, wpos.x, wpos.y, wpos.z, wpos.w);
        clip = mat4_mul_vec4(u.p, vw.x, vw.y, vw.z, vw.w);
        var[0] = wpos.x; var[1] = wpos.y; var[2] = wpos.z;
        var[3] = wnorm.x; var[4] = wnorm.y; var[5] = wnorm.z;
        var[6] = lvp.x / lvp.w; var[7] = lvp.y / lvp.w; var[8] = lvp.z / lvp.w;
    }
    static __device__ __forceinline__ void fragment(const Uniforms& u, const SamplerDev* samp, const float* d, float* frag) {
        // :100-133
        float nm = sqrtf(dot3(d[3], d[4], d[5], d[3], d[4], d[5]));
        float nx = d[3] / nm, ny = d[4] / nm, nz = d[5] / nm;                      // wnorm.normalized()
        float cx = u.cam_pos[0] - d[0], cy = u.cam_pos[1] - d[1], cz = u.cam_pos[2] - d[2];
        float cm = sqrtf(dot3(cx, cy, cz, cx, cy, cz));
        cx = cx / cm; cy = cy / cm; cz = cz / cm;                                   // cam_dir
        float lx = d[0] - u.light_pos[0], ly = d[1] - u.light_pos[1], lz = d[2] - u.light_pos[2];
        float lm = sqrtf(dot3(lx, ly, lz, lx, ly, lz));
        lx = lx / lm; ly = ly / lm; lz = lz / lm;                                   // light_dir
        float ambient = 0.1f;
        float diffuse = r_max(dot3(nx, ny, nz, -lx, -ly, -lz), 0.0f) * 0.5f;
        // (-light_dir).reflected(wnorm) = v - n * (2 * dot(v, n))
        float vx = -lx, vy = -ly, vz = -lz;
        float p2 = 2.0f * dot3(vx, vy, vz, nx, ny, nz);
        float rx = vx - nx * p2, ry = vy - ny * p2, rz = vz - nz * p2;
        float specular = powf(r_max(dot3(rx, ry, rz, -cx, -cy, -cz), 0.0f), 30.0f) * 3.0f;
        float sx = d[6] * 1.0f * 0.5f + 0.5f;
        float sy = d[7] * -1.0f * 0.5f + 0.5f;
        float tap;
        sample2d<1>(samp[0], sx, sy, &tap);
        float light_depth = tap + 0.0001f;
        bool in_light = d[8] < light_depth;
        float light = ambient + (in_light ? diffuse + specular : 0.0f);
        frag[0] = 1.0f * light; frag[1] = 0.8f * light; frag[2] = 0.7f * light; frag[3] = 1.0f * light;
    }
    static __device__ __forceinline__ uint32_t blend(uint32_t, const float* f) {  // :136-141, BGRA
        uint32_t r = r_as_u8(r_min(r_max(f[0], 0.0f), 1.0f) * 255.0f);
        uint32_t g = r_as_u8(r_min(r_max(f[1], 0.0f), 1.0f) * 255.0f);
        uint32_t b = r_as_u8(r_min(r_max(f[2], 0.0f), 1.0f) * 255.0f);
        uint32_t a = r_as_u8(r_min(r_max(f[3], 0.0f), 1.0f) * 255.0f);
        return pack_le(b, g, r, a);
    }
};

// examples/texture_mapping.rs:5-35
struct PipeTexCube {
    static constexpr int V = 2;
    static constexpr bool HAS_FRAGMENT = true;
    static constexpr bool BLEND_IGNORES_OLD = true;
    using Uniforms = euc_uniforms_tex_cube;
    static constexpr uint32_t VERTEX_BYTES = sizeof(euc_vertex_p4uv);
    static __device__ __forceinline__ void vertex(const Uniforms& u, const uint8_t* vp, float4& clip, float* var) {
        const float4 p = *(const float4*)vp;  // :20-25
        const float2 uv = *(const float2*)(vp + 16);
        clip = mat4_mul_vec4(u.mvp, p.x, p.y, p.z, p.w);
        var[0] = uv.x; var[1] = uv.y;
    }
    static __device__ __forceinline__ void fragment(const Uniforms&, const SamplerDev* samp, const float* uv, float* frag) {
        sample2d<4>(samp[0], uv[0], uv[1], frag);  // :28-30
    }
    static __device__ __forceinline__ uint32_t blend(uint32_t, const float* f) {  // :32-34
        return pack_le(r_as_u8(f[0]), r_as_u8(f[1]), r_as_u8(f[2]), r_as_u8(f[3]));
    }
};

// BASELINE config 4 (defined by this build; DESIGN.md "Synthetic workloads")
struct PipeBlendTris {
    static constexpr int V = 4;
    static constexpr bool HAS_FRAGMENT = true;
    static constexpr bool BLEND_IGNORES_OLD = false;
    struct Uniforms { float _unused[4]; };
    static constexpr uint32_t VERTEX_BYTES = sizeof(euc_vertex_p4c4);
    static __device__ __forceinline__ void vertex(const Uniforms&, const uint8_t* vp, float4& clip, float* var) {
        clip = *(const float4*)vp;
        const float4 c = *(const float4*)(vp + 16);
        var[0] = c.x; var[1] = c.y; var[2] = c.z; var[3] = c.w;
    }
    static __device__ __forceinline__ void fragment(const Uniforms&, const SamplerDev*, const float* v, float* frag) {
        frag[0] = v[0]; frag[1] = v[1]; frag[2] = v[2]; frag[3] = v[3];
    }
    static __device__ __forceinline__ uint32_t blend(uint32_t old, const float* n) {
        float a = n[3], ia = 1.0f - a;
        float c0 = (n[0] * 255.0f) * a + (float)(old & 0xffu) * ia;
        float c1 = (n[1] * 255.0f) * a + (float)((old >> 8) & 0xffu) * ia;
        float c2 = (n[2] * 255.0f) * a + (float)((old >> 16) & 0xffu) * ia;
        return pack_le(r_clamp255_as_u8(c0), r_clamp255_as_u8(c1),
                       r_clamp255_as_u8(c2), 255u);
    }
};

// BASELINE config 5 (defined by this build)
struct PipeVoxelIcon {
    static constexpr int V = 7;
    static constexpr bool HAS_FRAGMENT = true;
    static constexpr bool BLEND_IGNORES_OLD = false;
    using Uniforms = euc_uniforms_voxel_icon;
    static constexpr uint32_t VERTEX_BYTES = sizeof(euc_vertex_voxel);
    static __device__ __forceinline__ void vertex(const Uniforms& u, const uint8_t* vp, float4& clip, float* var) {
        const float4 a = *(const float4*)vp;         // pos.xyz, normal.x
        const float4 b = *(const float4*)(vp + 16);  // normal.yz, rgba(u8x4), pad
        clip = mat4_mul_vec4(u.mvp, a.x, a.y, a.z, 1.0f);
        uint32_t c = __float_as_uint(b.z);
        const float k = 1.0f / 255.0f;
        var[0] = a.w; var[1] = b.x; var[2] = b.y;
        var[3] = (float)(c & 0xffu) * k; var[4] = (float)((c >> 8) & 0xffu) * k;
        var[5] = (float)((c >> 16) & 0xffu) * k; var[6] = (float)(c >> 24) * k;
    }
    static __device__ __forceinline__ void fragment(const Uniforms& u, const SamplerDev*, const float* d, float* frag) {
        float s = 0.35f + 0.65f * r_max(dot3(d[0], d[1], d[2], u.light_dir[0], u.light_dir[1], u.light_dir[2]), 0.0f);
        frag[0] = d[3] * s; frag[1] = d[4] * s; frag[2] = d[5] * s; frag[3] = d[6];
    }
    static __device__ __forceinline__ uint32_t blend(uint32_t old, const float* n) {
        float a = n[3], ia = 1.0f - a;
        float ob = (float)(old & 0xffu), og = (float)((old >> 8) & 0xffu), orr = (float)((old >> 16) & 0xffu), oa = (float)(old >> 24);
        float r = (n[0] * 255.0f) * a + orr * ia;
        float g = (n[1] * 255.0f) * a + og * ia;
        float b = (n[2] * 255.0f) * a + ob * ia;
        float A = a * 255.0f + oa * ia;
        return pack_le(r_clamp255_as_u8(b), r_clamp255_as_u8(g),
                       r_clamp255_as_u8(r), r_clamp255_as_u8(A));
    }
};

// examples/triangle.rs:7-25, examples/spinning_cube.rs:5-29
struct PipeVertexColor {
    static constexpr int V = 4;
    static constexpr bool HAS_FRAGMENT = true;
    static constexpr bool BLEND_IGNORES_OLD = true;
    using Uniforms = euc_uniforms_vertex_color;
    static constexpr uint32_t VERTEX_BYTES = sizeof(euc_vertex_p4c4);
    static __device__ __forceinline__ void vertex(const Uniforms& u, const uint8_t* vp, float4& clip, float* var) {
        const float4 p = *(const float4*)vp;
        const float4 c = *(const float4*)(vp + 16);
        clip = mat4_mul_vec4(u.mvp, p.x, p.y, p.z, p.w);
        var[0] = c.x; var[1] = c.y; var[2] = c.z; var[3] = c.w;
    }
    static __device__ __forceinline__ void fragment(const Uniforms&, const SamplerDev*, const float* v, float* frag) {
        frag[0] = v[0]; frag[1] = v[1]; frag[2] = v[2]; frag[3] = v[3];
    }
    static __device__ __forceinline__ uint32_t blend(uint32_t, const float* f) {  // triangle.rs:23-25
        return pack_le(r_as_u8(f[0] * 255.0f), r_as_u8(f[1] * 255.0f), r_as_u8(f[2] * 255.0f), r_as_u8(f[3] * 255.0f));
    }
};

// examples/wireframes.rs:5-37
struct PipeWireframe {
    static constexpr int V = 0;
    static constexpr bool HAS_FRAGMENT = true;
    static constexpr bool BLEND_IGNORES_OLD = true;
    using Uniforms = euc_uniforms_wireframe;
    static constexpr uint32_t VERTEX_BYTES = sizeof(euc_vertex_pn);
    static __device__ __forceinline__ void vertex(const Uniforms& u, const uint8_t* vp, float4& clip, float*) {
        const float* f = (const float*)vp;  // :19-23
        const float4 wpos = mat4_mul_vec4(u.m, f[0], f[1], f[2], 1.0f);
        const float4 vw = mat4_mul_vec4(u.v, wpos.x, wpos.y, wpos.z, wpos.w);
        clip = mat4_mul_vec4(u.p, vw.x, vw.y, vw.z, vw.w);
    }
    static __device__ __forceinline__ void fragment(const Uniforms&, const SamplerDev*, const float*, float* frag) {
        frag[0] = 1.0f; frag[1] = 0.0f; frag[2] = 0.0f; frag[3] = 1.0f;  // Rgba::red()
    }
    static __device__ __forceinline__ uint32_t blend(uint32_t, const float* f) {  // :31-36: clamped(0,1)*255, BGRA
        uint32_t c[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { float e = f[i]; e = e < 0.0f ? 0.0f : (e > 1.0f ? 1.0f : e); c[i] = r_as_u8(e * 255.0f); }
        return pack_le(c[2], c[1], c[0], c[3]);
    }
};

}  // namespace eucb
