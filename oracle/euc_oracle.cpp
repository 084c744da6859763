// euc_oracle.cpp — the benchmarked pipelines' shader stages restated for the CPU oracle, and the C entry
// points tests/bench use through ctypes.  TEST INFRASTRUCTURE ONLY (see euc_oracle.hpp header).
//
// Shader-stage arithmetic follows benches/teapot.rs:10-142, examples/texture_mapping.rs:5-35,
// examples/triangle.rs:7-25 / examples/spinning_cube.rs:5-29 of the reference.  vek 0.17 (not in the
// reference tree) is restated from memory of its source — UNPINNED beliefs, kept behind EUC_VEK_FMA:
//   Mat4<f32> is column-major; Mat4 * Vec4 = cols[0]*v.x, then mul_add(cols[i], v[i], acc) for i=1..3 (fused);
//   dot = (a*b).sum() left-associated; normalized = v / sqrt(dot(v,v)); reflected(n) = v - n*(2*dot(v,n));
//   Rgba * f32 scales all four channels; as_() is Rust `as` (saturating truncation).
#include "euc_oracle.hpp"
#include "../include/euc_b200.h"
#include <chrono>

#ifndef EUC_VEK_FMA
#define EUC_VEK_FMA 1
#endif

namespace euc {

struct Mat4 { float c[4][4]; };  // c[col][row], column-major like vek
inline Mat4 load_mat4(const float* p) { Mat4 m; std::memcpy(m.c, p, 64); return m; }
inline f32x4 mat4_mul_vec4(const Mat4& m, f32x4 v) {
    f32x4 out;
    for (int r = 0; r < 4; ++r) {
#if EUC_VEK_FMA
        float acc = m.c[0][r] * v[0];
        acc = std::fma(m.c[1][r], v[1], acc);
        acc = std::fma(m.c[2][r], v[2], acc);
        acc = std::fma(m.c[3][r], v[3], acc);
#else
        float acc = m.c[0][r] * v[0] + m.c[1][r] * v[1] + m.c[2][r] * v[2] + m.c[3][r] * v[3];
#endif
        out[r] = acc;
    }
    return out;
}
inline float dot3(f32x3 a, f32x3 b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline f32x3 normalized(f32x3 v) { float m = std::sqrt(dot3(v, v)); return {v[0] / m, v[1] / m, v[2] / m}; }
inline f32x3 neg3(f32x3 v) { return {-v[0], -v[1], -v[2]}; }
inline f32x3 reflected(f32x3 v, f32x3 n) { float p = 2.0f * dot3(v, n); return {v[0] - n[0] * p, v[1] - n[1] * p, v[2] - n[2] * p}; }

inline uint32_t pack_le(uint8_t b0, uint8_t b1, uint8_t b2, uint8_t b3) {
    return (uint32_t)b0 | ((uint32_t)b1 << 8) | ((uint32_t)b2 << 16) | ((uint32_t)b3 << 24);
}

using F32Sampler = DynSampler<TexF32>;
using RgbaSampler = DynSampler<TexRgba8AsF32>;

// ---- benches/teapot.rs:10-51 ---------------------------------------------------------------------------
struct TeapotShadow {
    using Vertex = euc_vertex_pn; using VertexData = float; using Fragment = Unit; using Pixel = Unit;
    Mat4 mvp;
    std::pair<f32x4, float> vertex(const Vertex& v) const {  // :37-42
        return {mat4_mul_vec4(mvp, {v.pos[0], v.pos[1], v.pos[2], 1.0f}), 0.0f};
    }
    Unit fragment(float) const { return {}; }
    Unit blend(Unit, Unit) const { return {}; }
};

// ---- benches/teapot.rs:53-142 --------------------------------------------------------------------------
struct Teapot {
    using Vertex = euc_vertex_pn; using VertexData = VecN<9>; using Fragment = Rgba; using Pixel = uint32_t;
    Mat4 m, v, p, light_vp; f32x3 light_pos, cam_pos; F32Sampler shadow;
    std::pair<f32x4, VertexData> vertex(const Vertex& vx) const {  // :83-97
        f32x4 wpos = mat4_mul_vec4(m, {vx.pos[0], vx.pos[1], vx.pos[2], 1.0f});
        f32x4 wnorm = mat4_mul_vec4(m, {-vx.normal[0], -vx.normal[1], -vx.normal[2], 0.0f});
        f32x4 lvp = mat4_mul_vec4(light_vp, {wpos[0], wpos[1], wpos[2], 1.0f});
        f32x3 light_view_pos = {lvp[0] / lvp[3], lvp[1] / lvp[3], lvp[2] / lvp[3]};
        f32x4 clip = mat4_mul_vec4(p, mat4_mul_vec4(v, wpos));
        VertexData d{{wpos[0], wpos[1], wpos[2], wnorm[0], wnorm[1], wnorm[2], light_view_pos[0], light_view_pos[1], light_view_pos[2]}};
        return {clip, d};
    }
    Rgba fragment(const VertexData& d) const {  // :100-133
        f32x3 wpos = {d[0], d[1], d[2]};
        f32x3 wnorm = normalized({d[3], d[4], d[5]});
        f32x3 lvpos = {d[6], d[7], d[8]};
        f32x3 cam_dir = normalized({cam_pos[0] - wpos[0], cam_pos[1] - wpos[1], cam_pos[2] - wpos[2]});
        f32x3 light_dir = normalized({wpos[0] - light_pos[0], wpos[1] - light_pos[1], wpos[2] - light_pos[2]});
        const float surf_color[4] = {1.0f, 0.8f, 0.7f, 1.0f};
        float ambient = 0.1f;
        float diffuse = f32_max(dot3(wnorm, neg3(light_dir)), 0.0f) * 0.5f;
        float specular = std::pow(f32_max(dot3(reflected(neg3(light_dir), wnorm), neg3(cam_dir)), 0.0f), 30.0f) * 3.0f;
        float sx = lvpos[0] * 1.0f * 0.5f + 0.5f;
        float sy = lvpos[1] * -1.0f * 0.5f + 0.5f;
        float light_depth = shadow.sample(sx, sy) + 0.0001f;
        float depth = lvpos[2];
        bool in_light = depth < light_depth;
        float light = ambient + (in_light ? diffuse + specular : 0.0f);
        return Rgba{{surf_color[0] * light, surf_color[1] * light, surf_color[2] * light, surf_color[3] * light}};
    }
    uint32_t blend(uint32_t, const Rgba& rgba) const {  // :136-141  BGRA
        uint8_t c[4];
        for (int i = 0; i < 4; ++i) c[i] = f32_as_u8(f32_min(f32_max(rgba[i], 0.0f), 1.0f) * 255.0f);
        return pack_le(c[2], c[1], c[0], c[3]);
    }
};

// ---- examples/texture_mapping.rs:5-35 ------------------------------------------------------------------
struct TexCube {
    using Vertex = euc_vertex_p4uv; using VertexData = VecN<2>; using Fragment = Rgba; using Pixel = uint32_t;
    Mat4 mvp; RgbaSampler sampler;
    std::pair<f32x4, VertexData> vertex(const Vertex& v) const {  // :20-25
        return {mat4_mul_vec4(mvp, {v.pos[0], v.pos[1], v.pos[2], v.pos[3]}), VertexData{{v.uv[0], v.uv[1]}}};
    }
    Rgba fragment(const VertexData& uv) const { return sampler.sample(uv[0], uv[1]); }  // :28-30
    uint32_t blend(uint32_t, const Rgba& c) const {                                       // :32-34
        return pack_le(f32_as_u8(c[0]), f32_as_u8(c[1]), f32_as_u8(c[2]), f32_as_u8(c[3]));
    }
};

// ---- BASELINE config 4 (SURVEY §8d "C4"): defined by this build, not by the reference -----------------
// vertex: clip = pos, varying = rgba.  fragment = interpolated rgba.
// blend: o = unpack_u8(old) as f32 (RGBA little-endian); c = (new.c*255)*a + o.c*(1-a), unfused; alpha = 255;
//        pack `max(0).min(255) as u8` -> 0xAABBGGRR.
struct BlendTris {
    using Vertex = euc_vertex_p4c4; using VertexData = Rgba; using Fragment = Rgba; using Pixel = uint32_t;
    std::pair<f32x4, Rgba> vertex(const Vertex& v) const {
        return {{v.pos[0], v.pos[1], v.pos[2], v.pos[3]}, Rgba{{v.rgba[0], v.rgba[1], v.rgba[2], v.rgba[3]}}};
    }
    Rgba fragment(const Rgba& c) const { return c; }
    uint32_t blend(uint32_t old, const Rgba& n) const {
        float a = n[3], ia = 1.0f - a;
        uint8_t out[3];
        for (int i = 0; i < 3; ++i) {
            float o = (float)((old >> (8 * i)) & 0xffu);
            float c = (n[i] * 255.0f) * a + o * ia;
            out[i] = f32_as_u8(f32_min(f32_max(c, 0.0f), 255.0f));
        }
        return pack_le(out[0], out[1], out[2], 255);
    }
};

// ---- BASELINE config 5 (SURVEY §8d "C5"): defined by this build ----------------------------------------
// vertex: clip = mvp*(p,1); varyings = normal xyz, rgba = u8 as f32 * (1/255).
// fragment: s = 0.35 + 0.65*max(dot(n, L), 0); rgb *= s; alpha unchanged.
// blend: straight src-over in 8-bit space with alpha accumulation:
//   c = (new.c*255)*a + o.c*(1-a);  A = a*255 + o.a*(1-a);  pack as 0xAARRGGBB (BGRA bytes, like the teapot).
struct VoxelIcon {
    using Vertex = euc_vertex_voxel; using VertexData = VecN<7>; using Fragment = Rgba; using Pixel = uint32_t;
    Mat4 mvp; f32x3 light_dir;
    std::pair<f32x4, VertexData> vertex(const Vertex& v) const {
        const float k = 1.0f / 255.0f;
        VertexData d{{v.normal[0], v.normal[1], v.normal[2], (float)v.rgba[0] * k, (float)v.rgba[1] * k, (float)v.rgba[2] * k, (float)v.rgba[3] * k}};
        return {mat4_mul_vec4(mvp, {v.pos[0], v.pos[1], v.pos[2], 1.0f}), d};
    }
    Rgba fragment(const VertexData& d) const {
        float s = 0.35f + 0.65f * f32_max(dot3({d[0], d[1], d[2]}, light_dir), 0.0f);
        return Rgba{{d[3] * s, d[4] * s, d[5] * s, d[6]}};
    }
    uint32_t blend(uint32_t old, const Rgba& n) const {
        float a = n[3], ia = 1.0f - a;
        float ob = (float)(old & 0xffu), og = (float)((old >> 8) & 0xffu), orr = (float)((old >> 16) & 0xffu), oa = (float)(old >> 24);
        float r = (n[0] * 255.0f) * a + orr * ia;
        float g = (n[1] * 255.0f) * a + og * ia;
        float b = (n[2] * 255.0f) * a + ob * ia;
        float A = a * 255.0f + oa * ia;
        auto q = [](float c) { return f32_as_u8(f32_min(f32_max(c, 0.0f), 255.0f)); };
        return pack_le(q(b), q(g), q(r), q(A));
    }
};

// ---- examples/triangle.rs:7-25, examples/spinning_cube.rs:5-29 -----------------------------------------
struct VertexColor {
    using Vertex = euc_vertex_p4c4; using VertexData = Rgba; using Fragment = Rgba; using Pixel = uint32_t;
    Mat4 mvp;
    std::pair<f32x4, Rgba> vertex(const Vertex& v) const {  // spinning_cube.rs:17-19
        return {mat4_mul_vec4(mvp, {v.pos[0], v.pos[1], v.pos[2], v.pos[3]}), Rgba{{v.rgba[0], v.rgba[1], v.rgba[2], v.rgba[3]}}};
    }
    Rgba fragment(const Rgba& c) const { return c; }
    uint32_t blend(uint32_t, const Rgba& c) const {  // triangle.rs:23-25: (e * 255.0) as u8, RGBA
        return pack_le(f32_as_u8(c[0] * 255.0f), f32_as_u8(c[1] * 255.0f), f32_as_u8(c[2] * 255.0f), f32_as_u8(c[3] * 255.0f));
    }
};

// ---- examples/wireframes.rs:5-37 -------------------------------------------------------------------------
struct Wireframe {
    using Vertex = euc_vertex_pn; using VertexData = Unit; using Fragment = Rgba; using Pixel = uint32_t;
    Mat4 m, v, p;
    std::pair<f32x4, Unit> vertex(const Vertex& vx) const {  // :19-23
        f32x4 wpos = mat4_mul_vec4(m, {vx.pos[0], vx.pos[1], vx.pos[2], 1.0f});
        return {mat4_mul_vec4(p, mat4_mul_vec4(v, wpos)), Unit{}};
    }
    Rgba fragment(Unit) const { return Rgba{{1.0f, 0.0f, 0.0f, 1.0f}}; }  // Rgba::red()
    uint32_t blend(uint32_t, const Rgba& rgba) const {                    // :31-36, BGRA
        uint8_t c[4];
        for (int i = 0; i < 4; ++i) { float e = rgba[i]; e = e < 0.0f ? 0.0f : (e > 1.0f ? 1.0f : e); c[i] = f32_as_u8(e * 255.0f); }
        return pack_le(c[2], c[1], c[0], c[3]);
    }
};

// ---------------------------------------------------------------------------------------------------------
// src/pipeline.rs:248-300 — Pipeline::render: target-size selection, vertex stage, primitive assembly.
// ---------------------------------------------------------------------------------------------------------
struct HostTexture { const void* data; uint32_t w, h; };

struct RenderArgs {
    const euc_pipeline_desc* desc;
    const void* uniforms;
    const uint8_t* vertices; uint32_t stride, n_vertices;
    const uint32_t* indices; uint32_t n_indices;
    uint32_t first, count; int32_t base_vertex;  // stream range
    uint32_t* pixel; float* depth; uint32_t w, h;  // pixel/depth may be null (= Empty)
    HostTexture tex[EUC_MAX_SAMPLERS];
    unsigned n_threads;
    RenderStats* stats; SetupDump* dump;
    uint32_t row_begin, row_end;
};

template <class Pipe> int render_with(const Pipe& pipe, const RenderArgs& a) {
    const euc_pipeline_desc& d = *a.desc;
    PixelMode pm{d.pixel_write != 0};
    DepthMode dm{d.depth_test != EUC_DEPTH_NONE,
                 d.depth_test == EUC_DEPTH_LESS ? Ordering::Less : (d.depth_test == EUC_DEPTH_EQUAL ? Ordering::Equal : Ordering::Greater),
                 d.depth_write != 0};
    // pipeline.rs:256-270
    if (!pm.write && !dm.uses_depth()) return EUC_OK;
    if (pm.write && !a.pixel) return EUC_E_INVALID;        // Empty pixel target has size [0,0]: nothing to render into
    if (dm.uses_depth() && !a.depth) return EUC_E_INVALID;
    usize tgt_size[2] = {a.w, a.h};

    CoordinateMode coords{d.y_axis_up ? YAxisDirection::Up : YAxisDirection::Down, d.z_clip_enabled != 0, d.z_clip_min, d.z_clip_max};
    CullMode cull = d.cull_mode == EUC_CULL_NONE ? CullMode::None : (d.cull_mode == EUC_CULL_BACK ? CullMode::Back : CullMode::Front);
    PrimKind kind = d.primitive_kind == EUC_PRIM_TRIANGLE_LIST ? PrimKind::TriangleList
                  : (d.primitive_kind == EUC_PRIM_LINE_LIST ? PrimKind::LineList : PrimKind::LineTriangleList);

    // pipeline.rs:273-289 + index.rs:52-54: every stream element runs the vertex shader.
    using VOut = std::pair<f32x4, typename Pipe::VertexData>;
    std::vector<VOut> stream;
    stream.reserve(a.count);
    for (uint32_t i = 0; i < a.count; ++i) {
        uint64_t vi;
        if (a.indices) {
            if (a.first + i >= a.n_indices) return EUC_E_OUT_OF_BOUNDS;
            vi = (uint64_t)((int64_t)a.indices[a.first + i] + a.base_vertex);
        } else {
            vi = (uint64_t)a.first + i + (int64_t)a.base_vertex;
        }
        if (vi >= a.n_vertices) return EUC_E_OUT_OF_BOUNDS;
        typename Pipe::Vertex v;
        std::memcpy(&v, a.vertices + (size_t)vi * a.stride, sizeof(v));
        stream.push_back(pipe.vertex(v));
    }
    // primitive assembly: collect_primitive drops a trailing partial primitive (pipeline.rs:283)
    std::vector<VOut> assembled;
    if (kind == PrimKind::TriangleList) {
        stream.resize(stream.size() / 3 * 3);
        assembled.swap(stream);
    } else if (kind == PrimKind::LineList) {
        stream.resize(stream.size() / 2 * 2);
        assembled.swap(stream);
    } else {  // primitives.rs:56-76: a b, b c, c a
        for (size_t t = 0; t + 3 <= stream.size(); t += 3) {
            assembled.push_back(stream[t]); assembled.push_back(stream[t + 1]);
            assembled.push_back(stream[t + 1]); assembled.push_back(stream[t + 2]);
            assembled.push_back(stream[t + 2]); assembled.push_back(stream[t]);
        }
    }
    usize msaa_level = (usize)std::min(std::max(d.msaa_level, 0), 6);  // pipeline.rs:291-294
    usize rb = a.row_begin, re = a.row_end ? a.row_end : ~(usize)0;

    Buffer2d<float> depth_buf; depth_buf.items = a.depth; depth_buf.size[0] = a.w; depth_buf.size[1] = a.h;
    Empty<float> depth_empty;
    if constexpr (std::is_same<typename Pipe::Pixel, Unit>::value) {
        Empty<Unit> px;
        if (a.depth) render_par(pipe, assembled, kind, tgt_size, px, depth_buf, msaa_level, pm, dm, coords, cull, a.n_threads, a.stats, a.dump, rb, re);
        else render_par(pipe, assembled, kind, tgt_size, px, depth_empty, msaa_level, pm, dm, coords, cull, a.n_threads, a.stats, a.dump, rb, re);
    } else {
        Buffer2d<uint32_t> px; px.items = a.pixel; px.size[0] = a.w; px.size[1] = a.h;
        Empty<uint32_t> px_empty;
        if (a.pixel && a.depth) render_par(pipe, assembled, kind, tgt_size, px, depth_buf, msaa_level, pm, dm, coords, cull, a.n_threads, a.stats, a.dump, rb, re);
        else if (a.pixel) render_par(pipe, assembled, kind, tgt_size, px, depth_empty, msaa_level, pm, dm, coords, cull, a.n_threads, a.stats, a.dump, rb, re);
        else if (a.depth) render_par(pipe, assembled, kind, tgt_size, px_empty, depth_buf, msaa_level, pm, dm, coords, cull, a.n_threads, a.stats, a.dump, rb, re);
    }
    return EUC_OK;
}

static f32x3 load3(const float* p) { return {p[0], p[1], p[2]}; }

int render_dispatch(const RenderArgs& a) {
    const euc_pipeline_desc& d = *a.desc;
    if (a.w > 20000u << std::min(std::max(d.msaa_level, 0), 6)) return EUC_E_UNSUPPORTED;
    switch (d.pipeline_id) {
        case EUC_PIPE_TEAPOT_SHADOW: {
            if (d.uniform_bytes < sizeof(euc_uniforms_teapot_shadow)) return EUC_E_INVALID;
            auto* u = (const euc_uniforms_teapot_shadow*)a.uniforms;
            TeapotShadow p{load_mat4(u->mvp)};
            return render_with(p, a);
        }
        case EUC_PIPE_TEAPOT_PHONG: {
            if (d.uniform_bytes < sizeof(euc_uniforms_teapot_phong) || !a.tex[0].data) return EUC_E_INVALID;
            auto* u = (const euc_uniforms_teapot_phong*)a.uniforms;
            Teapot p{load_mat4(u->m), load_mat4(u->v), load_mat4(u->p), load_mat4(u->light_vp), load3(u->light_pos), load3(u->cam_pos),
                     F32Sampler{TexF32{(const float*)a.tex[0].data, a.tex[0].w, a.tex[0].h}, d.samplers[0].filter, d.samplers[0].wrap}};
            return render_with(p, a);
        }
        case EUC_PIPE_TEX_CUBE: {
            if (d.uniform_bytes < sizeof(euc_uniforms_tex_cube) || !a.tex[0].data) return EUC_E_INVALID;
            auto* u = (const euc_uniforms_tex_cube*)a.uniforms;
            TexCube p{load_mat4(u->mvp), RgbaSampler{TexRgba8AsF32{(const uint8_t*)a.tex[0].data, a.tex[0].w, a.tex[0].h}, d.samplers[0].filter, d.samplers[0].wrap}};
            return render_with(p, a);
        }
        case EUC_PIPE_BLEND_TRIS: return render_with(BlendTris{}, a);
        case EUC_PIPE_VOXEL_ICON: {
            if (d.uniform_bytes < sizeof(euc_uniforms_voxel_icon)) return EUC_E_INVALID;
            auto* u = (const euc_uniforms_voxel_icon*)a.uniforms;
            VoxelIcon p{load_mat4(u->mvp), load3(u->light_dir)};
            return render_with(p, a);
        }
        case EUC_PIPE_VERTEX_COLOR: {
            if (d.uniform_bytes < sizeof(euc_uniforms_vertex_color)) return EUC_E_INVALID;
            auto* u = (const euc_uniforms_vertex_color*)a.uniforms;
            VertexColor p{load_mat4(u->mvp)};
            return render_with(p, a);
        }
        case EUC_PIPE_WIREFRAME: {
            if (d.uniform_bytes < sizeof(euc_uniforms_wireframe)) return EUC_E_INVALID;
            auto* u = (const euc_uniforms_wireframe*)a.uniforms;
            Wireframe p{load_mat4(u->m), load_mat4(u->v), load_mat4(u->p)};
            return render_with(p, a);
        }
        default: return EUC_E_INVALID;
    }
}

}  // namespace euc

// ---------------------------------------------------------------------------------------------------------
// C entry points (ctypes).  Targets and textures are plain host arrays.
// ---------------------------------------------------------------------------------------------------------
extern "C" {

struct oracle_texture { const void* data; uint32_t w, h; };

struct oracle_stats { uint64_t primitives, fragments; double seconds; };

// One Pipeline::render.  pixel/depth may be NULL (= Empty target).  n_threads: 0 = all host cores, as
// `available_parallelism()`; 1 = the band loop run by a single thread (results are identical by construction).
// setup_dump: NULL or an array of n_primitives euc::SetupDump records (triangle lists only).
int oracle_render(const euc_pipeline_desc* desc, const void* vertices, uint32_t stride, uint32_t n_vertices,
                  const uint32_t* indices, uint32_t n_indices, uint32_t first, uint32_t count, int32_t base_vertex,
                  uint32_t* pixel, float* depth, uint32_t w, uint32_t h, const oracle_texture* textures,
                  unsigned n_threads, uint32_t row_begin, uint32_t row_end, oracle_stats* stats, void* setup_dump) {
    if (!desc) return EUC_E_INVALID;
    euc::RenderArgs a{};
    a.desc = desc; a.uniforms = desc->uniforms;
    a.vertices = (const uint8_t*)vertices; a.stride = stride; a.n_vertices = n_vertices;
    a.indices = indices; a.n_indices = n_indices; a.first = first; a.count = count; a.base_vertex = base_vertex;
    a.pixel = pixel; a.depth = depth; a.w = w; a.h = h;
    for (int i = 0; i < EUC_MAX_SAMPLERS; ++i) {
        if (textures) a.tex[i] = {textures[i].data, textures[i].w, textures[i].h};
        else a.tex[i] = {nullptr, 0, 0};
    }
    a.n_threads = n_threads; a.row_begin = row_begin; a.row_end = row_end;
    euc::RenderStats rs;
    a.stats = &rs; a.dump = (euc::SetupDump*)setup_dump;
    auto t0 = std::chrono::steady_clock::now();
    int rc = euc::render_dispatch(a);
    auto t1 = std::chrono::steady_clock::now();
    if (stats) {
        stats->primitives = rs.primitives.load(); stats->fragments = rs.fragments.load();
        stats->seconds = std::chrono::duration<double>(t1 - t0).count();
    }
    return rc;
}

// Target::clear (buffer.rs:213-218) for 4-byte texels.
void oracle_clear_u32(uint32_t* buf, uint64_t n, uint32_t v) { for (uint64_t i = 0; i < n; ++i) buf[i] = v; }

unsigned oracle_hardware_concurrency() { return std::max(1u, std::thread::hardware_concurrency()); }
uint32_t oracle_setup_dump_bytes() { return (uint32_t)sizeof(euc::SetupDump); }

// Leaf-semantics probes for known-answer tests.
uint64_t oracle_f32_as_usize(float f) { return euc::f32_as_usize(f); }
uint32_t oracle_f32_as_u8(float f) { return euc::f32_as_u8(f); }
float oracle_f32_min(float a, float b) { return euc::f32_min(a, b); }
float oracle_f32_max(float a, float b) { return euc::f32_max(a, b); }
float oracle_fract(float x) { return euc::f32_fract(x); }
float oracle_rem_euclid(float x, float r) { return euc::f32_rem_euclid(x, r); }
float oracle_wrap(int wrap, float e) {
    using L = euc::Linear<euc::TexF32>;
    switch (wrap) {
        case 1: return euc::Clamped<L>::map(e);
        case 2: return euc::Tiled<L>::map(e);
        case 3: return euc::Mirrored<L>::map(e);
        default: return e;
    }
}
float oracle_sample_f32(const float* data, uint32_t w, uint32_t h, int filter, int wrap, float x, float y) {
    return euc::F32Sampler{euc::TexF32{data, w, h}, filter, wrap}.sample(x, y);
}
void oracle_sample_rgba8(const uint8_t* data, uint32_t w, uint32_t h, int filter, int wrap, float x, float y, float* out4) {
    euc::Rgba r = euc::RgbaSampler{euc::TexRgba8AsF32{data, w, h}, filter, wrap}.sample(x, y);
    for (int i = 0; i < 4; ++i) out4[i] = r[i];
}
// band table (pipeline.rs:328-330): returns group_rows, writes needed_threads
uint64_t oracle_band_rows(uint64_t w, uint64_t h, uint32_t msaa_level, uint64_t threads, uint64_t* needed_threads) {
    uint64_t group_rows = 20000ull * (1ull << msaa_level) / std::max<uint64_t>(w, 1);
    if (needed_threads) *needed_threads = group_rows ? std::min(h / group_rows, threads) : 0;
    return group_rows;
}
}
