// euc_b200.hpp — header-only C++ mirror of euc's host-side surface over the C ABI (include/euc_b200.h).
// The reference is compiled code (Rust); with no Rust toolchain available this is the compiled-language host side.
// Names follow the reference: Buffer2d (src/buffer.rs), Empty (src/texture.rs:285-319), DepthMode / PixelMode /
// CoordinateMode / AaMode (src/pipeline.rs:14-163), CullMode (src/rasterizer/mod.rs:10-18), Pipeline::render
// (src/pipeline.rs:248).  Errors (the reference's panics) become euc::Error exceptions carrying the EUC_E_* code.
#pragma once
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "euc_b200.h"

namespace euc {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

class Context {
public:
    explicit Context(int device = 0) {
        int rc = euc_init(device, &ctx_);
        if (rc != EUC_OK) throw Error(rc, "euc_init failed (no CUDA device? there is no CPU fallback)");
    }
    ~Context() { if (ctx_) euc_shutdown(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    euc_ctx* raw() const { return ctx_; }
    void check(int rc) const { if (rc != EUC_OK) throw Error(rc, euc_last_error(ctx_)); }
    void sync() const { check(euc_sync(ctx_)); }
    // `pixel.clear(a); depth.clear(b); pipe.render(..)` with the clears fused into the next render (null = leave alone)
    void render_clear(const uint32_t* pixel_value, const float* depth_value) const { check(euc_render_clear(ctx_, pixel_value, depth_value)); }
private:
    euc_ctx* ctx_ = nullptr;
};

struct DepthMode {  // src/pipeline.rs:14-46
    int test; bool write;
    static constexpr DepthMode NONE() { return {EUC_DEPTH_NONE, false}; }
    static constexpr DepthMode LESS_WRITE() { return {EUC_DEPTH_LESS, true}; }
    static constexpr DepthMode GREATER_WRITE() { return {EUC_DEPTH_GREATER, true}; }
    static constexpr DepthMode LESS_PASS() { return {EUC_DEPTH_LESS, false}; }
    static constexpr DepthMode GREATER_PASS() { return {EUC_DEPTH_GREATER, false}; }
    bool uses_depth() const { return test != EUC_DEPTH_NONE || write; }
};
struct PixelMode { bool write; static constexpr PixelMode WRITE() { return {true}; } static constexpr PixelMode PASS() { return {false}; } };
struct CoordinateMode {  // src/pipeline.rs:94-157
    int handedness; bool y_up; bool z_clip; float z_min, z_max;
    static constexpr CoordinateMode OPENGL() { return {EUC_HAND_RIGHT, true, true, -1.0f, 1.0f}; }
    static constexpr CoordinateMode VULKAN() { return {EUC_HAND_LEFT, false, true, 0.0f, 1.0f}; }
    static constexpr CoordinateMode METAL() { return {EUC_HAND_RIGHT, false, true, 0.0f, 1.0f}; }
    static constexpr CoordinateMode DIRECTX() { return {EUC_HAND_LEFT, true, true, 0.0f, 1.0f}; }
    CoordinateMode without_z_clip() const { CoordinateMode c = *this; c.z_clip = false; return c; }
};
struct AaMode { int level; static constexpr AaMode None() { return {0}; } static constexpr AaMode Msaa(int l) { return {l}; } };
enum class CullMode { None = EUC_CULL_NONE, Back = EUC_CULL_BACK, Front = EUC_CULL_FRONT };

// Device-resident Buffer2d<T>, sizeof(T) == 4.
template <class T> class Buffer2d {
    static_assert(sizeof(T) == 4, "4-byte texels only");
public:
    static Buffer2d fill(const Context& c, unsigned w, unsigned h, T item, unsigned layers = 1) {  // buffer.rs:60-67
        Buffer2d b(c, w, h, layers);
        b.clear(item);
        return b;
    }
    Buffer2d(Buffer2d&& o) noexcept : c_(o.c_), h_(o.h_), w_(o.w_), hgt_(o.hgt_), layers_(o.layers_) { o.h_ = 0; }
    ~Buffer2d() { if (h_) euc_buf_destroy(c_->raw(), h_); }
    void clear(T item) { c_->check(euc_buf_clear(c_->raw(), h_, &item)); }                            // buffer.rs:213-218
    std::vector<T> raw() const {                                                                      // buffer.rs:104-107
        std::vector<T> v((size_t)w_ * hgt_ * layers_);
        c_->check(euc_buf_download(c_->raw(), h_, v.data(), v.size() * 4));
        return v;
    }
    void upload(const T* p) { c_->check(euc_buf_upload(c_->raw(), h_, p, (size_t)w_ * hgt_ * layers_ * 4)); }
    euc_buf handle() const { return h_; }
    unsigned width() const { return w_; }
    unsigned height() const { return hgt_; }
    // Texture::linear() / nearest() + Sampler::clamped() / tiled() / mirrored()
    euc_sampler_desc sampler(int format, int filter, int wrap) const { return euc_sampler_desc{h_, format, filter, wrap, 0}; }
private:
    Buffer2d(const Context& c, unsigned w, unsigned h, unsigned layers) : c_(&c), w_(w), hgt_(h), layers_(layers) {
        c.check(euc_buf_create(c.raw(), w, h, layers, 4, &h_));
    }
    const Context* c_; euc_buf h_ = 0; unsigned w_, hgt_, layers_;
};
struct Empty { euc_buf handle() const { return 0; } };  // texture.rs:285-319

// Mirror of `trait Pipeline`: a concrete pipeline supplies its id, uniform block and samplers; the getters carry
// the reference's defaults (src/pipeline.rs:178-209) and can be overridden.
struct Pipeline {
    int pipeline_id = -1;
    std::vector<unsigned char> uniforms;
    euc_sampler_desc samplers[EUC_MAX_SAMPLERS] = {};
    PixelMode pixel_mode = PixelMode::WRITE();
    DepthMode depth_mode = DepthMode::NONE();
    CoordinateMode coordinate_mode = CoordinateMode::VULKAN();
    AaMode aa_mode = AaMode::None();
    CullMode rasterizer_config = CullMode::Back;

    euc_pipeline_desc desc() const {
        euc_pipeline_desc d;
        std::memset(&d, 0, sizeof d);
        d.pipeline_id = pipeline_id; d.primitive_kind = EUC_PRIM_TRIANGLE_LIST; d.cull_mode = (int)rasterizer_config;
        d.depth_test = depth_mode.test; d.depth_write = depth_mode.write; d.pixel_write = pixel_mode.write;
        d.y_axis_up = coordinate_mode.y_up; d.handedness = coordinate_mode.handedness;
        d.z_clip_enabled = coordinate_mode.z_clip; d.z_clip_min = coordinate_mode.z_min; d.z_clip_max = coordinate_mode.z_max;
        d.msaa_level = aa_mode.level; d.uniforms = uniforms.data(); d.uniform_bytes = (uint32_t)uniforms.size();
        for (int i = 0; i < EUC_MAX_SAMPLERS; ++i) d.samplers[i] = samplers[i];
        return d;
    }
    // Pipeline::render(vertices, &mut pixel, &mut depth); `indices` = IndexedVertices (src/index.rs), may be null.
    template <class V, class P, class D>
    void render(const Context& c, const V* vertices, unsigned n_vertices, const uint32_t* indices, unsigned n_indices, P& pixel, D& depth) const {
        euc_pipeline_desc d = desc();
        c.check(euc_render(c.raw(), &d, vertices, (uint32_t)sizeof(V), n_vertices, indices, n_indices, pixel.handle(), depth.handle()));
    }
};

// benches/teapot.rs:10-51
inline Pipeline TeapotShadow(const float mvp[16]) {
    Pipeline p;
    p.pipeline_id = EUC_PIPE_TEAPOT_SHADOW;
    p.uniforms.assign((const unsigned char*)mvp, (const unsigned char*)mvp + 64);
    p.pixel_mode = PixelMode::PASS(); p.depth_mode = DepthMode::LESS_WRITE(); p.rasterizer_config = CullMode::None;
    return p;
}
// benches/teapot.rs:53-142
inline Pipeline Teapot(const euc_uniforms_teapot_phong& u, const Buffer2d<float>& shadow) {
    Pipeline p;
    p.pipeline_id = EUC_PIPE_TEAPOT_PHONG;
    p.uniforms.assign((const unsigned char*)&u, (const unsigned char*)&u + sizeof u);
    p.depth_mode = DepthMode::LESS_WRITE();
    p.samplers[0] = shadow.sampler(EUC_TEXEL_F32, EUC_FILTER_LINEAR, EUC_WRAP_CLAMP);  // (&shadow).linear().clamped()
    return p;
}

}  // namespace euc
