// A user-written pipeline for euc_pipeline_register(): vertex colours multiplied by a uniform tint, src-over blend.
// Same static interface as the built-in pipelines in euc_b200/csrc/shaders.cuh; compiled at run time by NVRTC with
// --fmad=false, so unfused f32 arithmetic here behaves like the Rust reference's.
struct TintPipe {
    static constexpr int V = 4;                          // VertexData: rgba
    static constexpr bool HAS_FRAGMENT = true;
    static constexpr bool BLEND_IGNORES_OLD = false;     // blend reads the old pixel -> immediate mode, submission order
    struct Uniforms { float mvp[16]; float tint[4]; };
    static constexpr uint32_t VERTEX_BYTES = 32;         // euc_vertex_p4c4
    static __device__ __forceinline__ void vertex(const Uniforms& u, const uint8_t* vp, float4& clip, float* var) {
        const float4 p = *(const float4*)vp;
        const float4 c = *(const float4*)(vp + 16);
        clip = mat4_mul_vec4(u.mvp, p.x, p.y, p.z, p.w);
        var[0] = c.x; var[1] = c.y; var[2] = c.z; var[3] = c.w;
    }
    static __device__ __forceinline__ void fragment(const Uniforms& u, const SamplerDev*, const float* v, float* frag) {
        frag[0] = v[0] * u.tint[0]; frag[1] = v[1] * u.tint[1]; frag[2] = v[2] * u.tint[2]; frag[3] = v[3] * u.tint[3];
    }
    static __device__ __forceinline__ uint32_t blend(uint32_t old, const float* n) {
        float a = n[3], ia = 1.0f - a;
        float c0 = (n[0] * 255.0f) * a + (float)(old & 0xffu) * ia;
        float c1 = (n[1] * 255.0f) * a + (float)((old >> 8) & 0xffu) * ia;
        float c2 = (n[2] * 255.0f) * a + (float)((old >> 16) & 0xffu) * ia;
        return pack_le(r_as_u8(r_min(r_max(c0, 0.0f), 255.0f)), r_as_u8(r_min(r_max(c1, 0.0f), 255.0f)),
                       r_as_u8(r_min(r_max(c2, 0.0f), 255.0f)), 255u);
    }
};
