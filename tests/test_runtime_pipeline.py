"""Row N3: pipelines compiled at run time (NVRTC) through euc_pipeline_register.  A user-written shader has no reference
counterpart, so parity is pinned by writing user pipelines whose arithmetic coincides with a built-in one for which
the oracle exists."""
import os

import numpy as np
import pytest

import euc_b200 as e
from euc_b200 import scenes
from oracle import oracle
from conftest import assert_colour_within_1lsb, assert_depth_bit_exact

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINT_SRC = open(os.path.join(ROOT, "examples", "user_pipeline_tint.cu")).read()

DEFERRED_SRC = r"""
struct UserVertexColor {   // same arithmetic as examples/triangle.rs / spinning_cube.rs, written by the "user"
    static constexpr int V = 4;
    static constexpr bool HAS_FRAGMENT = true;
    static constexpr bool BLEND_IGNORES_OLD = true;   // -> deferred: one fragment + blend per pixel (resolve kernel)
    struct Uniforms { float mvp[16]; };
    static constexpr uint32_t VERTEX_BYTES = 32;
    static __device__ __forceinline__ void vertex(const Uniforms& u, const uint8_t* vp, float4& clip, float* var) {
        const float4 p = *(const float4*)vp; const float4 c = *(const float4*)(vp + 16);
        clip = mat4_mul_vec4(u.mvp, p.x, p.y, p.z, p.w);
        var[0] = c.x; var[1] = c.y; var[2] = c.z; var[3] = c.w;
    }
    static __device__ __forceinline__ void fragment(const Uniforms&, const SamplerDev*, const float* v, float* f) {
        f[0] = v[0]; f[1] = v[1]; f[2] = v[2]; f[3] = v[3];
    }
    static __device__ __forceinline__ uint32_t blend(uint32_t, const float* f) {
        return pack_le(r_as_u8(f[0] * 255.0f), r_as_u8(f[1] * 255.0f), r_as_u8(f[2] * 255.0f), r_as_u8(f[3] * 255.0f));
    }
};
"""


def _tris(n, seed):
    r = scenes.u01(seed, n * 3 * 8).reshape(n, 3, 8)
    v = np.zeros((n, 3), dtype=e.VERTEX_P4C4)
    wv = 0.5 + 1.5 * r[:, :, 4]
    v["pos"][:, :, 0] = ((r[:, :1, 0] * 2.2 - 1.1) + (r[:, :, 2] - 0.5) * 0.5) * wv
    v["pos"][:, :, 1] = ((r[:, :1, 1] * 2.2 - 1.1) + (r[:, :, 3] - 0.5) * 0.5) * wv
    v["pos"][:, :, 2] = (r[:, :, 5] * 0.9 + 0.05) * wv
    v["pos"][:, :, 3] = wv
    v["rgba"][:, :, :3] = r[:, :, 5:8]
    v["rgba"][:, :, 3] = 0.25 + 0.5 * r[:, :, 6]
    return v.reshape(-1)


def _uniforms(tint):
    return np.eye(4, dtype=np.float32).T.tobytes() + np.asarray(tint, dtype=np.float32).tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("tint", [(1.0, 1.0, 1.0, 1.0), (0.5, 0.25, 1.0, 1.0)])
def test_user_tint_pipeline_matches_oracle(tint):
    ctx = e.default_context()
    pid = ctx.register_pipeline(TINT_SRC, "TintPipe")
    assert pid >= e.abi.PIPE_USER_BASE
    w, h = 800, 600
    verts = _tris(600, 42)
    pipe = e.UserPipeline(pid, e.VERTEX_P4C4, _uniforms(tint), depth=e.DepthMode.LESS_WRITE, cull=e.CullMode.NONE)
    px, z = e.Buffer2d.fill([w, h], 0xFF000000, dtype=np.uint32), e.Buffer2d.fill([w, h], 1.0)
    ctx.set_stats(True)
    pipe.render(verts, px, z)
    frags = ctx.get_stats()["fragments"]
    # oracle: BLEND_TRIS on colours pre-multiplied by the (power-of-two) tint -- scaling by 2^-k commutes with rounding
    ref_v = verts.copy()
    ref_v["rgba"] *= np.asarray(tint, dtype=np.float32)[None, :]
    rpx, rz = np.full((h, w), 0xFF000000, np.uint32), np.full((h, w), 1.0, np.float32)
    rs = oracle.render(e.BlendTris(), ref_v, rpx, rz, n_threads=0)
    assert frags == rs["fragments"] > 10000
    assert_depth_bit_exact(z.raw(), rz, "user pipeline depth")
    assert np.array_equal(px.raw(), rpx), "user pipeline colour (no transcendental: expected bit-exact)"


@pytest.mark.gpu
@pytest.mark.parametrize("aa", [None, 1])
def test_user_deferred_pipeline_matches_vertex_color(aa):
    ctx = e.default_context()
    pid = ctx.register_pipeline(DEFERRED_SRC, "UserVertexColor")
    w, h = 1000, 400
    verts = _tris(400, 7)
    kw = dict(depth=e.DepthMode.LESS_WRITE, cull=e.CullMode.Back, aa=e.AaMode.Msaa(aa) if aa else None)
    px, z = e.Buffer2d.fill([w, h], 0, dtype=np.uint32), e.Buffer2d.fill([w, h], 1.0)
    e.UserPipeline(pid, e.VERTEX_P4C4, np.eye(4, dtype=np.float32).tobytes(), **kw).render(verts, px, z)
    rpx, rz = np.zeros((h, w), np.uint32), np.full((h, w), 1.0, np.float32)
    oracle.render(e.VertexColor(**kw), verts, rpx, rz, n_threads=0)
    assert_depth_bit_exact(z.raw(), rz, "deferred user pipeline")
    assert_colour_within_1lsb(px.raw(), rpx, "deferred user pipeline")
    assert (rpx != 0).sum() > 10000


@pytest.mark.gpu
def test_user_pipeline_draws_lines():
    """Run-time pipelines accept LineList / LineTriangleList like the built-in ones."""
    ctx = e.default_context()
    pid = ctx.register_pipeline(DEFERRED_SRC, "UserVertexColor")
    w, h = 640, 360
    verts = _tris(300, 11)
    kw = dict(depth=e.DepthMode.LESS_WRITE, primitives=e.LineTriangleList)
    px, z = e.Buffer2d.fill([w, h], 0, dtype=np.uint32), e.Buffer2d.fill([w, h], 1.0)
    e.UserPipeline(pid, e.VERTEX_P4C4, np.eye(4, dtype=np.float32).tobytes(), **kw).render(verts, px, z)
    rpx, rz = np.zeros((h, w), np.uint32), np.full((h, w), 1.0, np.float32)
    rs = oracle.render(e.VertexColor(**kw), verts, rpx, rz, n_threads=0)
    assert rs["primitives"] == 900 and rs["fragments"] > 5000
    assert_depth_bit_exact(z.raw(), rz, "user pipeline lines")
    assert_colour_within_1lsb(px.raw(), rpx, "user pipeline lines")


@pytest.mark.gpu
def test_user_pipeline_compile_error_is_reported():
    ctx = e.default_context()
    with pytest.raises(e.EucError) as ei:
        ctx.register_pipeline("struct Broken { static constexpr int V = ; };", "Broken")
    assert ei.value.code == e.abi.E_INVALID and "user_pipeline.cu" in str(ei.value)


def _bundled_nvrtc():
    """The NVRTC that ships with torch's CUDA wheels (older than the toolkit's): it is the one dlopen("libnvrtc.so.12")
    finds when torch was imported first, so the headers must build with it too."""
    try:
        import nvidia.cuda_nvrtc as m
    except ImportError:
        return None
    path = os.path.join(list(m.__path__)[0], "lib", "libnvrtc.so.12")
    return path if os.path.exists(path) else None


@pytest.mark.parametrize("which", ["toolkit", "bundled"])
def test_headers_are_nvrtc_clean(which):
    """NVRTC needs no GPU to compile: the kernel headers plus the example user pipeline must build for sm_100a."""
    import subprocess, sys
    env = dict(os.environ)
    if which == "bundled":
        lib = _bundled_nvrtc()
        if lib is None:
            pytest.skip("no bundled NVRTC in this environment")
        env["NVRTC_LIB"] = lib
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "nvrtc_check.py"), os.path.join(ROOT, "examples", "user_pipeline_tint.cu"), "TintPipe"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.stdout.startswith("rc 0"), out.stdout + out.stderr
