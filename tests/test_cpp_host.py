"""The header-only C++ host mirror (include/euc_b200.hpp) compiles and links against the C-ABI library."""
import os
import subprocess
import textwrap

import euc_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_mirror_compiles_and_links(tmp_path):
    src = tmp_path / "host.cpp"
    src.write_text(textwrap.dedent("""
        #include "euc_b200.hpp"
        #include <cstdio>
        int main() {
            try {
                euc::Context ctx(0);
                auto shadow = euc::Buffer2d<float>::fill(ctx, 512, 512, 1.0f);
                auto color = euc::Buffer2d<uint32_t>::fill(ctx, 640, 480, 0u);
                auto depth = euc::Buffer2d<float>::fill(ctx, 640, 480, 1.0f);
                float mvp[16] = {1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1};
                euc_vertex_pn tri[3] = {{{-1,-1,0.5f},{0,0,1}}, {{1,-1,0.5f},{0,0,1}}, {{0,1,0.5f},{0,0,1}}};
                euc::Empty none;
                euc::TeapotShadow(mvp).render(ctx, tri, 3, nullptr, 0, none, shadow);
                euc_uniforms_teapot_phong u{};
                euc::Teapot(u, shadow).render(ctx, tri, 3, nullptr, 0, color, depth);
                std::printf("%zu\\n", color.raw().size());
            } catch (const euc::Error& e) { std::printf("error %d: %s\\n", e.code, e.what()); return e.code == EUC_E_CUDA ? 0 : 1; }
            return 0;
        }
    """))
    exe = tmp_path / "host"
    libdir = os.path.dirname(euc_b200.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-leuc_b200", f"-Wl,-rpath,{libdir}"])
    # without a GPU the program must report EUC_E_CUDA from euc_init (no CPU fallback) and exit 0
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([str(exe)], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "error -4" in out.stdout
