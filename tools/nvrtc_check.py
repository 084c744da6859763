"""Compiles a user pipeline source with NVRTC exactly like euc_pipeline_register does (no GPU needed): checks that the
kernels' headers are NVRTC-clean.  usage: python tools/nvrtc_check.py examples/user_pipeline_tint.cu TintPipe"""
import ctypes as C, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, name = open(sys.argv[1]).read(), sys.argv[2]
tu = '#include "kernels.cuh"\nnamespace eucb {\n#line 1 "user_pipeline.cu"\n' + src + "\n}\nusing EucUserPipe = eucb::" + name + ";\n" + \
     "constexpr bool EUC_USER_DEFER = EucUserPipe::HAS_FRAGMENT && EucUserPipe::BLEND_IGNORES_OLD;\n" + \
     'extern "C" __global__ void euc_user_info(unsigned int* out) { out[0] = EucUserPipe::V; out[5] = eucb::RecLayout<EucUserPipe>::BYTES; out[6] = (unsigned int)eucb::raster_smem_bytes<EucUserPipe, EUC_USER_DEFER>(); }\n'
n = C.CDLL(os.environ.get("NVRTC_LIB", "/usr/local/cuda/lib64/libnvrtc.so.12"))
prog = C.c_void_p()
assert n.nvrtcCreateProgram(C.byref(prog), tu.encode(), b"euc_user_pipeline.cu", 0, None, None) == 0
names = [b"eucb::setup_kernel<EucUserPipe, false>", b"eucb::raster_kernel<EucUserPipe, false, EUC_USER_DEFER, false>", b"eucb::raster_kernel<EucUserPipe, true, EUC_USER_DEFER, false>",
         b"eucb::resolve_kernel<EucUserPipe, false, false>", b"eucb::resolve_kernel<EucUserPipe, true, false>",
         b"eucb::setup_lines_kernel<EucUserPipe>", b"eucb::raster_kernel<EucUserPipe, false, EUC_USER_DEFER, true>", b"eucb::raster_kernel<EucUserPipe, true, EUC_USER_DEFER, true>",
         b"eucb::resolve_kernel<EucUserPipe, false, true>", b"eucb::resolve_kernel<EucUserPipe, true, true>"]
for nm in names:
    n.nvrtcAddNameExpression(prog, nm)
opts = [b"--gpu-architecture=sm_100a", b"--fmad=false", b"-std=c++17", b"-lineinfo", b"-device-int128", b"-default-device",
        ("-I" + os.path.join(root, "euc_b200", "csrc")).encode(), ("-I" + os.path.join(root, "include")).encode(), b"-I/usr/local/cuda/include"]
arr = (C.c_char_p * len(opts))(*opts)
import time; t = time.time()
rc = n.nvrtcCompileProgram(prog, len(opts), arr)
sz = C.c_size_t(); n.nvrtcGetProgramLogSize(prog, C.byref(sz)); log = C.create_string_buffer(sz.value); n.nvrtcGetProgramLog(prog, log)
print("rc", rc, "seconds", round(time.time() - t, 1)); print(log.value.decode()[:3000])
if rc == 0:
    cs = C.c_size_t(); n.nvrtcGetCUBINSize(prog, C.byref(cs)); print("cubin bytes", cs.value)
