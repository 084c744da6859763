"""ctypes wrapper over oracle/libeuc_oracle.so — the CPU restatement of euc's render path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under euc_b200/ imports this module.  PARITY UNPINNED (see euc_oracle.hpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

from euc_b200 import abi
from euc_b200.core import Buffer2d, IndexedVertices, Pipeline

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libeuc_oracle.so")
_lib = None


class OracleTexture(C.Structure):
    _fields_ = [("data", C.c_void_p), ("w", C.c_uint32), ("h", C.c_uint32)]


class OracleStats(C.Structure):
    _fields_ = [("primitives", C.c_uint64), ("fragments", C.c_uint64), ("seconds", C.c_double)]


class SetupDump(C.Structure):
    _fields_ = [("culled", C.c_uint32), ("w_hom_origin", C.c_float * 3), ("w_hom_dx", C.c_float * 3), ("w_hom_dy", C.c_float * 3),
                ("z_hom", C.c_float * 3), ("verts_by_y", C.c_float * 6), ("bounds_min", C.c_uint32 * 2), ("bounds_max", C.c_uint32 * 2),
                ("no_verts_clipped", C.c_uint32)]


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("euc_oracle.cpp", "euc_oracle.hpp", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(_SO)
        l.oracle_render.restype = C.c_int
        l.oracle_render.argtypes = [C.POINTER(abi.PipelineDesc), C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32,
                                    C.c_uint32, C.c_uint32, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                    C.POINTER(OracleTexture), C.c_uint, C.c_uint32, C.c_uint32, C.POINTER(OracleStats), C.c_void_p]
        l.oracle_hardware_concurrency.restype = C.c_uint
        l.oracle_setup_dump_bytes.restype = C.c_uint32
        l.oracle_f32_as_usize.restype = C.c_uint64
        l.oracle_f32_as_usize.argtypes = [C.c_float]
        l.oracle_f32_as_u8.restype = C.c_uint32
        l.oracle_f32_as_u8.argtypes = [C.c_float]
        for n in ("oracle_f32_min", "oracle_f32_max", "oracle_rem_euclid"):
            getattr(l, n).restype = C.c_float
            getattr(l, n).argtypes = [C.c_float, C.c_float]
        l.oracle_fract.restype = C.c_float
        l.oracle_fract.argtypes = [C.c_float]
        l.oracle_wrap.restype = C.c_float
        l.oracle_wrap.argtypes = [C.c_int, C.c_float]
        l.oracle_sample_f32.restype = C.c_float
        l.oracle_sample_f32.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_float, C.c_float]
        l.oracle_sample_rgba8.restype = None
        l.oracle_sample_rgba8.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]
        l.oracle_band_rows.restype = C.c_uint64
        l.oracle_band_rows.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64)]
        assert l.oracle_setup_dump_bytes() == C.sizeof(SetupDump)
        _lib = l
    return _lib


def hardware_concurrency():
    return int(lib().oracle_hardware_concurrency())


def render(pipe: Pipeline, vertices, pixel, depth, n_threads=1, rows=None, draw=None, dump_setup=False):
    """Pipeline::render on the CPU oracle.

    vertices: numpy vertex array (stream) or IndexedVertices.  pixel / depth: numpy (h, w) uint32 / float32 arrays,
    modified in place, or None (= euc `Empty`).  Samplers of `pipe` must wrap numpy arrays.
    draw = (first, count, base_vertex) restricts the stream (batch rendering).  Returns a stats dict.
    """
    l = lib()
    keepalive = []

    def tex_handle(s):
        keepalive.append(s)
        return 1  # non-zero marks the sampler as bound; the texture itself travels in `textures`

    d, keep = pipe.build_desc(tex_handle)
    texs = (OracleTexture * abi.MAX_SAMPLERS)()
    for i, s in enumerate(pipe.samplers()):
        t = s.texture
        if isinstance(t, Buffer2d):
            raise TypeError("oracle samplers must wrap numpy arrays, not device buffers")
        t = np.ascontiguousarray(t)
        keepalive.append(t)
        texs[i].data = t.ctypes.data
        texs[i].h, texs[i].w = t.shape[0], t.shape[1]
    if isinstance(vertices, IndexedVertices):
        v, idx = np.ascontiguousarray(vertices.verts), vertices.indices
    else:
        v, idx = np.ascontiguousarray(vertices), None
    stride = v.dtype.itemsize if v.ndim == 1 else v.strides[0]
    stream_len = idx.size if idx is not None else v.shape[0]
    first, count, base_vertex = draw if draw is not None else (0, stream_len, 0)
    tgt = pixel if pixel is not None else depth
    if pixel is not None and depth is not None and pixel.shape != depth.shape:
        raise AssertionError("Pixel target size is compatible with depth target size")  # pipeline.rs:262-266
    if tgt is None:
        h = w = 0
    else:
        h, w = tgt.shape
    for t in (pixel, depth):
        if t is not None:
            assert t.flags["C_CONTIGUOUS"] and t.dtype.itemsize == 4
    stats = OracleStats()
    dump = None
    if dump_setup:
        dump = (SetupDump * max(count // 3, 1))()
    r0, r1 = rows if rows is not None else (0, 0)
    rc = l.oracle_render(C.byref(d), v.ctypes.data, stride, v.shape[0], idx.ctypes.data if idx is not None else None,
                         0 if idx is None else idx.size, first, count, base_vertex,
                         pixel.ctypes.data if pixel is not None else None, depth.ctypes.data if depth is not None else None,
                         w, h, texs, n_threads, r0, r1, C.byref(stats), C.cast(dump, C.c_void_p) if dump is not None else None)
    if rc != 0:
        raise RuntimeError(f"oracle_render failed: {abi.STATUS_NAMES.get(rc, rc)}")
    out = {"primitives": stats.primitives, "fragments": stats.fragments, "seconds": stats.seconds}
    if dump is not None:
        out["setup"] = dump
    return out
