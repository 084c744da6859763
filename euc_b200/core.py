"""Host-side mirror of euc's public surface for the `Pipeline::render` path, over the C ABI.

Names follow the reference crate (paths relative to the euc crate root):
  DepthMode / PixelMode / CoordinateMode / AaMode   src/pipeline.rs:14-163
  CullMode                                          src/rasterizer/mod.rs:10-18
  TriangleList / LineList / LineTriangleList        src/primitives.rs:21, :82, :49
  IndexedVertices                                   src/index.rs:4-18
  Buffer2d (fill / clear / raw / size)              src/buffer.rs
  Empty                                             src/texture.rs:285-319
  Sampler (linear / nearest, clamped/tiled/mirrored) src/texture.rs:51-95, src/sampler/mod.rs:44-70
  Pipeline.render(vertices, pixel, depth)           src/pipeline.rs:248-300
"""
import ctypes as C
from dataclasses import dataclass, replace
from typing import Optional

import numpy as np

from . import abi
from ._lib import EucError, load


# ---- modes ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class DepthMode:
    test: Optional[str]  # None | "Less" | "Equal" | "Greater"   (Option<Ordering>)
    write: bool

    def uses_depth(self):
        return self.test is not None or self.write


DepthMode.NONE = DepthMode(None, False)
DepthMode.LESS_WRITE = DepthMode("Less", True)
DepthMode.GREATER_WRITE = DepthMode("Greater", True)
DepthMode.LESS_PASS = DepthMode("Less", False)
DepthMode.GREATER_PASS = DepthMode("Greater", False)
_DEPTH_TEST = {None: abi.DEPTH_NONE, "Less": abi.DEPTH_LESS, "Equal": abi.DEPTH_EQUAL, "Greater": abi.DEPTH_GREATER}


@dataclass(frozen=True)
class PixelMode:
    write: bool


PixelMode.WRITE = PixelMode(True)
PixelMode.PASS = PixelMode(False)


@dataclass(frozen=True)
class CoordinateMode:
    handedness: str        # "Left" | "Right"
    y_axis_direction: str  # "Down" | "Up"
    z_clip_range: Optional[tuple]

    def without_z_clip(self):
        return replace(self, z_clip_range=None)


CoordinateMode.OPENGL = CoordinateMode("Right", "Up", (-1.0, 1.0))
CoordinateMode.VULKAN = CoordinateMode("Left", "Down", (0.0, 1.0))
CoordinateMode.METAL = CoordinateMode("Right", "Down", (0.0, 1.0))
CoordinateMode.DIRECTX = CoordinateMode("Left", "Up", (0.0, 1.0))


@dataclass(frozen=True)
class AaMode:
    level: int  # 0 = AaMode::None

    @staticmethod
    def Msaa(level):
        return AaMode(int(level))


AaMode.NONE = AaMode(0)


class CullMode:
    NONE, Back, Front = abi.CULL_NONE, abi.CULL_BACK, abi.CULL_FRONT


class TriangleList:
    kind = abi.PRIM_TRIANGLE_LIST


class LineList:
    kind = abi.PRIM_LINE_LIST


class LineTriangleList:
    kind = abi.PRIM_LINE_TRIANGLE_LIST


class IndexedVertices:
    """IndexedVertices::new(indices, verts) (src/index.rs:11-18); indices are u32 on the device."""

    def __init__(self, indices, verts):
        self.indices = np.ascontiguousarray(indices, dtype=np.uint32)
        self.verts = verts


class Empty:
    """euc `Empty` target (src/texture.rs:285-319): size [0, 0], writes dropped."""
    handle = 0

    def size(self):
        return [0, 0]


# ---- context -------------------------------------------------------------------------------------------
class Context:
    """One CUDA device.  Single-owner: use from one host thread at a time."""

    def __init__(self, device=0):
        self._lib = load()
        p = C.c_void_p()
        rc = self._lib.euc_init(int(device), C.byref(p))
        if rc != abi.OK:
            raise EucError(rc, f"euc_init(device={device}) failed (is a CUDA device present? there is no CPU fallback)")
        self._p = p
        self.device = device

    def _check(self, rc):
        if rc != abi.OK:
            raise EucError(rc, self._lib.euc_last_error(self._p).decode())

    def set_stream(self, cuda_stream):
        self._check(self._lib.euc_set_stream(self._p, C.c_void_p(int(cuda_stream) if cuda_stream else 0)))

    def sync(self):
        self._check(self._lib.euc_sync(self._p))

    def render_clear(self, pixel_value=None, depth_value=None):
        """The next render of this context first clears its pixel target to `pixel_value` (u32) and its depth target to
        `depth_value` (f32), fused into the render's kernels (euc_render_clear)."""
        pv = C.c_uint32(int(pixel_value) & 0xFFFFFFFF) if pixel_value is not None else None
        zv = C.c_float(float(depth_value)) if depth_value is not None else None
        self._check(self._lib.euc_render_clear(self._p, C.byref(pv) if pv is not None else None, C.byref(zv) if zv is not None else None))

    def ticket_wait(self, ticket):
        self._check(self._lib.euc_ticket_wait(self._p, int(ticket)))

    def host_alloc(self, nbytes):
        """Pinned host memory (euc_host_alloc); returns its address.  Freed with host_free."""
        p = C.c_void_p()
        self._check(self._lib.euc_host_alloc(self._p, int(nbytes), C.byref(p)))
        return p.value

    def host_free(self, ptr):
        self._check(self._lib.euc_host_free(self._p, C.c_void_p(ptr)))

    def set_async(self, enabled):
        """euc_set_async: renders never wait for the device (default on); False = every render call is checked."""
        self._check(self._lib.euc_set_async(self._p, 1 if enabled else 0))

    def blocking_waits(self):
        return int(self._lib.euc_blocking_waits(self._p))

    def set_stats(self, enabled):
        self._check(self._lib.euc_set_stats(self._p, 1 if enabled else 0))

    def get_stats(self):
        s = abi.RenderStats()
        self._check(self._lib.euc_get_stats(self._p, C.byref(s)))
        return {"primitives": s.primitives, "binned_pairs": s.binned_pairs, "fragments": s.fragments}

    def set_profiling(self, enabled):
        self._check(self._lib.euc_set_profiling(self._p, 1 if enabled else 0))

    def get_profile(self, reset=True):
        """{stage: (milliseconds, launches)} accumulated since the last reset (blocking)."""
        n = len(abi.STAGE_NAMES)
        ms, calls = (C.c_float * n)(), (C.c_uint64 * n)()
        self._check(self._lib.euc_get_profile(self._p, ms, calls, 1 if reset else 0))
        return {abi.STAGE_NAMES[i]: (float(ms[i]), int(calls[i])) for i in range(n)}

    def launch_count(self):
        return int(self._lib.euc_launch_count(self._p))

    def register_pipeline(self, source: str, struct_name: str) -> int:
        """Compile a user-written pipeline (CUDA source with the static interface of csrc/shaders.cuh) at run time with
        NVRTC and return its pipeline id for this context.  The device analogue of `impl Pipeline for MyShader`."""
        out = C.c_int32()
        rc = self._lib.euc_pipeline_register(self._p, source.encode(), struct_name.encode(), C.byref(out))
        if rc != abi.OK:
            raise EucError(rc, self._lib.euc_last_error(self._p).decode() + "\n" + self._lib.euc_pipeline_log(self._p).decode())
        return out.value

    def close(self):
        if getattr(self, "_p", None):
            self._lib.euc_shutdown(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ---- Buffer2d ------------------------------------------------------------------------------------------
class Sampler:
    """texture.linear()/nearest() + .clamped()/.tiled()/.mirrored().  `texture` is a device Buffer2d (product path)
    or a numpy array (oracle path)."""

    def __init__(self, texture, fmt, filt, wrap=abi.WRAP_NONE):
        self.texture, self.format, self.filter, self.wrap = texture, fmt, filt, wrap

    def clamped(self):
        return Sampler(self.texture, self.format, self.filter, abi.WRAP_CLAMP)

    def tiled(self):
        return Sampler(self.texture, self.format, self.filter, abi.WRAP_TILE)

    def mirrored(self):
        return Sampler(self.texture, self.format, self.filter, abi.WRAP_MIRROR)


class Buffer2d:
    """Device-resident Buffer2d<T> with 4-byte texels (u32 colour, f32 depth, RGBA8 texture), row-major x + w*y.
    `layers` > 1 makes an array of equally sized targets (batch rendering)."""

    def __init__(self, size, dtype, ctx=None, layers=1, wrap_ptr=None):
        self.ctx = ctx or default_context()
        self.dtype = np.dtype(dtype)
        assert self.dtype.itemsize == 4
        self._size = [int(size[0]), int(size[1])]
        self.layers = int(layers)
        h = C.c_uint64()
        if wrap_ptr is None:
            self.ctx._check(self.ctx._lib.euc_buf_create(self.ctx._p, self._size[0], self._size[1], self.layers, 4, C.byref(h)))
        else:  # caller-owned device memory (e.g. a torch tensor's storage); the caller keeps it alive
            self.ctx._check(self.ctx._lib.euc_buf_wrap(self.ctx._p, C.c_void_p(int(wrap_ptr)), self._size[0], self._size[1],
                                                       self.layers, 4, C.byref(h)))
        self.handle = h.value

    @classmethod
    def wrap(cls, device_ptr, size, dtype, ctx=None, layers=1):
        return cls(size, dtype, ctx, layers, wrap_ptr=device_ptr)

    @classmethod
    def fill(cls, size, item, dtype=None, ctx=None, layers=1):
        """Buffer2d::fill(size, item) (src/buffer.rs:60-67)."""
        if dtype is None:
            dtype = np.float32 if isinstance(item, float) else np.uint32
        b = cls(size, dtype, ctx, layers)
        b.clear(item)
        return b

    @classmethod
    def from_array(cls, arr, ctx=None):
        """Upload a host image.  (h, w) of 4-byte texels, or (h, w, 4) uint8 RGBA."""
        arr = np.ascontiguousarray(arr)
        if arr.ndim == 3 and arr.dtype == np.uint8 and arr.shape[2] == 4:
            arr = arr.view(np.uint32).reshape(arr.shape[0], arr.shape[1])
        b = cls([arr.shape[1], arr.shape[0]], arr.dtype, ctx)
        b.upload(arr)
        return b

    def size(self):
        return list(self._size)

    def _texel(self, texel):
        return C.c_float(texel) if self.dtype == np.float32 else C.c_uint32(int(texel) & 0xFFFFFFFF)

    def clear(self, texel):
        """Target::clear (src/buffer.rs:213-218)."""
        rc = self.ctx._lib.euc_buf_clear(self.ctx._p, self.handle, C.byref(self._texel(texel)))
        if rc:
            self.ctx._check(rc)

    def clear_rows(self, texel, row_begin, row_end):
        """Target::clear restricted to rows [row_begin, row_end) (row-band rendering across ranks)."""
        rc = self.ctx._lib.euc_buf_clear_rows(self.ctx._p, self.handle, C.byref(self._texel(texel)), int(row_begin), int(row_end))
        if rc:
            self.ctx._check(rc)

    def as_torch(self):
        """Zero-copy torch view (int32/float32, flat) of this buffer through __cuda_array_interface__."""
        import torch
        ptr, nbytes = self.device_ptr()

        class _CAI:
            pass
        o = _CAI()
        o.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4" if self.dtype == np.float32 else "<i4",
                                      "data": (ptr, False), "version": 3}
        return torch.as_tensor(o, device=f"cuda:{self.ctx.device}")

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        self.ctx._check(self.ctx._lib.euc_buf_upload(self.ctx._p, self.handle, arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def raw(self, out=None):
        """Buffer::raw() (src/buffer.rs:104-107): downloads; shape (layers, h, w) squeezed to (h, w) for one layer."""
        shape = (self.layers, self._size[1], self._size[0]) if self.layers > 1 else (self._size[1], self._size[0])
        if out is None:
            out = np.empty(shape, dtype=self.dtype)
        self.ctx._check(self.ctx._lib.euc_buf_download(self.ctx._p, self.handle, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def download_async(self, host_ptr, rows=None):
        """Asynchronous read-back into (pinned) host memory at `host_ptr`: the whole buffer, or rows (r0, r1) of a single-layer
        buffer, tightly packed.  Returns a ticket for Context.ticket_wait."""
        t = C.c_uint64()
        if rows is None:
            self.ctx._check(self.ctx._lib.euc_buf_download_async(self.ctx._p, self.handle, C.c_void_p(host_ptr), self._size[0] * self._size[1] * self.layers * 4, C.byref(t)))
        else:
            self.ctx._check(self.ctx._lib.euc_buf_download_rows_async(self.ctx._p, self.handle, C.c_void_p(host_ptr), int(rows[0]), int(rows[1]), C.byref(t)))
        return t.value

    def ipc_export(self) -> bytes:
        """CUDA IPC handle of this buffer (for another process on the same node)."""
        h = C.create_string_buffer(abi.IPC_HANDLE_BYTES)
        self.ctx._check(self.ctx._lib.euc_buf_ipc_export(self.ctx._p, self.handle, h))
        return h.raw

    @classmethod
    def ipc_import(cls, handle: bytes, size, dtype, ctx=None, layers=1):
        """Map a buffer exported by another process; stores into it travel over NVLink."""
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        self.dtype = np.dtype(dtype)
        self._size = [int(size[0]), int(size[1])]
        self.layers = int(layers)
        out = C.c_uint64()
        hb = C.create_string_buffer(handle, abi.IPC_HANDLE_BYTES)
        self.ctx._check(self.ctx._lib.euc_buf_ipc_import(self.ctx._p, hb, self._size[0], self._size[1], self.layers, 4, C.byref(out)))
        self.handle = out.value
        return self

    def device_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        self.ctx._check(self.ctx._lib.euc_buf_device_ptr(self.ctx._p, self.handle, C.byref(p), C.byref(n)))
        return p.value, n.value

    # Texture::linear / nearest (src/texture.rs:51-95); `fmt` tells how a texel is mapped on read (Map, :140-176)
    def linear(self, fmt=None):
        return Sampler(self, self._fmt(fmt), abi.FILTER_LINEAR)

    def nearest(self, fmt=None):
        return Sampler(self, self._fmt(fmt), abi.FILTER_NEAREST)

    def _fmt(self, fmt):
        if fmt is not None:
            return fmt
        return abi.TEXEL_F32 if self.dtype == np.float32 else abi.TEXEL_RGBA8_TO_F32

    def destroy(self):
        if self.handle:
            self.ctx._lib.euc_buf_destroy(self.ctx._p, self.handle)
            self.handle = 0

    def __del__(self):
        try:
            if self.ctx._p:
                self.destroy()
        except Exception:
            pass


class Geometry:
    """Device-resident vertex (+ optional u32 index) buffer."""

    def __init__(self, vertices, indices=None, ctx=None):
        self.ctx = ctx or default_context()
        v = np.ascontiguousarray(vertices)
        self.n_vertices = v.shape[0]
        self.stride = v.dtype.itemsize if v.ndim == 1 else v.strides[0]
        idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.uint32)
        self.n_indices = 0 if idx is None else idx.size
        self.stream_len = self.n_indices if idx is not None else self.n_vertices
        h = C.c_uint64()
        self.ctx._check(self.ctx._lib.euc_geom_create(
            self.ctx._p, v.ctypes.data_as(C.c_void_p), self.stride, self.n_vertices,
            idx.ctypes.data_as(C.c_void_p) if idx is not None else None, self.n_indices, C.byref(h)))
        self.handle = h.value

    @classmethod
    def wrap(cls, vertices_ptr, stride, n_vertices, indices_ptr=None, n_indices=0, ctx=None):
        """Wrap caller-owned device memory (e.g. torch tensors a collective fills) as a Geometry."""
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        self.n_vertices, self.stride, self.n_indices = int(n_vertices), int(stride), int(n_indices) if indices_ptr else 0
        self.stream_len = self.n_indices if indices_ptr else self.n_vertices
        h = C.c_uint64()
        self.ctx._check(self.ctx._lib.euc_geom_wrap(self.ctx._p, C.c_void_p(int(vertices_ptr)), self.stride, self.n_vertices,
                                                    C.c_void_p(int(indices_ptr)) if indices_ptr else None, self.n_indices, C.byref(h)))
        self.handle = h.value
        return self

    def update(self, vertices_ptr, indices_ptr=None):
        """Re-upload from host memory (raw addresses, e.g. of pinned buffers); asynchronous on the context's stream."""
        self.ctx._check(self.ctx._lib.euc_geom_update(self.ctx._p, self.handle, C.c_void_p(vertices_ptr),
                                                      C.c_void_p(indices_ptr) if indices_ptr else None))

    def update_range(self, vertices_ptr, first_vertex, n_vertices, indices_ptr=None, first_index=0, n_indices=0):
        """Re-upload a slice (raw host addresses of the slice's first vertex / index); asynchronous when pinned."""
        self.ctx._check(self.ctx._lib.euc_geom_update_range(self.ctx._p, self.handle, C.c_void_p(vertices_ptr) if vertices_ptr else None,
                                                            int(first_vertex), int(n_vertices), C.c_void_p(indices_ptr) if indices_ptr else None,
                                                            int(first_index), int(n_indices)))

    def destroy(self):
        if self.handle:
            self.ctx._lib.euc_geom_destroy(self.ctx._p, self.handle)
            self.handle = 0

    def __del__(self):
        try:
            if self.ctx._p:
                self.destroy()
        except Exception:
            pass


# ---- Pipeline ------------------------------------------------------------------------------------------
class Pipeline:
    """Mirror of `trait Pipeline` (src/pipeline.rs:171-301).  Subclasses fix `pipeline_id`, the vertex layout and
    the uniform block; the mode getters have the reference's defaults and may be overridden per instance."""

    pipeline_id = -1
    vertex_dtype = None
    Primitives = TriangleList

    def pixel_mode(self):          # pipeline.rs:180-182
        return PixelMode.WRITE

    def depth_mode(self):          # pipeline.rs:186-188
        return DepthMode.NONE

    def coordinate_mode(self):     # pipeline.rs:192-194
        return CoordinateMode.VULKAN

    def aa_mode(self):             # pipeline.rs:198-200
        return AaMode.NONE

    def rasterizer_config(self):   # pipeline.rs:204-209 (CullMode default = Back)
        return CullMode.Back

    def uniform_block(self) -> bytes:
        return b""

    def samplers(self):
        return []

    # -- desc marshalling (shared by the device path and by the oracle wrapper) --
    def build_desc(self, sampler_handle_of):
        d = abi.PipelineDesc()
        dm, pm, cm, aa = self.depth_mode(), self.pixel_mode(), self.coordinate_mode(), self.aa_mode()
        d.pipeline_id = self.pipeline_id
        d.primitive_kind = self.Primitives.kind
        d.cull_mode = self.rasterizer_config()
        d.depth_test = _DEPTH_TEST[dm.test]
        d.depth_write = int(dm.write)
        d.pixel_write = int(pm.write)
        d.y_axis_up = int(cm.y_axis_direction == "Up")
        d.handedness = abi.HAND_LEFT if cm.handedness == "Left" else abi.HAND_RIGHT
        d.z_clip_enabled = int(cm.z_clip_range is not None)
        if cm.z_clip_range is not None:
            d.z_clip_min, d.z_clip_max = cm.z_clip_range
        d.msaa_level = aa.level
        ub = self.uniform_block()
        keep = C.create_string_buffer(ub, max(len(ub), 1))
        d.uniforms = C.cast(keep, C.c_void_p)
        d.uniform_bytes = len(ub)
        for i, s in enumerate(self.samplers()):
            d.samplers[i].buf = sampler_handle_of(s)
            d.samplers[i].format, d.samplers[i].filter, d.samplers[i].wrap = s.format, s.filter, s.wrap
        return d, keep

    def freeze(self):
        """Marshal the descriptor once and reuse it for every later render of this object (per-frame host overhead).
        The pipeline's fields must not change afterwards."""
        self._frozen = self.build_desc(lambda s: s.texture.handle)
        return self

    def render(self, vertices, pixel, depth, rows=None, mirrors=None, clear=None):
        """Pipeline::render (src/pipeline.rs:248).  `vertices`: numpy vertex array (stream), IndexedVertices, or a
        device-resident Geometry.  `pixel` / `depth`: Buffer2d or Empty().  Asynchronous on the context's stream.
        clear=(pixel_value | None, depth_value | None): same result as `pixel.clear(..); depth.clear(..)` (restricted to
        `rows`) followed by this render, with the clears fused into the render's kernels (euc_render_clear)."""
        ctx = None
        for t in (pixel, depth, vertices):
            if isinstance(t, (Buffer2d, Geometry)):
                ctx = t.ctx
                break
        ctx = ctx or default_context()
        d, keep = getattr(self, "_frozen", None) or self.build_desc(lambda s: s.texture.handle)
        lib = ctx._lib
        if rows is not None and not isinstance(vertices, Geometry):
            raise ValueError("row-restricted rendering needs a device-resident Geometry")
        if clear is not None:  # armed only once the arguments are known to be good: the request belongs to THIS render
            ctx.render_clear(*clear)
        if isinstance(vertices, Geometry):
            if mirrors:
                r0, r1 = rows if rows is not None else (0, 0xFFFFFFFF)
                arr = (C.c_uint64 * len(mirrors))(*[m.handle for m in mirrors])
                rc = lib.euc_render_geom_rows_mirrored(ctx._p, C.byref(d), vertices.handle, pixel.handle, depth.handle, r0, r1, arr, len(mirrors))
            elif rows is None:
                rc = lib.euc_render_geom(ctx._p, C.byref(d), vertices.handle, pixel.handle, depth.handle)
            else:
                rc = lib.euc_render_geom_rows(ctx._p, C.byref(d), vertices.handle, pixel.handle, depth.handle, rows[0], rows[1])
        else:
            if isinstance(vertices, IndexedVertices):
                v, idx = np.ascontiguousarray(vertices.verts), vertices.indices
            else:
                v, idx = np.ascontiguousarray(vertices), None
            rc = lib.euc_render(ctx._p, C.byref(d), v.ctypes.data_as(C.c_void_p), v.dtype.itemsize if v.ndim == 1 else v.strides[0],
                                v.shape[0], idx.ctypes.data_as(C.c_void_p) if idx is not None else None,
                                0 if idx is None else idx.size, pixel.handle, depth.handle)
        if rc:
            ctx._check(rc)

    def render_batch(self, geometry, draws, uniform_blocks, pixel, depth, clear=None):
        """n independent Pipeline::render calls in one launch sequence.  draws: iterable of (first, count,
        base_vertex, layer); uniform_blocks: bytes of len(draws) uniform blocks.  clear: as in render()."""
        ctx = geometry.ctx
        if clear is not None:
            ctx.render_clear(*clear)
        d, keep = self.build_desc(lambda s: s.texture.handle)
        # draws as one int64 -> 4 x 32-bit table (euc_batch_draw is four 32-bit words); a Python loop over ctypes structs
        # costs ~1 us per draw, which is the GPU time of a whole icon
        tbl = np.ascontiguousarray(np.asarray(draws, dtype=np.int64).reshape(-1, 4).astype(np.uint32))
        n = tbl.shape[0]
        ub = uniform_blocks if isinstance(uniform_blocks, (bytes, bytearray)) else bytes(uniform_blocks)
        if n:
            d.uniform_bytes = len(ub) // n
        ubuf = (C.c_char * max(len(ub), 1)).from_buffer_copy(ub) if len(ub) else C.create_string_buffer(1)
        ctx._check(ctx._lib.euc_render_batch(ctx._p, C.byref(d), geometry.handle, C.cast(tbl.ctypes.data, C.POINTER(abi.BatchDraw)), n,
                                             C.cast(ubuf, C.c_void_p), pixel.handle, depth.handle))
