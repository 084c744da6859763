"""Host I/O around `Pipeline::render`: the steps the reference's examples take on either side of the
render call.

  load_obj            wavefront::Obj::from_reader + Obj::vertices()       (benches/teapot.rs:151-152, :189)
  texture_from_image  image::open(..).to_rgba8() + Buffer2d::from_texture  (examples/texture_mapping.rs:111-116, src/buffer.rs:32-58)
  ReadbackRing        a pinned-memory ring for `win.update_with_buffer(color.raw(), ..)`-style consumers
                      (examples/teapot.rs:224): read-back of frame i overlaps rendering of frame i+1
"""
import ctypes as C

import numpy as np

from . import abi
from .core import Buffer2d, default_context
from .pipelines import VERTEX_PN


def load_obj(path_or_file):
    """Stream of `wavefront::Vertex` (position + normal) in face order, three face-vertices per triangle; polygons with
    more than three vertices are fan-triangulated.  Faces without normals get a zero normal (the reference's bench
    unwraps the normal and would panic)."""
    f = open(path_or_file) if isinstance(path_or_file, (str, bytes)) else path_or_file
    pos, nrm, out = [], [], []
    with f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                pos.append([float(x) for x in t[1:4]])
            elif t[0] == "vn":
                nrm.append([float(x) for x in t[1:4]])
            elif t[0] == "f":
                corners = []
                for c in t[1:]:
                    parts = c.split("/")
                    vi = int(parts[0])
                    ni = int(parts[2]) if len(parts) > 2 and parts[2] else 0
                    corners.append((vi - 1 if vi > 0 else len(pos) + vi, (ni - 1 if ni > 0 else len(nrm) + ni) if ni else -1))
                for k in range(1, len(corners) - 1):
                    out.extend([corners[0], corners[k], corners[k + 1]])
    pos = np.asarray(pos, dtype=np.float32).reshape(-1, 3)
    nrm = np.asarray(nrm, dtype=np.float32).reshape(-1, 3)
    stream = np.zeros(len(out), dtype=VERTEX_PN)
    idx = np.asarray(out, dtype=np.int64).reshape(-1, 2)
    if len(out):
        stream["pos"] = pos[idx[:, 0]]
        has_n = idx[:, 1] >= 0
        if has_n.any():
            stream["normal"][has_n] = nrm[idx[has_n, 1]]
    return stream


def texture_from_image(image, ctx=None):
    """`Buffer2d::from_texture(&image::open(path).to_rgba8())`: a PIL image, a path, or an (h, w, 3|4) uint8 array becomes
    a device RGBA8 texture; sample it with `.linear()` / `.nearest()` (texels map to f32 0..255 on read)."""
    if isinstance(image, (str, bytes)):
        from PIL import Image
        image = Image.open(image)
    if hasattr(image, "convert"):
        image = np.asarray(image.convert("RGBA"), dtype=np.uint8)
    image = np.ascontiguousarray(image, dtype=np.uint8)
    if image.ndim == 3 and image.shape[2] == 3:
        image = np.concatenate([image, np.full(image.shape[:2] + (1,), 255, np.uint8)], axis=2)
    assert image.ndim == 3 and image.shape[2] == 4
    return Buffer2d.from_array(image, ctx)


class ReadbackRing:
    """`depth` pinned host frames.  `submit(buf)` queues an asynchronous read-back of a device buffer behind the work
    already issued on the context's stream and returns a slot; `wait(slot)` blocks until that frame is in host memory
    and returns it as a numpy array (valid until the slot is reused)."""

    def __init__(self, buf: Buffer2d, depth=3, ctx=None):
        self.ctx = ctx or buf.ctx or default_context()
        self.shape = (buf.layers, buf._size[1], buf._size[0]) if buf.layers > 1 else (buf._size[1], buf._size[0])
        self.dtype, self.nbytes = buf.dtype, int(np.prod(self.shape)) * 4
        self._ptrs, self._views, self._tickets, self._next = [], [], [None] * depth, 0
        for _ in range(depth):
            p = C.c_void_p()
            self.ctx._check(self.ctx._lib.euc_host_alloc(self.ctx._p, self.nbytes, C.byref(p)))
            self._ptrs.append(p)
            self._views.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(self.nbytes // 4,)).view(self.dtype).reshape(self.shape))

    def submit(self, buf: Buffer2d) -> int:
        slot = self._next
        self._next = (self._next + 1) % len(self._ptrs)
        if self._tickets[slot] is not None:
            self.wait(slot)
        t = C.c_uint64()
        self.ctx._check(self.ctx._lib.euc_buf_download_async(self.ctx._p, buf.handle, self._ptrs[slot], self.nbytes, C.byref(t)))
        self._tickets[slot] = t.value
        return slot

    def wait(self, slot: int) -> np.ndarray:
        if self._tickets[slot] is not None:
            self.ctx._check(self.ctx._lib.euc_ticket_wait(self.ctx._p, self._tickets[slot]))
            self._tickets[slot] = None
        return self._views[slot]

    def close(self):
        for i, p in enumerate(self._ptrs):
            if self._tickets[i] is not None:
                self.wait(i)
            self.ctx._lib.euc_host_free(self.ctx._p, p)
        self._ptrs, self._views = [], []
