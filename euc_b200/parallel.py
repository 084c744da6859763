"""Work partitioning and the multi-GPU group, one process (or host thread) per GPU of one node.

Two partitions, as SURVEY §8(e):
  * independent frames (icon batches): contiguous frame ranges per rank, no data-path collective;
  * one large frame: contiguous screen-space row bands per rank; the band's colour rows are stored by the raster kernel
    into the root's framebuffer over NVLink (gather), or into every peer's (all-gather).  euc's own row bands are
    independent of each other (src/pipeline.rs:340-362) and the raster kernel derives every band-dependent quantity per
    row, so any 16-row-aligned split reproduces the single-GPU frame bit for bit.

The partition arithmetic lives in the C ABI (euc_group_rows / euc_group_frames, pure functions); `Group` wraps the
collective entry points (euc_group_create / share_buf / barrier / render).  Nothing here needs torch or NCCL.
"""
import ctypes as C

from . import abi
from ._lib import load

TILE = 16


def row_band(height, rank, world):
    """(row_begin, row_end) of `rank`: tile rows split evenly; ranks beyond the last tile row get the empty band (0, 0)."""
    a, b = C.c_uint32(), C.c_uint32()
    rc = load().euc_group_rows(int(height), int(rank), int(world), C.byref(a), C.byref(b))
    if rc != abi.OK:
        raise ValueError(f"euc_group_rows({height}, {rank}, {world}) -> {rc}")
    return a.value, b.value


def row_band_slots(height, world):
    """Even split of the 16-px tile rows.  Returns (slot_rows, [(row_begin, row_end)] per rank); every rank owns a slot of
    slot_rows rows in a gather buffer (world * slot_rows >= height), the last ranks may own fewer (or zero) real rows."""
    tile_rows = (height + TILE - 1) // TILE
    per = (tile_rows + world - 1) // world
    return per * TILE, [row_band(height, r, world) for r in range(world)]


def frame_shards(n_frames, world):
    """Contiguous [begin, end) frame ranges, sizes differing by at most one."""
    out = []
    lib = load()
    for r in range(world):
        a, b = C.c_uint32(), C.c_uint32()
        rc = lib.euc_group_frames(int(n_frames), r, int(world), C.byref(a), C.byref(b))
        if rc != abi.OK:
            raise ValueError(f"euc_group_frames({n_frames}, {r}, {world}) -> {rc}")
        out.append((a.value, b.value))
    return out


class Group:
    """The ranks of one job.  Collective calls must be made by every rank, in the same order."""

    def __init__(self, ctx, name, rank, world):
        self.ctx, self.rank, self.world = ctx, int(rank), int(world)
        ctx._check(ctx._lib.euc_group_create(ctx._p, str(name).encode(), self.rank, self.world))

    def share(self, buf):
        """Every rank passes its buffer (same size); returns the list of all ranks' buffers as mapped here (own rank: buf)."""
        from .core import Buffer2d
        out = (C.c_uint64 * self.world)()
        self.ctx._check(self.ctx._lib.euc_group_share_buf(self.ctx._p, buf.handle, out))
        peers = []
        for r in range(self.world):
            if r == self.rank:
                peers.append(buf)
            else:
                b = Buffer2d.__new__(Buffer2d)
                b.ctx, b.dtype, b._size, b.layers, b.handle = self.ctx, buf.dtype, list(buf._size), buf.layers, out[r]
                peers.append(b)
        return peers

    def barrier(self):
        """Stream-ordered device barrier (no host wait)."""
        rc = self.ctx._lib.euc_group_barrier(self.ctx._p)
        if rc:
            self.ctx._check(rc)

    def render(self, pipeline, geometry, pixel_peers, depth, gather=abi.GATHER_ROOT, clear=None):
        """This rank's row band of one frame + gather of the colour rows (fused into the raster kernel) + barrier."""
        d, keep = getattr(pipeline, "_frozen", None) or pipeline.build_desc(lambda s: s.texture.handle)
        if clear is not None:
            self.ctx.render_clear(*clear)
        arr = getattr(self, "_peer_arr", None)
        if arr is None or self._peer_key != tuple(b.handle for b in pixel_peers):
            self._peer_key = tuple(b.handle for b in pixel_peers)
            arr = self._peer_arr = (C.c_uint64 * self.world)(*self._peer_key)
        rc = self.ctx._lib.euc_group_render(self.ctx._p, C.byref(d), geometry.handle, arr, depth.handle, int(gather))
        if rc:
            self.ctx._check(rc)

    def allgather_geom(self, geometry):
        """Every rank has uploaded its 1/world slice (Geometry.update_range over frame_shards of the vertex and index
        counts); the slices are exchanged over NVLink.  Collective, stream-ordered."""
        rc = self.ctx._lib.euc_group_allgather_geom(self.ctx._p, geometry.handle)
        if rc:
            self.ctx._check(rc)

    def rows(self, height):
        return row_band(height, self.rank, self.world)

    def close(self):
        if self.ctx is not None and getattr(self.ctx, "_p", None):
            self.ctx._lib.euc_group_destroy(self.ctx._p)
        self.ctx = None
