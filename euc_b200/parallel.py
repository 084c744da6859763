"""Work partitioning across ranks (one process per GPU).  Pure host logic: no device calls here.

Two partitions, as SURVEY §8(e):
  * independent frames (icon batches): contiguous frame ranges per rank, no data-path collective;
  * one large frame: contiguous screen-space row bands per rank, then an all-gather of the colour rows.  euc's own
    row bands are independent of each other (src/pipeline.rs:348-350) and the raster kernel derives every
    band-dependent quantity per row, so any 16-row-aligned split reproduces the single-GPU frame bit for bit.
"""
TILE = 16


def row_band_slots(height, world):
    """Even split of the 16-px tile rows.  Returns (slot_rows, [(row_begin, row_end)] per rank); every rank owns a slot of
    slot_rows rows in the gather buffer (world * slot_rows >= height), the last ranks may own fewer (or zero) real rows."""
    tile_rows = (height + TILE - 1) // TILE
    per = (tile_rows + world - 1) // world
    slot_rows = per * TILE
    return slot_rows, [(min(r * slot_rows, height), min((r + 1) * slot_rows, height)) for r in range(world)]


def frame_shards(n_frames, world):
    """Contiguous [begin, end) frame ranges, sizes differing by at most one."""
    base, extra = divmod(n_frames, world)
    out, b = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((b, b + n))
        b += n
    return out
