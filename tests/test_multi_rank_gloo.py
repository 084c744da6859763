"""world_size-2 test of the multi-GPU host logic on CPU (gloo): each rank renders its row band with the CPU oracle,
the colour rows are all-gathered into one buffer exactly as bench.py does over NCCL, and the result must equal the
single-rank frame bit for bit.  Also checks the frame-shard partition used for icon batches."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import euc_b200 as e
    from euc_b200 import parallel, scenes
    from oracle import oracle
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h = 640, 200  # 13 tile rows -> 7 + 6; euc bands are 31 rows, so the split cuts through a band
    verts, idx = scenes.blend_tris(600, w, h, seed=11, size_px=(4.0, 30.0))
    slot_rows, bands = parallel.row_band_slots(h, world)
    r0, r1 = bands[rank]
    gather = torch.zeros(world * slot_rows * w, dtype=torch.int32)
    color = gather.numpy().view(np.uint32)[: h * w].reshape(h, w)
    color[:] = 0xFF000000
    depth = np.full((h, w), 1.0, np.float32)
    # the oracle renders whole euc bands that intersect [r0, r1); rows outside the slot are masked out afterwards,
    # which is what the device's row-restricted render produces directly
    tmp_c = np.full((h, w), 0xFF000000, np.uint32)
    oracle.render(e.BlendTris(), e.IndexedVertices(idx, verts), tmp_c, depth, n_threads=1, rows=(r0, r1))
    color[r0:r1] = tmp_c[r0:r1]
    mine = gather[rank * slot_rows * w:(rank + 1) * slot_rows * w].clone()
    dist.all_gather_into_tensor(gather, mine)
    if rank == 0:
        np.save(out_path, gather.numpy().view(np.uint32)[: h * w].reshape(h, w))
    # frame shards agree across ranks
    shards = parallel.frame_shards(4096, world)
    t = torch.tensor([shards[rank][1] - shards[rank][0]])
    dist.all_reduce(t)
    assert int(t) == 4096
    dist.barrier()
    dist.destroy_process_group()


def test_row_bands_all_gather_equals_single_rank(tmp_path):
    sys.path.insert(0, ROOT)
    import euc_b200 as e
    from euc_b200 import scenes
    from oracle import oracle
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    w, h = 640, 200
    verts, idx = scenes.blend_tris(600, w, h, seed=11, size_px=(4.0, 30.0))
    full = np.full((h, w), 0xFF000000, np.uint32)
    oracle.render(e.BlendTris(), e.IndexedVertices(idx, verts), full, np.full((h, w), 1.0, np.float32))
    got = np.load(out)
    assert (full != 0xFF000000).sum() > 5000
    assert np.array_equal(got, full)


def test_partitions():
    from euc_b200 import parallel
    slot, bands = parallel.row_band_slots(2160, 8)
    assert slot == 272 and bands[0] == (0, 272) and bands[-1] == (1904, 2160)
    assert all(b[0] % 16 == 0 for b in bands) and sum(b[1] - b[0] for b in bands) == 2160
    slot, bands = parallel.row_band_slots(480, 4)
    assert 4 * slot >= 480 and bands[-1][1] == 480
    assert parallel.frame_shards(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert parallel.row_band_slots(100, 1) == (112, [(0, 100)])
