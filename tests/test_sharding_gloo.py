"""Multi-GPU host logic on CPU: world_size-2 gloo processes agree on a disjoint, complete tile-row partition
(the same bvht_shard_tile_rows the launches use), and a frame assembled from per-rank row sets equals the
full frame (rendered here by the oracle, since there is no GPU)."""
import os
import socket
import sys

import numpy as np
import pytest
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


def _worker(rank, world, port, height, tile, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bvhtracer_b200 import _ffi, examples
    import oracle_lib as O
    import scene_build as SB
    width = 96
    rows = _ffi.shard_tile_rows((0, 0, width, height), tile, rank, world)
    n_rows_total = (height + tile - 1) // tile
    owned = torch.zeros(n_rows_total, dtype=torch.int32)
    owned[rows] = 1
    dist.all_reduce(owned)                                   # every row owned exactly once
    ok_partition = bool((owned == 1).all())
    # each rank renders only its rows into a zeroed frame; the sum over ranks must be the full frame
    scene, cam = SB.oracle_scene(examples.cube())
    part = np.zeros(width * height, O.HIT)
    for r in rows:
        scene.render(cam, width, height, tile=tile, region=(0, r * tile, width, min(height, (r + 1) * tile)), out=part)
    t = torch.from_numpy(part.view(np.uint8).reshape(-1).astype(np.int32))
    dist.all_reduce(t)
    if rank == 0:
        full = scene.render(cam, width, height, tile=tile)
        ok_frame = bool(np.array_equal(t.numpy().astype(np.uint8), full.view(np.uint8).reshape(-1)))
        ret.put((ok_partition, ok_frame, len(rows)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("height,tile", [(64, 8), (70, 8), (45, 5)])
def test_two_rank_tile_row_partition(height, tile):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, height, tile, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok_partition, ok_frame, n0 = ret.get(timeout=5)
    assert ok_partition and ok_frame and n0 > 0


def test_shard_rows_properties():
    from bvhtracer_b200 import _ffi
    for (y0, y1, tile, n) in [(0, 2160, 8, 8), (16, 2000, 8, 3), (7, 130, 5, 4), (0, 8, 8, 4), (40, 40, 8, 2)]:
        rows = [_ffi.shard_tile_rows((0, y0, 64, y1), tile, i, n) for i in range(n)]
        flat = sorted(r for rr in rows for r in rr)
        expect = list(range(y0 // tile, (y1 + tile - 1) // tile)) if y1 > y0 else []
        assert flat == expect
        assert all(r % n == i for i, rr in enumerate(rows) for r in rr)
        assert max(len(rr) for rr in rows) - min(len(rr) for rr in rows) <= 1
