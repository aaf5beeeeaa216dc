"""The N > 1 path on CPU: world_size-2 gloo processes shard the tile list like EvaluationOperation's workers would and sum
their frames with one all-reduce — the host logic bench.py --workload render runs over NCCL. The per-tile "renderer" here is
the oracle (this is a test), so the reduced frame must equal a single-process render bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from echorenderer_b200 import host, scenes, shard_tiles, structs  # noqa: E402

WIDTH, HEIGHT, TILE = 48, 32, 16


def render_frame(tiles):
    from tests import oracle_lib as ol
    prepared = host.prepare(scenes.cornell_box())
    oracle = ol.OracleScene(prepared)
    params = structs.render_params(WIDTH, HEIGHT, TILE, extend=2, seed=4)
    out, _ = oracle.render_tiles(params, tiles, threads=1)
    frame = np.zeros((HEIGHT, WIDTH, 4), dtype=np.float32)
    for tile, (tx, ty) in zip(out, tiles):
        frame[ty * TILE:(ty + 1) * TILE, tx * TILE:(tx + 1) * TILE, :3] = tile[..., :3]
        frame[ty * TILE:(ty + 1) * TILE, tx * TILE:(tx + 1) * TILE, 3] = 1.0  # "rendered" weight, like render_frame_device
    return frame


def worker(rank, world, port, result_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tiles = shard_tiles(scenes.tile_grid(WIDTH, HEIGHT, TILE), rank, world)
    frame = torch.from_numpy(render_frame(tiles))
    dist.all_reduce(frame)  # sum of disjoint tiles
    resolved = frame[..., :3] / frame[..., 3:4]
    if rank == 0:
        np.save(result_path, resolved.numpy())
    dist.destroy_process_group()


def test_shard_tiles_partition():
    tiles = scenes.tile_grid(100, 70, 16)
    for world in (1, 2, 4, 8):
        parts = [shard_tiles(tiles, rank, world) for rank in range(world)]
        merged = np.concatenate(parts)
        assert len(merged) == len(tiles)
        assert sorted(map(tuple, merged)) == sorted(map(tuple, tiles))          # a partition: no tile lost or duplicated
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1     # balanced round-robin


def test_two_rank_tile_sharding_matches_single_process(tmp_path):
    result = str(tmp_path / "frame.npy")
    mp.spawn(worker, args=(2, 29517, result), nprocs=2, join=True)
    reduced = np.load(result)
    single = render_frame(scenes.tile_grid(WIDTH, HEIGHT, TILE))[..., :3]
    assert np.array_equal(reduced.view(np.uint32), single.view(np.uint32))
