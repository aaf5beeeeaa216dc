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

from echorenderer_b200 import host, scenes, shard_epochs, shard_tiles, structs  # noqa: E402

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
        frame[ty * TILE:(ty + 1) * TILE, tx * TILE:(tx + 1) * TILE, 3] = 1.0  # weight 1 per rendered pixel (render_frame_device uses the sample count)
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

        # block-cyclic dealing (what bench.py uses with the Hilbert sequence): still a partition, runs of `block` tiles stay together
        blocks = [shard_tiles(tiles, rank, world, block=4) for rank in range(world)]
        assert sorted(map(tuple, np.concatenate(blocks))) == sorted(map(tuple, tiles))
        assert max(len(p) for p in blocks) - min(len(p) for p in blocks) <= 4
        assert np.array_equal(blocks[0][:4], tiles[:4])

        # ownership by position (bench.py's default): a partition, every rank in every row of tiles, sequence order kept
        owned = [shard_tiles(tiles, rank, world, by="position") for rank in range(world)]
        assert sorted(map(tuple, np.concatenate(owned))) == sorted(map(tuple, tiles))
        assert max(len(p) for p in owned) - min(len(p) for p in owned) <= 5
        for rank, part in enumerate(owned):
            assert np.all((part[:, 0] + part[:, 1]) % world == rank)
            order = {tuple(t): i for i, t in enumerate(map(tuple, tiles))}
            assert [order[tuple(t)] for t in part] == sorted(order[tuple(t)] for t in part)


def test_two_rank_tile_sharding_matches_single_process(tmp_path):
    result = str(tmp_path / "frame.npy")
    mp.spawn(worker, args=(2, 29517, result), nprocs=2, join=True)
    reduced = np.load(result)
    single = render_frame(scenes.tile_grid(WIDTH, HEIGHT, TILE))[..., :3]
    assert np.array_equal(reduced.view(np.uint32), single.view(np.uint32))


EPOCHS, EXTEND = 4, 2


def render_epochs(first, count):
    """All tiles, epochs [first, first + count): (mean * count, count) per pixel, the layout render_frame_device accumulates."""
    from tests import oracle_lib as ol
    prepared = host.prepare(scenes.cornell_box())
    oracle = ol.OracleScene(prepared)
    tiles = scenes.tile_grid(WIDTH, HEIGHT, TILE)
    frame = np.zeros((HEIGHT, WIDTH, 4), dtype=np.float32)
    if count == 0:
        return frame
    params = structs.render_params(WIDTH, HEIGHT, TILE, extend=EXTEND, min_epoch=count, max_epoch=count, seed=4, epoch_offset=first)
    out, _ = oracle.render_tiles(params, tiles, threads=1)
    for tile, (tx, ty) in zip(out, tiles):
        frame[ty * TILE:(ty + 1) * TILE, tx * TILE:(tx + 1) * TILE, :3] = tile[..., :3] * count
        frame[ty * TILE:(ty + 1) * TILE, tx * TILE:(tx + 1) * TILE, 3] = count
    return frame


def sample_worker(rank, world, port, result_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_epochs(EPOCHS, rank, world)
    frame = torch.from_numpy(render_epochs(first, count))
    dist.all_reduce(frame)  # sum of (mean * epochs, epochs)
    resolved = frame[..., :3] / frame[..., 3:4]
    if rank == 0:
        np.save(result_path, resolved.numpy())
    dist.destroy_process_group()


def test_shard_epochs_partition():
    for world in (1, 2, 3, 4, 8):
        for epochs in (1, 4, 7, 20):
            blocks = [shard_epochs(epochs, rank, world) for rank in range(world)]
            covered = [e for first, count in blocks for e in range(first, first + count)]
            assert covered == list(range(epochs))


def test_two_rank_sample_sharding_matches_single_process(tmp_path):
    """every rank renders its block of epochs of all tiles; the sum of (mean * n, n) divided by n is the
    mean over all samples — equal to the single-process render of all epochs up to float rounding (a Welford mean over 8
    samples vs the average of two 4-sample means)."""
    result = str(tmp_path / "frame.npy")
    mp.spawn(sample_worker, args=(2, 29519, result), nprocs=2, join=True)
    reduced = np.load(result)
    single = render_epochs(0, EPOCHS)
    single = single[..., :3] / single[..., 3:4]
    assert np.allclose(reduced, single, rtol=2e-6, atol=1e-7)
