"""Round 2: the 8-GPU C5 step under a sweep of the wavefront's host-side switches, in ONE torchrun session (the scene is built
and uploaded once per rank; echo_b200_debug_set_option changes the switches between configurations).

For every configuration: `--steps` epochs of 256 spp over the 3840x2160 frame, tiles sharded over the ranks, per step the host
wall time of the render call on every rank, the all-reduce time (CUDA events) and the device-timed total. Prints one JSON
line per configuration on rank 0 — where the step time goes at N ranks ("the unexplained ~90 ms").

  torchrun --nproc-per-node 8 variants/r2_sweep_c5.py --steps 3
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from echorenderer_b200 import PreparedScene, _native, host, scenes, structs, shard_tiles, hilbert_curve_pattern  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--steps", type=int, default=3)
    parser.add_argument("--spp", type=int, default=256)
    parser.add_argument("--scene", default="large")
    parser.add_argument("--width", type=int, default=3840)
    parser.add_argument("--height", type=int, default=2160)
    parser.add_argument("--bounce-limit", type=int, default=128)
    parser.add_argument("--quick", action="store_true", help="fewer configurations")
    parser.add_argument("--tail-sweep", action="store_true", help="sweep TAIL_LIMIT / NARROW_LIMIT instead of the host-side switches")
    parser.add_argument("--wait-sweep", action="store_true", help="sweep the wait mode (0 spin, 1 blocking event, 2 poll + yield) x run-ahead x pipelines, tiles owned by position")
    args = parser.parse_args()

    import torch
    import torch.distributed as dist
    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    builders = {"large": scenes.large_scene, "mixed": scenes.mixed_material_scene, "lights": scenes.many_lights_scene, "cornell": scenes.cornell_box}
    prepared = host.prepare(builders[args.scene]())
    scene = PreparedScene(prepared, device=local_rank)
    tile = 16
    count = ((args.width + tile - 1) // tile, (args.height + tile - 1) // tile)
    sequences = {"hilbert": hilbert_curve_pattern(count), "ordered": scenes.tile_grid(args.width, args.height, tile)}
    frame = torch.zeros(args.height * args.width * 4, dtype=torch.float32, device=device)
    stream = torch.cuda.current_stream().cuda_stream

    def gather(value):
        if world == 1:
            return [value]
        tensor = torch.tensor([value], dtype=torch.float64, device=device)
        out = [torch.zeros_like(tensor) for _ in range(world)]
        dist.all_gather(out, tensor)
        return [round(float(t.item()), 2) for t in out]

    def run(tag, pattern="hilbert", block=64, reduce_every=1, by="sequence", **switches):
        defaults = {"RUN_AHEAD": -1, "BLOCKING_SYNC": -1, "RENDER_WORKERS": 8, "BATCH_PATHS": 1 << 24, "TAIL_LIMIT": 8192, "NARROW_LIMIT": 1048576}
        for name, value in {**defaults, **switches}.items():
            _native.set_option(name, value)
        tiles = shard_tiles(sequences[pattern], rank, world, block=block, by=by)
        render_ms, reduce_ms = [], []
        for index in range(args.steps + 1):
            if index == 1:
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                start.record()
            params = structs.render_params(args.width, args.height, tile, extend=args.spp, min_epoch=1, max_epoch=1, bounce_limit=args.bounce_limit, seed=1, epoch_offset=index)
            begin = time.perf_counter()
            stats = scene.render_frame_device(params, tiles, frame.data_ptr(), stream)
            if index > 0:
                render_ms.append((time.perf_counter() - begin) * 1e3)
            if index % reduce_every == 0:
                first, last = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                first.record()
                if world > 1:
                    dist.all_reduce(frame)
                last.record()
                scene.frame_resolve_device(frame.data_ptr(), args.width, args.height, stream)
                frame.zero_()
                if index > 0:
                    reduce_ms.append((first, last))
        stop.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        per_rank_mean = gather(float(np.mean(render_ms)))
        per_rank_max = gather(float(np.max(render_ms)))
        reduce_mean = float(np.mean([a.elapsed_time(b) for a, b in reduce_ms])) if reduce_ms else 0.0
        if rank == 0:
            samples = args.width * args.height * args.spp * args.steps
            print(json.dumps({"tag": tag, "world": world, "ms_per_step": round(float(total.item()) / args.steps, 2), "msamples_per_s": round(samples / (float(total.item()) * 1e-3) / 1e6, 1),
                              "rank_render_ms_mean": per_rank_mean, "rank_render_ms_max": per_rank_max, "all_reduce_ms": round(reduce_mean, 3), "launches_rank0": int(stats["kernelLaunches"][0]),
                              "pattern": pattern, "block": block, "by": by, "reduce_every": reduce_every, "switches": switches}), flush=True)

    if args.tail_sweep:
        run("warm-up (discard)", by="position")
        for tail in (4096, 16384):
            run(f"tail limit {tail}", by="position", TAIL_LIMIT=tail)
        for narrow in (262144, 2097152, 4194304):
            run(f"narrow limit {narrow}", by="position", NARROW_LIMIT=narrow)
        run("default again", by="position")
        scene.close()
        if world > 1:
            dist.destroy_process_group()
        return

    if args.wait_sweep:
        run("warm-up (discard)", by="position")
        run("automatic", by="position")
        for wait, name in ((1, "blocking event"), (2, "poll + yield"), (0, "spinning event")):
            for ahead in (0, 1, 2):
                run(f"{name}, run-ahead {ahead}", by="position", BLOCKING_SYNC=wait, RUN_AHEAD=ahead)
        run("poll + yield, lock step, 4 pipelines", by="position", BLOCKING_SYNC=2, RUN_AHEAD=0, RENDER_WORKERS=4)
        run("poll + yield, lock step, 6 pipelines", by="position", BLOCKING_SYNC=2, RUN_AHEAD=0, RENDER_WORKERS=6)
        run("spinning event, lock step, 4 pipelines", by="position", BLOCKING_SYNC=0, RUN_AHEAD=0, RENDER_WORKERS=4)
        run("poll + yield, lock step, 8 Mi-path batches", by="position", BLOCKING_SYNC=2, RUN_AHEAD=0, BATCH_PATHS=1 << 23)
        run("automatic again", by="position")
        scene.close()
        if world > 1:
            dist.destroy_process_group()
        return

    run("warm-up (discard)")
    run("hilbert, tiles owned by position", by="position")
    run("default (automatic run-ahead / blocking)")
    run("lock step, blocking waits", RUN_AHEAD=0, BLOCKING_SYNC=1)
    run("lock step, spinning waits", RUN_AHEAD=0, BLOCKING_SYNC=0)
    run("run-ahead 2, blocking waits", RUN_AHEAD=2, BLOCKING_SYNC=1)
    run("run-ahead 2, spinning waits", RUN_AHEAD=2, BLOCKING_SYNC=0)
    run("reduce every 4 steps", reduce_every=4)
    if not args.quick:
        run("run-ahead 4, blocking waits", RUN_AHEAD=4, BLOCKING_SYNC=1)
        run("4 pipelines", RENDER_WORKERS=4)
        run("16 pipelines", RENDER_WORKERS=16)
        run("8 Mi-path batches", BATCH_PATHS=1 << 23)
        run("ordered sequence, tile by tile", pattern="ordered", block=1)
        run("hilbert, tile by tile", pattern="hilbert", block=1)
        run("hilbert, blocks of 256", pattern="hilbert", block=256)
        run("default again")

    scene.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
