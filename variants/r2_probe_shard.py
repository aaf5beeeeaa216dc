"""Round 2 probe: where do the ~90 ms per rank of the 8-GPU C5 step come from (VERDICT weak #6)?

One GPU renders what rank `--rank` of `--world` would render of the C5 frame (tile i -> rank i mod world) at `--spp`
samples per pixel per step, with no other rank, no NCCL and no barrier. If this step is as slow as the step inside the
8-GPU job, the remainder is intra-rank (few, in-phase batches per pipeline); if it matches the 1/world share of the
one-GPU step, the remainder is between ranks. The prepared scene is cached in /tmp so that one gpurun call can run several
configurations (the library reads its environment switches once per process).

  python variants/r2_probe_shard.py --world 8 --spp 256 --steps 3 [--scene large]
"""
import argparse
import json
import os
import pickle
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from echorenderer_b200 import PreparedScene, host, scenes, structs, shard_tiles, hilbert_curve_pattern  # noqa: E402


def prepared_scene(name):
    path = f"/tmp/echo_probe_{name}.pkl"
    if os.path.exists(path):
        with open(path, "rb") as file:
            return pickle.load(file)
    builders = {"large": scenes.large_scene, "mixed": scenes.mixed_material_scene, "lights": scenes.many_lights_scene, "cornell": scenes.cornell_box}
    prepared = host.prepare(builders[name]())
    with open(path, "wb") as file:
        pickle.dump(prepared, file, protocol=4)
    return prepared


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--scene", default="large")
    parser.add_argument("--width", type=int, default=3840)
    parser.add_argument("--height", type=int, default=2160)
    parser.add_argument("--world", type=int, default=8)
    parser.add_argument("--rank", type=int, default=0)
    parser.add_argument("--spp", type=int, default=256)
    parser.add_argument("--steps", type=int, default=3)
    parser.add_argument("--bounce-limit", type=int, default=128)
    parser.add_argument("--pattern", default="ordered")
    parser.add_argument("--tag", default="")
    parser.add_argument("--by", default="sequence", choices=["sequence", "position"])
    args = parser.parse_args()

    import torch
    started = time.perf_counter()
    prepared = prepared_scene(args.scene)
    seconds_prepare = time.perf_counter() - started
    scene = PreparedScene(prepared, device=0)
    device = torch.device("cuda", 0)
    tile = 16
    count = ((args.width + tile - 1) // tile, (args.height + tile - 1) // tile)
    all_tiles = hilbert_curve_pattern(count) if args.pattern == "hilbert" else scenes.tile_grid(args.width, args.height, tile)
    tiles = shard_tiles(all_tiles, args.rank, args.world, by=args.by)
    frame = torch.zeros(args.height * args.width * 4, dtype=torch.float32, device=device)
    stream = torch.cuda.current_stream().cuda_stream
    times, samples, launches = [], 0, 0

    for index in range(args.steps + 1):
        params = structs.render_params(args.width, args.height, tile, extend=args.spp, min_epoch=1, max_epoch=1, bounce_limit=args.bounce_limit, seed=1, epoch_offset=index)
        frame.zero_()
        torch.cuda.synchronize()
        begin = time.perf_counter()
        stats = scene.render_frame_device(params, tiles, frame.data_ptr(), stream)
        torch.cuda.synchronize()
        if index > 0:
            times.append((time.perf_counter() - begin) * 1e3)
            samples += int(stats["sampleEvaluated"][0])
            launches += int(stats["kernelLaunches"][0])

    line = {"tag": args.tag, "scene": args.scene, "world": args.world, "rank": args.rank, "spp": args.spp, "tiles": int(len(tiles)), "ms_per_step": float(np.mean(times)), "ms_steps": times,
            "msamples_per_s": samples / (sum(times) * 1e-3) / 1e6, "launches_per_step": launches / args.steps, "prepare_seconds": seconds_prepare,
            "env": {k: v for k, v in os.environ.items() if k.startswith("ECHO_B200_")}}
    print(json.dumps(line))
    scene.close()


if __name__ == "__main__":
    main()
