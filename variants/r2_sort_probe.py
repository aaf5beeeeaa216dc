"""Round 2: what would sorting the ray queue buy? The secondary-ray batch of the C2 bench (rays leaving the surfaces the first
batch hit, cosine-hemisphere directions, 17 node visits per ray) is traced as spawned and after a HOST-side sort by (a) the Morton code of the
origin cell, (b) direction octant, then origin cell — an upper bound for a device-side sort, whose own cost is not counted. Results are the same
rays, so hits are identical up to their order.

  python variants/r2_sort_probe.py [--rays 16777216]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from echorenderer_b200 import PreparedScene, scenes, structs  # noqa: E402


def spread(v):
    v = v.astype(np.uint64) & np.uint64(0x3FF)
    v = (v | (v << np.uint64(16))) & np.uint64(0x30000FF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x300F00F)
    v = (v | (v << np.uint64(4))) & np.uint64(0x30C30C3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x9249249)
    return v


def morton(points, low, high, bits=10):
    scale = (1 << bits) - 1
    q = np.clip((points - low) / np.maximum(high - low, 1e-30) * scale, 0, scale).astype(np.uint64)
    return (spread(q[:, 0]) << np.uint64(2)) | (spread(q[:, 1]) << np.uint64(1)) | spread(q[:, 2])


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--rays", type=int, default=1 << 24)
    parser.add_argument("--steps", type=int, default=5)
    args = parser.parse_args()

    import torch
    device = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    inputs = argparse.Namespace(rays=args.rays, quads=[1000, 500], instanced=False, tree="sah")
    prepared, rays, _ = bench.build_trace_inputs(inputs, 0)
    scene = PreparedScene(prepared, device=0)
    hits = scene.trace(rays)
    spawned = scenes.secondary_rays(prepared, rays, hits)
    spawned = np.tile(spawned, (len(rays) + len(spawned) - 1) // len(spawned))[:len(rays)]
    low, high = (np.asarray(b, dtype=np.float32) for b in prepared.bounds)

    origin_key = morton(spawned["origin"], low, high)
    octant = ((spawned["direction"][:, 0] > 0).astype(np.uint64) << np.uint64(2)) | ((spawned["direction"][:, 1] > 0).astype(np.uint64) << np.uint64(1)) | (spawned["direction"][:, 2] > 0).astype(np.uint64)
    orders = {"as spawned": None, "sorted by origin cell (30-bit Morton)": np.argsort(origin_key, kind="stable"),
              "sorted by direction octant, then origin cell": np.argsort((octant << np.uint64(30)) | origin_key, kind="stable"),
              "sorted by origin cell (15-bit Morton), then octant": np.argsort(((origin_key >> np.uint64(15)) << np.uint64(3)) | octant, kind="stable"),
              "shuffled": np.random.default_rng(3).permutation(len(spawned))}

    n = len(spawned)
    d_out = torch.empty(n * 16, dtype=torch.uint8, device=device)
    d_flags = torch.empty(n, dtype=torch.uint8, device=device)
    results = {}
    reference_tokens = None

    for name, order in orders.items():
        batch = spawned if order is None else np.ascontiguousarray(spawned[order])
        d_rays = torch.from_numpy(batch.view(np.uint8).reshape(-1)).to(device)
        for _ in range(3):
            scene.trace_device(d_rays.data_ptr(), n, d_out.data_ptr(), stream)
            scene.occlude_device(d_rays.data_ptr(), n, d_flags.data_ptr(), stream)
        events = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        trace_ms = occlude_ms = 0.0
        for _ in range(args.steps):
            events[0].record()
            scene.trace_device(d_rays.data_ptr(), n, d_out.data_ptr(), stream)
            events[1].record()
            scene.occlude_device(d_rays.data_ptr(), n, d_flags.data_ptr(), stream)
            events[2].record()
            torch.cuda.synchronize()
            trace_ms += events[0].elapsed_time(events[1]) / args.steps
            occlude_ms += events[1].elapsed_time(events[2]) / args.steps
        tokens = d_out.cpu().numpy().view(structs.HIT)["token"]
        if order is not None:
            restored = np.empty_like(tokens)
            restored[order] = tokens
            tokens = restored
        if reference_tokens is None:
            reference_tokens = tokens
        results[name] = {"closest_mrays_per_s": n / (trace_ms * 1e-3) / 1e6, "any_hit_mrays_per_s": n / (occlude_ms * 1e-3) / 1e6, "same_hits": bool(np.array_equal(tokens, reference_tokens))}
        print(json.dumps({name: results[name]}), flush=True)
        del d_rays

    scene.close()


if __name__ == "__main__":
    main()
