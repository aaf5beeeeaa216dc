#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the hot path (see DESIGN.md "Measurement").

Default workload = BASELINE.json configs[1] (C2): 1 M-triangle synthetic QBVH (+10 k spheres), 16 Mi fully incoherent rays,
one closest-hit pass + one occlusion pass per step. Metric: Mrays/s, whole job, inputs resident in HBM (`value`), and the
same through the host-buffer C ABI with pinned host memory (`e2e`). `--workload render` measures the path tracer
(samples/s) on the C3-style mixed-material scene instead; it is reported with the same JSON shape.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload trace|render] [--rays R]

N > 1 is launched by torch.distributed.run, one rank per GPU: the scene is replicated, every rank traces its own ray batch
(weak scaling, no data-path collective); the render workload shards tiles across ranks and all-reduces the frame (NCCL).
Only this file's cpu_baseline / --impl reference legs execute oracle/ (as the timed CPU stand-in for the reference's C# path).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from echorenderer_b200 import host, scenes, structs  # noqa: E402

MRAYS = 1e6


def parse():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=5)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", default="b200", choices=["b200", "reference"])
    parser.add_argument("--workload", default="trace", choices=["trace", "render"])
    parser.add_argument("--rays", type=int, default=1 << 24, help="rays per pass per GPU (C2: 16 Mi)")
    parser.add_argument("--quads", type=int, nargs=2, default=[1000, 500], help="terrain quads (C2: 1000 x 500 = 1 M triangles)")
    parser.add_argument("--cpu-sample", type=int, default=1 << 24,
                        help="rays per pass of the bounded CPU baseline sample (default: the whole 16 Mi-ray batch, ~1 s on 16 threads = ~13 core-seconds)")
    parser.add_argument("--width", type=int, default=1920)
    parser.add_argument("--height", type=int, default=1080)
    parser.add_argument("--spp", type=int, default=16, help="render workload: samples per pixel per step (one epoch)")
    parser.add_argument("--scene", default="mixed", choices=["cornell", "mixed", "lights", "large", "instanced", "textured"],
                        help="render workload scene: C1 / C3 / C4 / C5 / 2 304 placements of two packs (SURVEY.md 8f rank 2)")
    parser.add_argument("--instanced", action="store_true", help="trace workload: the instanced scene instead of the C2 terrain (same ray recipe)")
    parser.add_argument("--bounce-limit", type=int, default=8, help="render workload: PathTracedEvaluator.BounceLimit (C3: 8; reference default 128)")
    parser.add_argument("--tree", default="sah", choices=["sah", "device"],
                        help="trace workload: the SweepBuilder mirror's tree (default) or the device-built linear BVH (echo_b200_build_qbvh)")
    parser.add_argument("--shard", default="tiles", choices=["tiles", "samples"],
                        help="render workload, N > 1: tile sharding (tile i -> rank i mod N) or sample sharding (every rank renders spp / N samples of every tile)")
    parser.add_argument("--pattern", default="ordered", choices=["hilbert", "ordered"],
                        help="render workload: tile sequence, an OrderedPattern or the HilbertCurvePattern of EvaluationProfile.Pattern (A/B in profiles/README.md)")
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--no-secondary", action="store_true", help="skip the secondary-ray batch reported beside the headline (SURVEY.md 8d)")
    return parser.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as file:
            return float(json.load(file)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed `ncu --set full` capture of this
    command (profiles/ncu_traffic.json); null when no capture has been recorded."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        with open(path) as file:
            return json.load(file).get(key)
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe). nvidia-smi needs a moment to
    start, so the sampler is started before the warm-up and `mark()`s delimit the timed region; samples are stamped on arrival."""
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.lines, self.process, self.index, self.window = [], None, index, [None, None]

    def __enter__(self):
        try:
            self.process = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20"],
                                            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            deadline = time.time() + 3.0
            while not self.lines and time.time() < deadline:
                time.sleep(0.02)
        except OSError:
            self.process = None
        return self

    def _read(self):
        for line in self.process.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self, which):
        self.window[which] = time.time()

    def __exit__(self, *_):
        if self.process:
            self.process.terminate()
            try:
                self.process.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.process.kill()

    def summary(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        begin, end = self.window
        rows = []
        for stamp, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                rows.append((stamp, float(parts[0]), float(parts[1]), [n for n, f in zip(names, parts[4:8]) if f.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [r for r in rows if begin is not None and end is not None and begin - 0.03 <= r[0] <= end + 0.03]
        scope = "timed region"
        if not inside:  # region shorter than the sampling period: fall back to every sample taken under load (warm-up + timed)
            inside, scope = [r for r in rows if begin is None or r[0] >= begin - 1.0], "warm-up + timed region"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted({reason for r in inside for reason in r[3]})
        return {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": max(r[2] for r in inside), "reasons": reasons,
                "samples": len(inside), "scope": scope}


def instanced_bench_scene():
    return scenes.instanced_scene(grid=48, rings=128, segments=130)


def build_trace_inputs(args, rank):
    description = instanced_bench_scene() if getattr(args, "instanced", False) else scenes.terrain_scene(args.quads[0], args.quads[1], 10000)
    tree = None
    if getattr(args, "tree", "sah") == "device" and not getattr(args, "instanced", False):
        from echorenderer_b200 import build_qbvh_device
        build_qbvh_device(description.triangles[:64], description.spheres[:0], device=int(os.environ.get("LOCAL_RANK", "0")))  # context + module load
        started = time.perf_counter()
        tree = build_qbvh_device(description.triangles, description.spheres, device=int(os.environ.get("LOCAL_RANK", "0")))
        args.tree_build_seconds = time.perf_counter() - started
    started = time.perf_counter()
    prepared = host.prepare(description, tree=tree)
    args.prepare_seconds = time.perf_counter() - started
    # every rank traces its own batch (distinct seeds per rank)
    rays = scenes.random_rays(prepared.bounds, args.rays, seed=11 + 1000 * rank)
    shadow = rays.copy()
    low, high = prepared.bounds
    diagonal = float(np.linalg.norm(np.asarray(high, np.float64) - np.asarray(low, np.float64)))
    shadow["distance"] = scenes.uniform(17 + 1000 * rank, np.arange(args.rays, dtype=np.uint64)) * np.float32(diagonal)
    return prepared, rays, shadow


def secondary_batch(scene, prepared, rays, d_hits, device, stream, args, max_over_ranks):
    """The second batch SURVEY.md §8(d) asks to report beside the headline: rays leaving the surfaces the first batch hit
    (cosine-hemisphere directions, ignore = hit token, as TraceQuery.SpawnTrace does), tiled up to the size of the first batch.
    Reported separately; it is not part of `value`."""
    import torch
    hits = d_hits.cpu().numpy().view(structs.HIT)
    spawned = scenes.secondary_rays(prepared, rays, hits)
    if len(spawned) == 0:
        return None
    unique = len(spawned)
    spawned = np.tile(spawned, (len(rays) + unique - 1) // unique)[:len(rays)]
    n = len(spawned)
    shadow = spawned.copy()
    low, high = prepared.bounds
    diagonal = float(np.linalg.norm(np.asarray(high, np.float64) - np.asarray(low, np.float64)))
    shadow["distance"] = scenes.uniform(29, np.arange(n, dtype=np.uint64)) * np.float32(diagonal)

    d_rays = torch.from_numpy(spawned.view(np.uint8).reshape(-1)).to(device)
    d_shadow = torch.from_numpy(shadow.view(np.uint8).reshape(-1)).to(device)
    d_out = torch.empty(n * 16, dtype=torch.uint8, device=device)
    d_flags = torch.empty(n, dtype=torch.uint8, device=device)
    d_counts = torch.zeros(6, dtype=torch.int64, device=device)
    scene.trace_device(d_rays.data_ptr(), n, d_out.data_ptr(), stream, d_counts.data_ptr())
    scene.occlude_device(d_shadow.data_ptr(), n, d_flags.data_ptr(), stream, d_counts.data_ptr() + 24)
    torch.cuda.synchronize()
    counts = d_counts.cpu().numpy().astype(np.float64) / n
    hit_rate = float(np.mean(d_out.cpu().numpy().view(structs.HIT)["token"] != structs.TOKEN_EMPTY))

    for _ in range(args.warmup):
        scene.trace_device(d_rays.data_ptr(), n, d_out.data_ptr(), stream)
        scene.occlude_device(d_shadow.data_ptr(), n, d_flags.data_ptr(), stream)
    events = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    trace_ms = occlude_ms = 0.0
    torch.cuda.synchronize()
    for _ in range(args.steps):
        events[0].record()
        scene.trace_device(d_rays.data_ptr(), n, d_out.data_ptr(), stream)
        events[1].record()
        scene.occlude_device(d_shadow.data_ptr(), n, d_flags.data_ptr(), stream)
        events[2].record()
        torch.cuda.synchronize()
        trace_ms += events[0].elapsed_time(events[1]) / args.steps
        occlude_ms += events[1].elapsed_time(events[2]) / args.steps
    trace_ms, occlude_ms = max_over_ranks(trace_ms), max_over_ranks(occlude_ms)
    bytes_trace = 32 + 16 + 128 * counts[0] + 36 * counts[1] + 16 * counts[2]
    bytes_occlude = 32 + 1 + 128 * counts[3] + 36 * counts[4] + 16 * counts[5]
    return {"rays_per_pass_per_gpu": n, "distinct_rays": unique, "hit_rate": hit_rate,
            "closest_hit": {"mrays_per_s": n / (trace_ms * 1e-3) / MRAYS, "ms_per_launch": trace_ms, "algorithmic_bytes_per_query": bytes_trace,
                            "achieved_gbs": bytes_trace * n / (trace_ms * 1e-3) / 1e9,
                            "visits_per_query": {"nodes": counts[0], "triangles": counts[1], "spheres": counts[2]}},
            "occlusion": {"mrays_per_s": n / (occlude_ms * 1e-3) / MRAYS, "ms_per_launch": occlude_ms, "algorithmic_bytes_per_query": bytes_occlude,
                          "achieved_gbs": bytes_occlude * n / (occlude_ms * 1e-3) / 1e9,
                          "visits_per_query": {"nodes": counts[3], "triangles": counts[4], "spheres": counts[5]}}}


def cpu_baseline_trace(prepared, rays, shadow, sample, threads=0):
    """The oracle (C++ restatement of the reference's CPU path) on a bounded sample, all host threads."""
    from tests import oracle_lib
    oracle = oracle_lib.OracleScene(prepared)
    sample = min(sample, len(rays))
    cores = threads or os.cpu_count() or 1
    oracle.trace(rays[:4096], threads=cores)  # warm-up (page in the scene)
    start = time.perf_counter()
    oracle.trace(rays[:sample], threads=cores)
    oracle.occlude(shadow[:sample], threads=cores)
    seconds = time.perf_counter() - start
    whole = "all" if sample == len(rays) else "first"
    return 2 * sample / seconds / MRAYS, cores, seconds, f"{whole} {sample} closest-hit + {sample} occlusion queries of the rank-0 batch, {cores} threads"


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path. Echo is C#/.NET 6 and cannot be built or run
    here (no dotnet), so the C++ oracle port stands in, on all host threads; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return

    if args.workload == "trace":
        prepared, rays, shadow = build_trace_inputs(argparse.Namespace(**{**vars(args), "rays": args.cpu_sample}), 0)
        from tests import oracle_lib
        oracle = oracle_lib.OracleScene(prepared)
        cores = os.cpu_count() or 1
        times = []
        for step in range(args.warmup + args.steps):
            start = time.perf_counter()
            oracle.trace(rays, threads=cores)
            oracle.occlude(shadow, threads=cores)
            if step >= args.warmup:
                times.append(time.perf_counter() - start)
        seconds = sum(times)
        value = 2 * len(rays) * args.steps / seconds / MRAYS
        sample = f"{len(rays)} closest-hit + {len(rays)} occlusion queries per step ({'the whole' if len(rays) == args.rays else 'a bounded sample of the'} {args.rays}-ray batch), {cores} threads"
        line = base_line(args, "Mrays/s", value, seconds / args.steps * 1e3, trace_config(args, len(rays)), "f32")
    else:
        prepared = host.prepare(RENDER_SCENES[args.scene][1]())
        from tests import oracle_lib
        oracle = oracle_lib.OracleScene(prepared)
        cores = os.cpu_count() or 1
        width, height = 256, 144  # bounded sample: a 256 x 144 crop of the frame's tile grid at the same spp
        params = structs.render_params(args.width, args.height, 16, extend=args.spp, min_epoch=1, max_epoch=1, bounce_limit=args.bounce_limit)
        tiles = scenes.tile_grid(args.width, args.height, 16)
        tiles_x = (args.width + 15) // 16
        crop = tiles.reshape(-1, tiles_x, 2)[(args.height // 32) - 4:(args.height // 32) + 5, (tiles_x // 2) - 8:(tiles_x // 2) + 8].reshape(-1, 2)
        times, samples = [], 0
        for step in range(args.warmup + args.steps):
            start = time.perf_counter()
            _, stats = oracle.render_tiles(params, crop, threads=cores)
            if step >= args.warmup:
                times.append(time.perf_counter() - start)
                samples += int(stats["sampleEvaluated"][0])
        seconds = sum(times)
        value = samples / seconds
        sample = f"{len(crop)} central tiles ({width}x{height} px region) of the {args.width}x{args.height} frame at {args.spp} spp per step"
        line = base_line(args, "samples/s", value, seconds / args.steps * 1e3, render_config(args), "f32")

    line["impl"] = "reference"
    line["cpu_baseline"] = {"value": line["value"], "unit": line["unit"], "cores": cores, "kind": "port", "sample": sample}
    line["e2e"] = {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    line["gpu_launches"] = 0
    emit(line)


def trace_config(args, rays):
    if getattr(args, "instanced", False):
        return {"workload": "incoherent ray-batch intersection through instanced packs (48 x 48 placements of a 33 k-triangle pack and of a pack of three of those): closest-hit + occlusion",
                "rays_per_pass_per_gpu": rays, "passes_per_step": 2, "parallelism": f"replicated scene x{args.gpus}",
                "l2": "inputs larger than L2 (ray batch 512 MiB + hit buffer 256 MiB per pass vs 126 MB L2)"}
    return {"workload": "C2 incoherent ray-batch intersection: closest-hit + occlusion", "triangles": args.quads[0] * args.quads[1] * 2,
            "spheres": 10000, "rays_per_pass_per_gpu": rays, "passes_per_step": 2, "parallelism": f"replicated scene x{args.gpus}",
            "l2": "inputs larger than L2 (ray batch 512 MiB + hit buffer 256 MiB per pass vs 126 MB L2)"}


RENDER_SCENES = {
    "cornell": ("C1 Cornell box (CornellBox.cs / cornell.echo)", scenes.cornell_box),
    "mixed": ("C3 mixed-material scene (Dielectric, Conductor GGX, Oren-Nayar)", scenes.mixed_material_scene),
    "lights": ("C4 many-lights scene (10 k emissive triangles, light-tree NEE)", scenes.many_lights_scene),
    "large": ("C5 ~10 M-triangle terrain with the C3 material mix", scenes.large_scene),
    "textured": ("image textures in every material slot, normal map, alpha cut-out (SURVEY.md 8f rank 3)", lambda: scenes.textured_scene(rings=128, segments=130)),
    "instanced": ("2 304 placements of two packs (one nests three placements of the other), lights inside the packs", instanced_bench_scene),
}


def render_config(args):
    return {"workload": f"{RENDER_SCENES[args.scene][0]}, path tracer bounce limit {args.bounce_limit}", "width": args.width,
            "height": args.height, "spp_per_step": args.spp, "parallelism": f"{'sample' if args.shard == 'samples' else 'tile'}-sharded x{args.gpus} + NCCL all-reduce of the frame",
            "tile_pattern": args.pattern, "l2": "wavefront state (up to ~4 GB) larger than L2"}


def base_line(args, unit, value, ms_per_step, config, dtype):
    # trace: every rank traces its own batch (weak); render: one fixed frame sharded by tiles (strong)
    return {"metric": "Mrays/s" if unit == "Mrays/s" else "samples/sec", "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if unit == "Mrays/s" else "strong", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "config": config}


_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints its version banner from C when the box
    sets NCCL_DEBUG=VERSION), so file descriptor 1 points at stderr while the bench runs and emit() writes to the real one."""
    global _STDOUT
    if _STDOUT is None:
        sys.stdout.flush()
        _STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _STDOUT if _STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse()
    if args.warmup < 3:
        args.warmup = 3
    quiet_stdout()

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from echorenderer_b200 import PreparedScene

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)

    if distributed:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(milliseconds):
        if not distributed:
            return milliseconds
        tensor = torch.tensor([milliseconds], dtype=torch.float64, device=device)
        dist.all_reduce(tensor, op=dist.ReduceOp.MAX)
        return float(tensor.item())

    peak, peak_source = peaks()

    if args.workload == "trace":
        prepared, rays, shadow = build_trace_inputs(args, rank)
        scene = PreparedScene(prepared, device=local_rank)
        n = len(rays)

        def to_u8(array):
            return torch.from_numpy(array.view(np.uint8).reshape(-1))

        host_rays, host_shadow = to_u8(rays).pin_memory(), to_u8(shadow).pin_memory()
        host_hits = torch.empty(n * 16, dtype=torch.uint8).pin_memory()
        host_occluded = torch.empty(n, dtype=torch.uint8).pin_memory()

        d_rays, d_shadow = host_rays.to(device), host_shadow.to(device)
        d_hits = torch.empty(n * 16, dtype=torch.uint8, device=device)
        d_occluded = torch.empty(n, dtype=torch.uint8, device=device)
        d_counts = torch.zeros(6, dtype=torch.int64, device=device)
        stream = torch.cuda.current_stream().cuda_stream

        # visit counters -> algorithmic bytes per query; one untimed counted pass
        scene.trace_device(d_rays.data_ptr(), n, d_hits.data_ptr(), stream, d_counts.data_ptr())
        scene.occlude_device(d_shadow.data_ptr(), n, d_occluded.data_ptr(), stream, d_counts.data_ptr() + 24)
        torch.cuda.synchronize()
        counts = d_counts.cpu().numpy().astype(np.float64) / n
        bytes_trace = 32 + 16 + 128 * counts[0] + 36 * counts[1] + 16 * counts[2]
        bytes_occlude = 32 + 1 + 128 * counts[3] + 36 * counts[4] + 16 * counts[5]

        def step():
            scene.trace_device(d_rays.data_ptr(), n, d_hits.data_ptr(), stream)
            scene.occlude_device(d_shadow.data_ptr(), n, d_occluded.data_ptr(), stream)

        events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]

        with ClockSampler(local_rank) as clocks:
            for _ in range(args.warmup):
                step()
            barrier()

            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            clocks.mark(0)
            start.record()
            for first, middle, last in events:
                first.record()
                scene.trace_device(d_rays.data_ptr(), n, d_hits.data_ptr(), stream)
                middle.record()
                scene.occlude_device(d_shadow.data_ptr(), n, d_occluded.data_ptr(), stream)
                last.record()
            stop.record()
            barrier()
            clocks.mark(1)
            total_ms = max_over_ranks(start.elapsed_time(stop))

        trace_ms = float(np.mean([a.elapsed_time(b) for a, b, _ in events]))
        occlude_ms = float(np.mean([b.elapsed_time(c) for _, b, c in events]))
        value = world * 2 * n * args.steps / (total_ms * 1e-3) / MRAYS

        # end to end through the host-buffer C ABI: pinned host rays in, hits/occlusion flags out, every step
        for _ in range(2):
            scene.trace_pointers(host_rays.data_ptr(), n, host_hits.data_ptr())
        barrier()
        wall = time.perf_counter()
        for _ in range(args.steps):
            scene.trace_pointers(host_rays.data_ptr(), n, host_hits.data_ptr())
            scene.occlude_pointers(host_shadow.data_ptr(), n, host_occluded.data_ptr())
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - wall) * 1e3)
        e2e_value = world * 2 * n * args.steps / (e2e_ms * 1e-3) / MRAYS

        line = base_line(args, "Mrays/s", value, total_ms / args.steps, trace_config(args, n), "f32")
        achieved = bytes_trace * n / (trace_ms * 1e-3) / 1e9
        line["roofline"] = {"bound": "hbm", "kernel": "instanced closest-hit kernel" if args.instanced else "persistent_batch_kernel<48, false> (closest hit)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": ncu_traffic("closest_hit_dram_bytes_per_launch"), "peak_source": peak_source, "algorithmic_bytes_per_query": bytes_trace, "ms_per_launch": trace_ms,
                            "mrays_per_s": n / (trace_ms * 1e-3) / MRAYS,
                            "visits_per_query": {"nodes": counts[0], "triangles": counts[1], "spheres": counts[2]},
                            "occlusion": {"kernel": "persistent_batch_kernel<48, true>", "traffic": ncu_traffic("occlusion_dram_bytes_per_launch"), "achieved": bytes_occlude * n / (occlude_ms * 1e-3) / 1e9, "algorithmic_bytes_per_query": bytes_occlude,
                                          "ms_per_launch": occlude_ms, "mrays_per_s": n / (occlude_ms * 1e-3) / MRAYS,
                                          "visits_per_query": {"nodes": counts[3], "triangles": counts[4], "spheres": counts[5]}}}
        line["e2e"] = {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 2 * n * 32, "d2h_bytes_per_step": n * 16 + n, "ms_per_step": e2e_ms / args.steps}
        line["gpu_launches"] = 2 * args.steps
        line["clocks"] = clocks.summary()
        line["config"]["tree"] = {"builder": args.tree, "nodes": int(len(prepared.nodes)), "quad_depth": int(prepared.max_depth),
                                  "device_build_seconds": getattr(args, "tree_build_seconds", None), "host_prepare_seconds": getattr(args, "prepare_seconds", None)}

        if not args.no_secondary and not args.instanced:
            line["secondary"] = secondary_batch(scene, prepared, rays, d_hits, device, stream, args, max_over_ranks)

        if rank == 0 and not args.no_cpu_baseline:
            cpu_value, cores, seconds, sample = cpu_baseline_trace(prepared, rays, shadow, args.cpu_sample)
            line["cpu_baseline"] = {"value": cpu_value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample, "seconds": seconds}
    else:
        prepared = host.prepare(RENDER_SCENES[args.scene][1]())
        scene = PreparedScene(prepared, device=local_rank)
        width, height, tile = args.width, args.height, 16
        from echorenderer_b200 import hilbert_curve_pattern, shard_tiles
        tile_count = ((width + tile - 1) // tile, (height + tile - 1) // tile)
        all_tiles = hilbert_curve_pattern(tile_count) if args.pattern == "hilbert" else scenes.tile_grid(width, height, tile)
        by_samples = args.shard == "samples" and world > 1
        tiles = all_tiles if by_samples else shard_tiles(all_tiles, rank, world)
        extend = max(1, args.spp // world) if by_samples else args.spp
        frame = torch.zeros(height * width * 4, dtype=torch.float32, device=device)
        host_frame = torch.empty(height * width * 4, dtype=torch.float32).pin_memory()
        stream = torch.cuda.current_stream().cuda_stream
        launches = 0
        samples = 0

        def step(index):
            nonlocal launches, samples
            # sample sharding: rank r renders epoch index * world + r, spp / world samples per pixel; the summed weight is `world`
            epoch = index * world + rank if by_samples else index
            params = structs.render_params(width, height, tile, extend=extend, min_epoch=1, max_epoch=1, bounce_limit=args.bounce_limit, seed=1, epoch_offset=epoch)
            frame.zero_()
            stats = scene.render_frame_device(params, tiles, frame.data_ptr(), stream)
            if distributed:
                dist.all_reduce(frame)  # the accumulation-buffer reduce over NVLink (disjoint tiles: sum with zeros)
            scene.frame_resolve_device(frame.data_ptr(), width, height, stream)
            launches += int(stats["kernelLaunches"][0]) + 1
            samples += int(stats["sampleEvaluated"][0])
            return stats

        with ClockSampler(local_rank) as clocks:
            for index in range(args.warmup):
                step(index)

            launches = samples = 0
            barrier()
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            clocks.mark(0)
            start.record()
            for index in range(args.steps):
                stats = step(args.warmup + index)
            stop.record()
            barrier()
            clocks.mark(1)
            total_ms = max_over_ranks(start.elapsed_time(stop))

        total_samples = samples
        if distributed:
            tensor = torch.tensor([samples], dtype=torch.float64, device=device)
            dist.all_reduce(tensor)
            total_samples = int(tensor.item())
        value = total_samples / (total_ms * 1e-3)

        # end to end: render + all-reduce + device->host read of the resolved frame, every step
        barrier()
        wall = time.perf_counter()
        for index in range(args.steps):
            step(args.warmup + index)
            host_frame.copy_(frame, non_blocking=True)
            torch.cuda.synchronize()
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - wall) * 1e3)

        line = base_line(args, "samples/s", value, total_ms / args.steps, render_config(args), "f32")
        queries = int(stats["traceQueries"][0]) + int(stats["occludeQueries"][0])
        line["roofline"] = {"bound": "hbm", "kernel": "extend_kernel + shadow_kernel (wavefront)", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                            "peak_source": peak_source, "queries_per_step_last": queries}
        line["e2e"] = {"value": total_samples / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": int(tiles.nbytes), "d2h_bytes_per_step": height * width * 16,
                       "ms_per_step": e2e_ms / args.steps}
        line["gpu_launches"] = launches
        line["clocks"] = clocks.summary()
        line["stats_last_step"] = {label: int(stats[name][0]) for label, name in zip(structs.STATS_LABELS, structs.STATS_FIELDS)}

    if rank == 0:
        emit(line)

    scene.close()
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
