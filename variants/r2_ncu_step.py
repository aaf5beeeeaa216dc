"""One step of a render workload between cudaProfilerStart / cudaProfilerStop, for `ncu --profile-from-start off`: the launch list, the
DRAM traffic of a whole wavefront step and `--set full` captures of its kernels (variants/r2_ncu.sh). ONE pipeline, so that the launches
arrive in program order: raygen, then per bounce extend, classify, rotate | shade x6, shadow.

  python variants/r2_ncu_step.py --scene mixed --spp 16
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("ECHO_B200_RENDER_WORKERS", "1")

from echorenderer_b200 import PreparedScene, host, scenes, structs, hilbert_curve_pattern  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--scene", default="mixed")
    parser.add_argument("--width", type=int, default=1920)
    parser.add_argument("--height", type=int, default=1080)
    parser.add_argument("--spp", type=int, default=16)
    parser.add_argument("--bounce-limit", type=int, default=8)
    args = parser.parse_args()

    import torch
    builders = {"large": scenes.large_scene, "mixed": scenes.mixed_material_scene, "lights": scenes.many_lights_scene, "cornell": scenes.cornell_box}
    prepared = host.prepare(builders[args.scene]())
    scene = PreparedScene(prepared, device=0)
    tile = 16
    tiles = hilbert_curve_pattern(((args.width + tile - 1) // tile, (args.height + tile - 1) // tile))
    frame = torch.zeros(args.height * args.width * 4, dtype=torch.float32, device="cuda:0")
    stream = torch.cuda.current_stream().cuda_stream

    def step(index):
        params = structs.render_params(args.width, args.height, tile, extend=args.spp, min_epoch=1, max_epoch=1, bounce_limit=args.bounce_limit, seed=1, epoch_offset=index)
        return scene.render_frame_device(params, tiles, frame.data_ptr(), stream)

    step(0)  # warm-up: allocations, module load
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    stats = step(1)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(json.dumps({"scene": args.scene, "samples": int(stats["sampleEvaluated"][0]), "launches": int(stats["kernelLaunches"][0])}))
    scene.close()


if __name__ == "__main__":
    main()
