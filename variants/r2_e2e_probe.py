"""Round 2: per-call wall time of echo_b200_render_tiles (host buffers) against echo_b200_render_frame_device on the same tiles — the
plugin-call leg of the render records showed an erratic extra cost (C3: +5 ... +240 ms per step between runs)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from echorenderer_b200 import PreparedScene, _native, host, scenes, structs, hilbert_curve_pattern  # noqa: E402

import torch

prepared = host.prepare(scenes.mixed_material_scene())
scene = PreparedScene(prepared, device=0)
width, height, tile, spp = 1920, 1080, 16, 64
tiles = hilbert_curve_pattern(((width + 15) // 16, (height + 15) // 16))
frame = torch.zeros(height * width * 4, dtype=torch.float32, device="cuda:0")
stream = torch.cuda.current_stream().cuda_stream
buffer = _native.HostBuffer(len(tiles) * tile * tile * 4, np.float32)


def params(index):
    return structs.render_params(width, height, tile, extend=spp, min_epoch=1, max_epoch=1, bounce_limit=8, seed=1, epoch_offset=index)


for label, call in (("frame_device", lambda i: scene.render_frame_device(params(i), tiles, frame.data_ptr(), stream)),
                    ("render_tiles (host)", lambda i: scene.render_tiles_pointers(params(i), tiles, buffer.address)),
                    ("frame_device again", lambda i: scene.render_frame_device(params(i), tiles, frame.data_ptr(), stream)),
                    ("render_tiles (host) again", lambda i: scene.render_tiles_pointers(params(i), tiles, buffer.address))):
    times = []
    for index in range(6):
        torch.cuda.synchronize()
        begin = time.perf_counter()
        call(index)
        torch.cuda.synchronize()
        times.append(round((time.perf_counter() - begin) * 1e3, 1))
    print(label, times, flush=True)
scene.close()
