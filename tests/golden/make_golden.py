"""Generates tests/golden/oracle_regression.npz from the ORACLE (the reference is C#/.NET and cannot run here, so these are
regression fixtures that freeze the oracle's bits, not reference-derived vectors). Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from echorenderer_b200 import host, scenes, structs  # noqa: E402
from tests import oracle_lib as ol  # noqa: E402

out = {}
cornell = host.prepare(scenes.cornell_box())
terrain = host.prepare(scenes.terrain_scene(96, 48, 400))

for name, prepared in (("cornell", cornell), ("terrain", terrain)):
    oracle = ol.OracleScene(prepared)
    hits = oracle.trace(scenes.random_rays(prepared.bounds, 4096, seed=21))
    out[f"{name}_token"] = hits["token"]
    out[f"{name}_distance_bits"] = hits["distance"].view(np.uint32)

oracle = ol.OracleScene(cornell)
params = structs.render_params(32, 32, 16, extend=4, seed=9)
ys, xs = np.meshgrid(np.arange(0, 32, 2), np.arange(0, 32, 2), indexing="ij")
pixels = np.repeat(np.stack([xs.reshape(-1), ys.reshape(-1)], axis=-1), 2, axis=0).astype(np.int32)
index = np.tile(np.arange(2, dtype=np.uint32), len(pixels) // 2)
out["cornell_radiance_bits"] = oracle.evaluate_samples(params, pixels, index).view(np.uint32)

# the rows added after the hot path: instanced packs, image textures + environment light, the auxiliary passes
instanced = host.prepare(scenes.instanced_scene(grid=3, rings=8, segments=10))
oracle = ol.OracleScene(instanced)
hits, layers = oracle.trace_hierarchy(scenes.random_rays(instanced.bounds, 4096, seed=21))
out["instanced_token"], out["instanced_distance_bits"] = hits["token"], hits["distance"].view(np.uint32)
out["instanced_layers"] = np.concatenate([layers["instanceCount"][:, None], layers["instances"]], axis=1)
params = structs.render_params(32, 32, 16, extend=2, seed=9, bounce_limit=12)
half_pixels, half_index = np.ascontiguousarray(pixels[::2]), np.ascontiguousarray(index[::2])
out["instanced_radiance_bits"] = oracle.evaluate_samples(params, half_pixels, half_index).view(np.uint32)

textured = host.prepare(scenes.textured_scene(rings=8, segments=10))
oracle = ol.OracleScene(textured)
out["textured_radiance_bits"] = oracle.evaluate_samples(params, half_pixels, half_index).view(np.uint32)
for name, code in (("albedo", structs.EVALUATOR_ALBEDO | structs.EVALUATOR_DIVERGE_ONCE), ("normal_depth", structs.EVALUATOR_NORMAL_DEPTH)):
    aux = structs.render_params(32, 32, 16, extend=2, seed=9, evaluator=code)
    value = np.zeros((len(half_index), 4), dtype=np.float32)
    oracle.lib.oracle_evaluate_samples4(oracle.handle, ol.ptr(aux), ol.ptr(half_pixels), ol.ptr(half_index), len(value), ol.ptr(value), 4, 0)
    out[f"textured_{name}_bits"] = value.view(np.uint32)

np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_regression.npz"), **out)
print({k: v.shape for k, v in out.items()})
