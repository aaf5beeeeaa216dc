"""r2ad: the device light-tree build timed on its own: C4's emitters (10 000 of 309 502 triangles) and larger random emitter sets, five calls
each (echo_b200_debug_last_light_build phases), beside the host mirror's recursive build. Prints one JSON line per case."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from echorenderer_b200 import _native, build_light_tree_device, host, scenes  # noqa: E402
from tests.test_light_build import random_emitters  # noqa: E402

cases = [("c4", scenes.many_lights_scene()), ("random_24k", random_emitters(8, 40_000, 2000, 100, 30.0)), ("random_140k", random_emitters(9, 200_000, 0, 0, 30.0)),
         ("same_facing_10k", None)]
for name, description in cases:
    if description is None:  # emitters that all face down (a ceiling of lights): the cone never grows to the whole sphere, every step runs the full union
        from echorenderer_b200 import structs
        description = scenes.many_lights_scene()
        lit = np.flatnonzero(description.materials["type"][description.triangles["material"]] == structs.MATERIAL_EMISSIVE)
        side = np.linalg.norm(description.triangles["edge1"][lit], axis=1, keepdims=True)
        description.triangles["edge1"][lit] = side * np.array([1.0, 0.0, 0.0], dtype=np.float32)
        description.triangles["edge2"][lit] = side * np.array([0.0, 0.0, 1.0], dtype=np.float32)  # edge1 x edge2 = -y: every light faces down
    for _ in range(3):  # context, module load, clocks
        build_light_tree_device(description)
    calls = []
    for _ in range(5):
        started = time.perf_counter()
        nodes, tokens, paths, power = build_light_tree_device(description)
        calls.append({"call_ms": (time.perf_counter() - started) * 1e3, **_native.last_light_build()})
    started = time.perf_counter()
    expected = host.build_light_tree(description)
    host_ms = (time.perf_counter() - started) * 1e3
    identical = nodes.tobytes() == expected[0].tobytes() and tokens.tobytes() == expected[1].tobytes() and paths.tobytes() == expected[2].tobytes()
    print(json.dumps({"case": name, "emitters": len(tokens), "device_build_ms": sorted(c["device_build_ms"] for c in calls), "call_ms": sorted(c["call_ms"] for c in calls),
                      "levels": calls[0]["levels"], "host_mirror_ms": host_ms, "identical": bool(identical)}), flush=True)
