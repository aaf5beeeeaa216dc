"""One device SweepBuilder call on C2's geometry (after a warm-up call on 64 triangles): the command the ncu launch list of the build is taken from."""
import sys, time
sys.path.insert(0, ".")
from echorenderer_b200 import _native, build_qbvh_device, scenes

description = scenes.terrain_scene()
build_qbvh_device(description.triangles[:64], description.spheres[:0])
started = time.perf_counter()
nodes, depth = build_qbvh_device(description.triangles, description.spheres)
print({"call_ms": (time.perf_counter() - started) * 1e3, "nodes": len(nodes), "depth": depth, **_native.last_build()})
