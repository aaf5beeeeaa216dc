"""Host-side scene preparation: the Python face of libecho_host.so (include/echo_host.h).

Mirrors what Echo's C# host does once per render (ScenePreparer -> PreparedPack, Aggregation/Preparation/PreparedPack.cs:17-25):
build the SweepBuilder hierarchy, collapse it into the QuadBoundingVolumeHierarchy node array, build the LightTree and
compute the infinite-light threshold (PreparedScene.cs:34-39). The results are the arrays echo_b200.h takes.
"""
import ctypes
import os
from dataclasses import dataclass, field

import numpy as np

from . import structs

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _library():
    global _lib
    if _lib is not None:
        return _lib

    path = os.path.join(_HERE, "libecho_host.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (or make -C echorenderer_b200/csrc)")

    lib = ctypes.CDLL(path)
    p = ctypes.c_void_p
    u32 = ctypes.c_uint32
    lib.echo_host_build_qbvh.argtypes = [p, u32, p, u32, ctypes.c_int32, ctypes.POINTER(p), ctypes.POINTER(u32), ctypes.POINTER(u32)]
    lib.echo_host_build_qbvh.restype = ctypes.c_int32
    lib.echo_host_build_light_tree.argtypes = [p, u32, p, u32, p, u32, p, u32, ctypes.POINTER(p), ctypes.POINTER(u32),
                                               ctypes.POINTER(p), ctypes.POINTER(p), ctypes.POINTER(u32), ctypes.POINTER(ctypes.c_float)]
    lib.echo_host_build_light_tree.restype = ctypes.c_int32
    lib.echo_host_infinite_threshold.argtypes = [ctypes.c_float, ctypes.c_float]
    lib.echo_host_infinite_threshold.restype = ctypes.c_float
    lib.echo_host_ambient_power.argtypes = [p, p]
    lib.echo_host_ambient_power.restype = ctypes.c_float
    lib.echo_host_emissive_power.argtypes = [p]
    lib.echo_host_emissive_power.restype = ctypes.c_float
    lib.echo_host_fresnel_diffuse_reflectance.argtypes = [ctypes.c_float]
    lib.echo_host_fresnel_diffuse_reflectance.restype = ctypes.c_float
    lib.echo_host_fresnel_diffuse_reflectance_fast.argtypes = [ctypes.c_float]
    lib.echo_host_fresnel_diffuse_reflectance_fast.restype = ctypes.c_float
    lib.echo_host_free.argtypes = [p]
    lib.echo_host_free.restype = None
    _lib = lib
    return lib


def _pointer(array):
    return ctypes.c_void_p(array.ctypes.data) if array is not None and array.size else ctypes.c_void_p(0)


def _take(pointer, count, dtype):
    """Copies a malloc'd result array into numpy and frees the original."""
    lib = _library()
    if count == 0 or not pointer.value:
        if pointer.value:
            lib.echo_host_free(pointer)
        return np.zeros(0, dtype=dtype)
    buffer = (ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(pointer.value)
    result = np.frombuffer(buffer, dtype=dtype, count=count).copy()
    lib.echo_host_free(pointer)
    return result


@dataclass
class SceneDescription:
    """What Echo's Scene + ScenePreparer hand to PreparedScene, with constant (Pure) textures flattened."""
    triangles: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=structs.TRIANGLE))
    spheres: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=structs.SPHERE))
    materials: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=structs.MATERIAL))
    point_lights: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=structs.POINT_LIGHT))
    infinite_lights: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=structs.INFINITE_LIGHT))
    camera: np.ndarray = field(default_factory=lambda: np.zeros(1, dtype=structs.CAMERA))
    name: str = "scene"


@dataclass
class PreparedArrays:
    """The flattened, immutable scene: the exact payload of the echo_b200_scene_set_* calls."""
    description: SceneDescription
    nodes: np.ndarray
    max_depth: int
    light_nodes: np.ndarray
    emitter_tokens: np.ndarray
    emitter_bitpaths: np.ndarray
    infinite_threshold: float
    infinite_pdf: float
    scene_power: float

    @property
    def triangles(self):
        return self.description.triangles

    @property
    def spheres(self):
        return self.description.spheres

    @property
    def materials(self):
        return self.description.materials

    @property
    def bounds(self):
        """BoxBound of the whole accelerator (Accelerator.BoxBound): min/max over the root node's children."""
        root = self.nodes[0]
        valid = root["token4"] != structs.TOKEN_EMPTY
        low = np.array([root[k][valid].min() for k in ("minX", "minY", "minZ")], dtype=np.float32)
        high = np.array([root[k][valid].max() for k in ("maxX", "maxY", "maxZ")], dtype=np.float32)
        return low, high


def fresnel_diffuse_reflectance(eta, fast=False):
    """CoatedLambertianReflection.FresnelDiffuseReflectance[Fast] (Evaluation/Scattering/Lambertian.cs:168-230)."""
    lib = _library()
    return float((lib.echo_host_fresnel_diffuse_reflectance_fast if fast else lib.echo_host_fresnel_diffuse_reflectance)(float(eta)))


def build_qbvh(triangles, spheres, threads=0):
    """SweepBuilder + QuadBoundingVolumeHierarchy constructor (SweepBuilder.cs:24-36, QuadBoundingVolumeHierarchy.cs:24-36)."""
    lib = _library()
    triangles = np.ascontiguousarray(triangles, dtype=structs.TRIANGLE)
    spheres = np.ascontiguousarray(spheres, dtype=structs.SPHERE)
    nodes, count, depth = ctypes.c_void_p(), ctypes.c_uint32(), ctypes.c_uint32()
    status = lib.echo_host_build_qbvh(_pointer(triangles), len(triangles), _pointer(spheres), len(spheres), threads,
                                      ctypes.byref(nodes), ctypes.byref(count), ctypes.byref(depth))
    if status != 0:
        raise ValueError(f"echo_host_build_qbvh failed with status {status} (needs 2..2^28-1 primitives)")
    return _take(nodes, count.value, structs.QBVH_NODE), int(depth.value)


def build_light_tree(description):
    """LightCollection.CreateBounds + LightTree constructor (LightCollection.cs:91-137, LightTree.cs:21-38)."""
    lib = _library()
    d = description
    nodes, tokens, paths = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    node_count, emitter_count, power = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_float()
    status = lib.echo_host_build_light_tree(_pointer(d.triangles), len(d.triangles), _pointer(d.spheres), len(d.spheres),
                                            _pointer(d.materials), len(d.materials), _pointer(d.point_lights), len(d.point_lights),
                                            ctypes.byref(nodes), ctypes.byref(node_count), ctypes.byref(tokens), ctypes.byref(paths),
                                            ctypes.byref(emitter_count), ctypes.byref(power))
    if status != 0:
        raise ValueError(f"echo_host_build_light_tree failed with status {status}")
    return (_take(nodes, node_count.value, structs.LIGHT_NODE), _take(tokens, emitter_count.value, np.uint32),
            _take(paths, emitter_count.value, np.uint64), float(power.value))


def prepare(description, threads=0):
    """ScenePreparer.Prepare -> PreparedScene constructor (PreparedScene.cs:26-40) for a scene without instances."""
    lib = _library()
    d = description
    d.triangles = np.ascontiguousarray(d.triangles, dtype=structs.TRIANGLE)
    d.spheres = np.ascontiguousarray(d.spheres, dtype=structs.SPHERE)
    d.materials = np.ascontiguousarray(d.materials, dtype=structs.MATERIAL)
    d.point_lights = np.ascontiguousarray(d.point_lights, dtype=structs.POINT_LIGHT)
    d.infinite_lights = np.ascontiguousarray(d.infinite_lights, dtype=structs.INFINITE_LIGHT)

    nodes, max_depth = build_qbvh(d.triangles, d.spheres, threads)
    light_nodes, tokens, paths, scene_power = build_light_tree(d)

    # FilterLights / SumInfiniteLightsPower / CalculateThreshold (PreparedScene.cs:279-325)
    infinite_power = np.float32(0)
    keep = []
    for i, light in enumerate(d.infinite_lights):
        radiance = np.ascontiguousarray(light["radiance"], dtype=np.float32)
        power = lib.echo_host_ambient_power(_pointer(radiance), _pointer(nodes[:1]))
        if power >= 8e-7:
            keep.append(i)
            infinite_power = np.float32(infinite_power + np.float32(power))
    d.infinite_lights = np.ascontiguousarray(d.infinite_lights[keep])

    if len(d.infinite_lights) == 0 and not scene_power >= 8e-7:
        # "Degenerate case with literally zero light contributor": a black ambient light (PreparedScene.cs:293-305)
        d.infinite_lights = np.zeros(1, dtype=structs.INFINITE_LIGHT)
        d.infinite_lights["directlyVisible"] = 1

    threshold = float(lib.echo_host_infinite_threshold(float(infinite_power), scene_power))
    count = len(d.infinite_lights)
    pdf = float(np.float32(threshold) / np.float32(count)) if count else float("nan")
    if count == 0:
        pdf = 0.0  # never read: Pick only takes the infinite branch when sample < threshold == 0

    return PreparedArrays(d, nodes, max_depth, light_nodes, tokens, paths, threshold, pdf, scene_power)
