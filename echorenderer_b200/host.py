"""Host-side scene preparation: the Python face of libecho_host.so (include/echo_host.h).

Mirrors what Echo's C# host does once per render (ScenePreparer -> PreparedPack, Aggregation/Preparation/PreparedPack.cs:17-25):
build the SweepBuilder hierarchy, collapse it into the QuadBoundingVolumeHierarchy node array, build the LightTree and
compute the infinite-light threshold (PreparedScene.cs:34-39). The results are the arrays echo_b200.h takes.
"""
import ctypes
import math
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
    lib.echo_host_build_qbvh_instanced.argtypes = [p, u32, p, u32, p, u32, ctypes.c_int32, ctypes.POINTER(p), ctypes.POINTER(u32), ctypes.POINTER(u32)]
    lib.echo_host_build_qbvh_instanced.restype = ctypes.c_int32
    lib.echo_host_build_light_tree_instanced.argtypes = [p, u32, p, u32, p, u32, p, u32, p, u32, ctypes.POINTER(p), ctypes.POINTER(u32),
                                                         ctypes.POINTER(p), ctypes.POINTER(p), ctypes.POINTER(u32), ctypes.POINTER(ctypes.c_float)]
    lib.echo_host_build_light_tree_instanced.restype = ctypes.c_int32
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
class InstanceDescription:
    """Scenic/Hierarchies/PackInstance.cs: one placement of an EntityPack. `pack` indexes SceneDescription.packs (0-based);
    rotation is Euler degrees, scale is uniform; `materials` optionally replaces the pack's swatch for this placement."""
    pack: int
    position: tuple = (0.0, 0.0, 0.0)
    rotation: tuple = (0.0, 0.0, 0.0)
    scale: float = 1.0
    materials: np.ndarray = None
    material_textures: np.ndarray = None  # texture slots of the replacement swatch


@dataclass
class PackDescription:
    """Scenic/Hierarchies/EntityPack.cs: geometry that is prepared once (PreparedPack) and instanced any number of times."""
    triangles: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=structs.TRIANGLE))
    spheres: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=structs.SPHERE))
    materials: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=structs.MATERIAL))
    instances: list = field(default_factory=list)
    point_lights: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=structs.POINT_LIGHT))
    material_textures: np.ndarray = None


@dataclass
class TextureDescription:
    """Textures/Grids/TextureGrid.cs: texels[height, width, 4] RGBA128 (row 0 at the bottom), IFilter and IWrapper."""
    texels: np.ndarray
    filter: int = structs.FILTER_BILINEAR   # TextureGrid.Filter defaults to bilinear (TextureGrid.cs:35)
    wrapper: int = structs.WRAPPER_CLAMP    # TextureGrid.Wrapper defaults to clamp (TextureGrid.cs:34)


@dataclass
class SceneDescription:
    """What Echo's Scene + ScenePreparer hand to PreparedScene, with constant (Pure) textures flattened."""
    triangles: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=structs.TRIANGLE))
    spheres: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=structs.SPHERE))
    materials: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=structs.MATERIAL))
    instances: list = field(default_factory=list)  # InstanceDescription placed directly in the scene
    packs: list = field(default_factory=list)      # PackDescription referred to by instances
    textures: list = field(default_factory=list)   # TextureDescription, shared by every pack of the scene
    material_textures: np.ndarray = None           # structs.MATERIAL_TEXTURES per material; None = constants only
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
    # instanced scenes: every pack's arrays back to back (include/echo_b200.h EchoPack); None for a single pack
    packs: np.ndarray = None
    instances: np.ndarray = None
    all_triangles: np.ndarray = None
    all_spheres: np.ndarray = None
    all_materials: np.ndarray = None
    all_point_lights: np.ndarray = None
    textures: np.ndarray = None           # structs.TEXTURE records, None without image textures
    texels: np.ndarray = None             # [n, 4] float32
    material_textures: np.ndarray = None  # structs.MATERIAL_TEXTURES per entry of `materials`
    distributions: np.ndarray = None      # DiscreteDistribution2D cdf values of the environment lights
    _bound_radius: float = None           # cache of bound_radius

    @property
    def point_lights(self):
        return self.description.point_lights if self.all_point_lights is None else self.all_point_lights

    @property
    def triangles(self):
        return self.description.triangles if self.all_triangles is None else self.all_triangles

    @property
    def spheres(self):
        return self.description.spheres if self.all_spheres is None else self.all_spheres

    @property
    def materials(self):
        return self.description.materials if self.all_materials is None else self.all_materials

    @property
    def bound_radius(self):
        """Accelerator.SphereBound.radius (Accelerator.cs:43-63): the near-minimal sphere around the corners of the bounds two
        quad levels below the root — what the auxiliary evaluators report for escaped rays and the infinite lights' power uses."""
        if self._bound_radius is None:
            self._bound_radius = float(accelerator_sphere_bound(self.nodes).radius)
        return self._bound_radius

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


def build_qbvh(triangles, spheres, threads=0, instance_bounds=None):
    """SweepBuilder + QuadBoundingVolumeHierarchy constructor (SweepBuilder.cs:24-36, QuadBoundingVolumeHierarchy.cs:24-36).
    instance_bounds: [n, 6] float32 boxes (min xyz, max xyz) of the pack's instances, tokenized after the spheres."""
    lib = _library()
    triangles = np.ascontiguousarray(triangles, dtype=structs.TRIANGLE)
    spheres = np.ascontiguousarray(spheres, dtype=structs.SPHERE)
    boxes = np.zeros((0, 6), dtype=np.float32) if instance_bounds is None else np.ascontiguousarray(instance_bounds, dtype=np.float32).reshape(-1, 6)
    nodes, count, depth = ctypes.c_void_p(), ctypes.c_uint32(), ctypes.c_uint32()
    status = lib.echo_host_build_qbvh_instanced(_pointer(triangles), len(triangles), _pointer(spheres), len(spheres), _pointer(boxes), len(boxes),
                                                threads, ctypes.byref(nodes), ctypes.byref(count), ctypes.byref(depth))
    if status != 0:
        raise ValueError(f"echo_host_build_qbvh failed with status {status} (needs 2..2^28-1 primitives)")
    return _take(nodes, count.value, structs.QBVH_NODE), int(depth.value)


def fill_bounds(nodes, depth=6):
    """QuadBoundingVolumeHierarchy.FillBounds (QuadBoundingVolumeHierarchy.cs:62-118): the child boxes found `depth // 2`
    quad levels below the root (leaves met earlier are kept), as [n, 6] min/max rows."""
    iteration = depth // 2 - 1
    boxes, stack = [], [0]

    def bound(node, j):
        return [node["minX"][j], node["minY"][j], node["minZ"][j], node["maxX"][j], node["maxY"][j], node["maxZ"][j]]

    for _ in range(iteration):
        following = []
        while stack:
            node = nodes[stack.pop()]
            for j in range(4):
                child = int(node["token4"][j])
                if child == structs.TOKEN_EMPTY:
                    continue
                if structs.token_type(child) == structs.TOKEN_TYPE_NODE:
                    following.append(structs.token_index(child))
                else:
                    boxes.append(bound(node, j))
        stack = following

    while stack:
        node = nodes[stack.pop()]
        for j in range(4):
            if int(node["token4"][j]) != structs.TOKEN_EMPTY:
                boxes.append(bound(node, j))

    return np.asarray(boxes, dtype=np.float32).reshape(-1, 6)


def transformed_bound(boxes, matrix):
    """BoxBound(ReadOnlySpan<BoxBound>, Float4x4) (BoxBound.cs:39-57): centre/extent transform of every box, then the union."""
    matrix = np.asarray(matrix, dtype=np.float32).reshape(3, 4)
    center = (boxes[:, 3:] + boxes[:, :3]) / np.float32(2)
    extend = (boxes[:, 3:] - boxes[:, :3]) / np.float32(2)
    center = (center @ matrix[:, :3].T + matrix[:, 3]).astype(np.float32)
    extend = (extend @ np.abs(matrix[:, :3]).T).astype(np.float32)
    return np.concatenate([(center - extend).min(axis=0), (center + extend).max(axis=0)]).astype(np.float32)


def instance_matrices(instance):
    """PreparedInstance constructor (PreparedInstance.cs:15-27) from a PackInstance's position / rotation / uniform scale:
    inverseTransform = Float4x4.Transformation (local -> parent), forwardTransform = its inverse, and the two scale multipliers."""
    from .scenes import rotation_matrix
    rotation = np.asarray(rotation_matrix(*instance.rotation), dtype=np.float64)
    inverse = np.eye(4)
    inverse[:3, :3] = rotation * float(instance.scale)
    inverse[:3, 3] = instance.position
    inverse = inverse.astype(np.float32)
    forward = np.linalg.inv(inverse.astype(np.float64)).astype(np.float32)
    x, y, z = inverse[0, :3]
    inverse_scale = np.sqrt((x * x + y * y) + (z * z + np.float32(0)), dtype=np.float32)  # GetRow(0).XYZ_.Magnitude, Float4.cs:51-61,73-81
    forward_scale = np.float32(1) / inverse_scale
    return forward[:3].reshape(-1), inverse[:3].reshape(-1), forward_scale, inverse_scale


def build_light_tree(description, instance_lights=None):
    """LightCollection.CreateBounds + LightTree constructor (LightCollection.cs:91-137, LightTree.cs:21-38).
    instance_lights: [n, 12] float32 PreparedInstance.LightBound rows (box min, box max, cone axis, cosOffset, cosExtend, power)."""
    lib = _library()
    d = description
    triangles = np.ascontiguousarray(d.triangles, dtype=structs.TRIANGLE)
    spheres = np.ascontiguousarray(d.spheres, dtype=structs.SPHERE)
    materials = np.ascontiguousarray(d.materials, dtype=structs.MATERIAL)
    points = np.ascontiguousarray(d.point_lights, dtype=structs.POINT_LIGHT)
    bounds = np.zeros((0, 12), dtype=np.float32) if instance_lights is None else np.ascontiguousarray(instance_lights, dtype=np.float32).reshape(-1, 12)
    nodes, tokens, paths = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    node_count, emitter_count, power = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_float()
    status = lib.echo_host_build_light_tree_instanced(_pointer(triangles), len(triangles), _pointer(spheres), len(spheres),
                                                      _pointer(materials), len(materials), _pointer(points), len(points), _pointer(bounds), len(bounds),
                                                      ctypes.byref(nodes), ctypes.byref(node_count), ctypes.byref(tokens), ctypes.byref(paths),
                                                      ctypes.byref(emitter_count), ctypes.byref(power))
    if status != 0:
        raise ValueError(f"echo_host_build_light_tree failed with status {status}")
    return (_take(nodes, node_count.value, structs.LIGHT_NODE), _take(tokens, emitter_count.value, np.uint32),
            _take(paths, emitter_count.value, np.uint64), float(power.value))


def _prepare_packs(description, threads, tree_builder=None, light_tree_builder=None):
    """ScenePreparer: every EntityPack becomes one PreparedPack (children before parents), then all arrays are laid back to
    back with an EchoPack record per pack (pack 0 = the scene) and an EchoInstance record per placement."""
    sources = [description] + list(description.packs)  # pack k+1 = description.packs[k]
    built = {}

    def build(index, trail=()):
        if index in built:
            return built[index]
        if index in trail or len(trail) > structs.MAX_INSTANCE_LAYERS:
            raise ValueError("instancing is cyclic or deeper than TokenHierarchy.MaxLayer")
        source = sources[index]
        boxes, records, lights = [], [], []
        for instance in source.instances:
            child = instance.pack + 1
            child_nodes, _, child_lights = build(child, trail + (index,))
            forward, inverse, forward_scale, inverse_scale = instance_matrices(instance)
            boxes.append(transformed_bound(fill_bounds(child_nodes), inverse))  # PreparedInstance.BoxBound
            records.append((forward, inverse, forward_scale, inverse_scale, child, instance))

            # PreparedInstance.LightBound / Power (PreparedInstance.cs:31-40): the pack's light-tree root moved to parent space
            light_nodes = child_lights[0]
            if len(light_nodes):
                root = light_nodes[0]
                box = transformed_bound(np.concatenate([root["boxMin"], root["boxMax"]])[None, :], inverse)
                axis = inverse.reshape(3, 4)[:, :3].astype(np.float64) @ root["coneAxis"].astype(np.float64)
                axis = (axis / max(np.linalg.norm(axis), 1e-300)).astype(np.float32)  # ConeBound operator *, ConeBound.cs:68-72
                power = np.float32(root["power"]) * inverse_scale * inverse_scale
                lights.append(np.concatenate([box, axis, [root["cosOffset"], root["cosExtend"], power]]).astype(np.float32))
            else:
                lights.append(np.zeros(12, dtype=np.float32))
        pack_boxes = np.asarray(boxes, dtype=np.float32) if boxes else None
        if tree_builder is not None and len(source.triangles) + len(source.spheres) + len(boxes) >= 2:
            nodes, depth = tree_builder(source.triangles, source.spheres, pack_boxes)
        else:
            nodes, depth = build_qbvh(source.triangles, source.spheres, threads, pack_boxes)
        built[index] = (nodes, depth, (light_tree_builder or build_light_tree)(source, np.asarray(lights, dtype=np.float32) if lights else None))
        source._records = records
        return built[index]

    build(0)
    used = sorted(built)  # packs nothing refers to are dropped
    remap = {old: new for new, old in enumerate(used)}
    packs = np.zeros(len(used), dtype=structs.PACK)
    all_nodes, all_triangles, all_spheres, all_materials, all_instances = [], [], [], [], []
    all_light_nodes, all_tokens, all_paths, all_points = [], [], [], []
    slot_blocks, slot_overrides = [], []

    def slots_of(source, count):
        given = getattr(source, "material_textures", None)
        return structs.material_textures(count) if given is None else np.ascontiguousarray(given, dtype=structs.MATERIAL_TEXTURES)

    counts = dict(node=0, triangle=0, sphere=0, material=0, instance=0, light=0, emitter=0, point=0)
    overrides = []

    for new, old in enumerate(used):
        source, (nodes, depth, (light_nodes, tokens, paths, _)) = sources[old], built[old]
        points = np.ascontiguousarray(source.point_lights, dtype=structs.POINT_LIGHT)
        triangles = np.ascontiguousarray(source.triangles, dtype=structs.TRIANGLE)
        spheres = np.ascontiguousarray(source.spheres, dtype=structs.SPHERE)
        materials = np.ascontiguousarray(source.materials, dtype=structs.MATERIAL)
        record = packs[new]
        record["nodeOffset"], record["nodeCount"], record["maxDepth"] = counts["node"], len(nodes), depth
        record["triangleOffset"], record["triangleCount"] = counts["triangle"], len(triangles)
        record["sphereOffset"], record["sphereCount"] = counts["sphere"], len(spheres)
        record["instanceOffset"], record["instanceCount"] = counts["instance"], len(source.instances)
        record["materialOffset"] = counts["material"]
        record["lightNodeOffset"], record["lightNodeCount"] = counts["light"], len(light_nodes)
        record["emitterOffset"], record["emitterCount"] = counts["emitter"], len(tokens)
        record["pointLightOffset"], record["pointLightCount"] = counts["point"], len(points)
        all_light_nodes.append(light_nodes), all_tokens.append(tokens), all_paths.append(paths), all_points.append(points)
        counts["light"] += len(light_nodes)
        counts["emitter"] += len(tokens)
        counts["point"] += len(points)
        slot_blocks.append(slots_of(source, len(materials)))
        all_nodes.append(nodes), all_triangles.append(triangles), all_spheres.append(spheres), all_materials.append(materials)
        counts["node"] += len(nodes)
        counts["triangle"] += len(triangles)
        counts["sphere"] += len(spheres)
        counts["material"] += len(materials)
        counts["instance"] += len(source.instances)
        all_instances.extend((new, r) for r in getattr(source, "_records", []))

    instances = np.zeros(len(all_instances), dtype=structs.INSTANCE)

    for i, (_, (forward, inverse, forward_scale, inverse_scale, child, instance)) in enumerate(all_instances):
        record = instances[i]
        record["forward"], record["inverse"] = forward, inverse
        record["forwardScale"], record["inverseScale"] = forward_scale, inverse_scale
        record["pack"] = remap[child]
        if instance.materials is None:
            record["materialOffset"] = packs[remap[child]]["materialOffset"]  # PackInstance without a swatch: the pack's own
        else:
            record["materialOffset"] = counts["material"]
            override = np.ascontiguousarray(instance.materials, dtype=structs.MATERIAL)
            overrides.append(override)
            slot_overrides.append(slots_of(instance, len(override)))
            counts["material"] += len(override)

    # OneSided.base indexes the swatch it lives in: make it absolute
    blocks = all_materials + overrides
    materials = np.concatenate(blocks) if blocks else np.zeros(0, dtype=structs.MATERIAL)
    offset = 0
    for block in blocks:
        view = materials[offset:offset + len(block)]
        view["base"][view["type"] == structs.MATERIAL_ONESIDED] += offset
        offset += len(block)

    lights = (np.concatenate(all_light_nodes), np.concatenate(all_tokens), np.concatenate(all_paths), np.concatenate(all_points), float(built[0][2][3]))
    slots = np.concatenate(slot_blocks + slot_overrides) if blocks else structs.material_textures(0)
    return (packs, instances, np.concatenate(all_nodes), np.concatenate(all_triangles), np.concatenate(all_spheres), materials,
            max(int(p["maxDepth"]) for p in packs), lights, slots)


def _wrap_index(index, size, wrapper):
    if wrapper == structs.WRAPPER_CLAMP:
        return np.clip(index, 0, size - 1)
    if wrapper == structs.WRAPPER_REPEAT:
        return np.mod(index, size)
    folded = np.mod(index, 2 * size)
    return np.minimum(folded, 2 * size - 1 - folded)


def sample_texture(texture, uv):
    """TextureGrid.this[Float2] (IFilter.cs:17-68) in float64 numpy: host-side preparation only (distributions, averages)."""
    texels = np.asarray(texture.texels, dtype=np.float64)
    height, width = texels.shape[:2]
    scaled = np.asarray(uv, dtype=np.float64) * [width, height]
    if texture.filter == structs.FILTER_POINT:
        x = _wrap_index(np.floor(scaled[..., 0]).astype(int), width, texture.wrapper)
        y = _wrap_index(np.floor(scaled[..., 1]).astype(int), height, texture.wrapper)
        return texels[y, x]
    base = np.floor(scaled - 0.5).astype(int)
    time = scaled - 0.5 - base
    x0, x1 = _wrap_index(base[..., 0], width, texture.wrapper), _wrap_index(base[..., 0] + 1, width, texture.wrapper)
    y0, y1 = _wrap_index(base[..., 1], height, texture.wrapper), _wrap_index(base[..., 1] + 1, height, texture.wrapper)
    tx, ty = time[..., :1], time[..., 1:]
    low = texels[y0, x0] * (1 - tx) + texels[y0, x1] * tx
    high = texels[y1, x0] * (1 - tx) + texels[y1, x1] * tx
    return low * (1 - ty) + high * ty


def _distribution_1d(values):
    """DiscreteDistribution1D constructor (DiscreteDistribution1D.cs:12-56): returns (cdfValues float32, sum float32)."""
    values = np.asarray(values, dtype=np.float32)
    length = len(values)
    count_r = np.float32(1) / np.float32(length)
    cdf = np.cumsum(values.astype(np.float64)).astype(np.float32)
    total = np.float32(np.sum(values.astype(np.float64)))

    if not total >= np.float32(8e-7):  # FastMath.AlmostZero(sum)
        cdf = (np.arange(length, dtype=np.float64) * np.float64(count_r) + np.float64(count_r)).astype(np.float32)  # FastMath.FMA(i, countR, countR)
        total = np.float32(0)
    else:
        cdf = (cdf * (np.float32(1) / total)).astype(np.float32)

    index = length - 1
    last = cdf[index]
    while True:  # "Assign the last identical values to one to ensure no leaking when sampling"
        cdf[index] = 1.0
        index -= 1
        if not (index > 0 and last == cdf[index]):
            break
    return cdf, total


def build_environment(texture):
    """CylindricalTexture.Prepare (CylindricalTexture.cs:33-96): the sin-weighted luminance of every cell of the grid, the
    DiscreteDistribution2D over it (vertical cdf, then one cdf per row) and the texture's Average."""
    texels = np.asarray(texture.texels, dtype=np.float64)
    height, width = texels.shape[:2]
    ys, xs = np.meshgrid(np.arange(height + 1), np.arange(width + 1), indexing="ij")
    corners = sample_texture(texture, np.stack([xs / width, ys / height], axis=-1))[..., :3]  # texture[(x, y) * sizeR]
    sines = np.clip(np.sin(np.pi * np.arange(height + 1) / height), 0.0, 1.0)
    weighted = corners * sines[:, None, None]
    grab = weighted[:-1] + weighted[1:]                 # Grab(x): lower * sin0 + upper * sin1, per row y
    average = (grab[:, :-1] + grab[:, 1:]) / 4.0         # the four corners of every cell
    weights = (average[..., 0] * 0.212671 + average[..., 1] * 0.715160 + average[..., 2] * 0.072169).astype(np.float32)

    rows, sums = [], []
    for y in range(height):
        cdf, total = _distribution_1d(weights[y])
        rows.append(cdf)
        sums.append(total)
    vertical, _ = _distribution_1d(np.asarray(sums, dtype=np.float32))
    values = np.concatenate([vertical] + rows).astype(np.float32)
    mean = (average.reshape(-1, 3).sum(axis=0) * np.pi / 2 / (width * height)).astype(np.float32)
    return values, mean


def cubemap_average(faces, count=128):
    """Stand-in for Cubemap.Prepare's AverageConverge (a seeded-by-time Monte Carlo, IDirectionalTexture.cs:28-63): the mean of
    Cubemap.Evaluate over a stratified set of directions."""
    v, u = np.meshgrid((np.arange(count) + 0.5) / count, (np.arange(2 * count) + 0.5) / (2 * count), indexing="ij")
    z = 1 - 2 * v
    r = np.sqrt(np.maximum(0, 1 - z * z))
    d = np.stack([r * np.cos(2 * np.pi * u), r * np.sin(2 * np.pi * u), z], axis=-1).reshape(-1, 3)
    axis = np.argmax(np.abs(d), axis=1)
    part = d[np.arange(len(d)), axis]
    index = axis * 2 + (part < 0)
    uv = np.zeros((len(d), 2))
    table = {0: (-d[:, 2], d[:, 1]), 1: (d[:, 2], d[:, 1]), 2: (d[:, 0], -d[:, 2]), 3: (d[:, 0], d[:, 2]), 4: (d[:, 0], d[:, 1]), 5: (-d[:, 0], d[:, 1])}
    total = np.zeros(3)
    for face in range(6):
        mask = index == face
        if not mask.any():
            continue
        a, b = table[face]
        uv = np.stack([a[mask], b[mask]], axis=-1) * (0.5 / np.abs(part[mask]))[:, None] + 0.5
        total += sample_texture(faces[face], uv)[..., :3].sum(axis=0)
    return (total / len(d)).astype(np.float32)


# ---- Aggregation/Bounds/SphereBound.cs: near-minimal bounding sphere (extremal points + exact solver + grow), float32 like the reference ----
_F = np.float32


def _sq(v):  # Float3.SquaredMagnitude, X * X + Y * Y + Z * Z
    return _F(_F(_F(v[0] * v[0]) + _F(v[1] * v[1])) + _F(v[2] * v[2]))


def _cross(a, b):  # Float3.Cross rounds fp64 lanes once (Float3.cs:268-273)
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    return np.array([a64[1] * b64[2] - a64[2] * b64[1], a64[2] * b64[0] - a64[0] * b64[2], a64[0] * b64[1] - a64[1] * b64[0]]).astype(np.float32)


def _determinant(m):  # Float4x4.Determinant, Float4x4.cs:95-121
    (f00, f01, f02, f03), (f10, f11, f12, f13), (f20, f21, f22, f23), (f30, f31, f32, f33) = [[_F(x) for x in row] for row in m]
    d21_32, d21_33, d22_33 = f21 * f32 - f22 * f31, f21 * f33 - f23 * f31, f22 * f33 - f23 * f32
    d20_31, d20_32, d20_33 = f20 * f31 - f21 * f30, f20 * f32 - f22 * f30, f20 * f33 - f23 * f30
    m00 = f11 * d22_33 - f12 * d21_33 + f13 * d21_32
    m01 = f10 * d22_33 - f12 * d20_33 + f13 * d20_32
    m02 = f10 * d21_33 - f11 * d20_33 + f13 * d20_31
    m03 = f10 * d21_32 - f11 * d20_32 + f12 * d20_31
    return _F(f00 * m00 - f01 * m01 + f02 * m02 - f03 * m03)


def _sphere_normals():
    """`new Versor(45f, 45f, 45f) * {Right, Up, Forward}` (SphereBound.cs:50-63, Versor.cs:21-35,223-240)."""
    half = _F(45.0) * _F(_F(math.pi / 180.0) * _F(0.5))
    s, c = _F(math.sin(half)), _F(math.cos(half))
    d = np.array([s * c * c + c * s * s, c * s * c - s * c * s, c * c * s - s * s * c, c * c * c + s * s * s], dtype=np.float32)
    dd = d * d
    dw = (d[3] * _F(2)) * d[:3]
    dz = (d[2] * _F(2)) * d[:2]
    dy_x = d[1] * _F(2) * d[0]

    def rotate(v):
        x, y, z = (_F(k) for k in v)
        return np.array([
            dd[3] * x + dd[0] * x - dw[2] * y + dy_x * y + dw[1] * z + dz[0] * z - dd[2] * x - dd[1] * x,
            dy_x * x + dw[2] * x + dd[1] * y - dd[2] * y + dz[1] * z - dw[0] * z + dd[3] * y - dd[0] * y,
            dz[0] * x - dw[1] * x + dz[1] * y + dw[0] * y + dd[2] * z - dd[1] * z - dd[0] * z + dd[3] * z], dtype=np.float32)

    return [rotate(axis) for axis in ((1, 0, 0), (0, 1, 0), (0, 0, 1))]


_SPHERE_NORMALS = _sphere_normals()


class SphereBound:
    """Aggregation/Bounds/SphereBound.cs: `SphereBound(ReadOnlySpan<Float3> points)` -> center, radius."""

    def __init__(self, points):
        points = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        assert len(points) > 0
        if len(points) > len(_SPHERE_NORMALS) * 2:
            center, radius2 = self._solve_exact(self._extremes(points))
            center, radius2 = self._grow(points, center, radius2)
        else:
            center, radius2 = self._solve_exact(points)
        self.center = center
        radius = _F(np.sqrt(radius2)) if radius2 > 0 else _F(0)  # FastMath.Sqrt0
        self.radius = _F(radius * _F(_F(1) + _F(8e-7)))          # "increase the radius of the sphere by an epsilon", :37-38

    def contains(self, point):
        return _sq(np.asarray(point, dtype=np.float32) - self.center) <= self.radius * self.radius

    @staticmethod
    def _contains(point, center, radius2):
        return _sq(point - center) <= radius2

    @staticmethod
    def _extremes(points):  # FillExtremes, :75-107: first minimum / maximum along each normal
        out = []
        for normal in _SPHERE_NORMALS:
            values = (points[:, 0] * normal[0] + points[:, 1] * normal[1]) + points[:, 2] * normal[2]  # Float3.Dot
            out += [points[int(np.argmin(values))], points[int(np.argmax(values))]]
        return np.array(out, dtype=np.float32)

    @classmethod
    def _grow(cls, points, center, radius2):  # GrowSphere, :109-125
        for current in points:
            if cls._contains(current, center, radius2):
                continue
            offset = center - current
            length = _F(np.sqrt(_sq(offset)))
            radius = _F(_F(_F(np.sqrt(radius2)) + length) / _F(2))
            center = (current + (offset / length) * radius).astype(np.float32)
            radius2 = _F(radius * radius)
        return center, radius2

    @staticmethod
    def _diameter(a, b):  # SolveFromDiameterPoints, :273-277
        return ((a + b) / _F(2)).astype(np.float32), _F(_sq(a - b) / _F(4))

    @classmethod
    def _triangle(cls, a, b, c):  # SolveFromTriangle, :201-224
        pba, pca = b - a, c - a
        normal = _cross(pba, pca)
        magnitude2 = _sq(normal)
        if magnitude2 > 0:
            center = (_cross((_sq(pba) * pca - _sq(pca) * pba).astype(np.float32), normal) / magnitude2 / _F(2) + a).astype(np.float32)
            return center, _sq(a - center)
        if (pba[0] * pca[0] + pba[1] * pca[1]) + pba[2] * pca[2] > 0:
            return cls._diameter(a, b)
        return cls._diameter(b, c)

    @staticmethod
    def _circumsphere(a, b, c, d):  # SolveCircumSphereFromFourExtremes, :226-268
        rows = [a, b, c, d]
        squares = [_sq(p) for p in rows]
        one = _F(1)
        det_r = one / _determinant([[p[0], p[1], p[2], one] for p in rows])
        center = (det_r / _F(2)) * np.array([
            _determinant([[q, p[1], p[2], one] for p, q in zip(rows, squares)]),
            _determinant([[p[0], q, p[2], one] for p, q in zip(rows, squares)]),
            _determinant([[p[0], p[1], q, one] for p, q in zip(rows, squares)])], dtype=np.float32)
        center = center.astype(np.float32)
        return center, _sq(a - center)

    @classmethod
    def _solve_exact(cls, points):  # SolveExact, :130-149
        if len(points) == 1:
            return points[0].copy(), _F(0)
        center, radius2 = cls._diameter(points[0], points[1])
        for index in range(2, len(points)):
            if cls._contains(points[index], center, radius2):
                continue
            center, radius2 = cls._solve_recursive(points, index, index)
        return center, radius2

    @classmethod
    def _solve_recursive(cls, points, end, pin1, pin2=-1, pin3=-1):  # SolveExactRecursive, :154-197
        index = 0
        if pin2 < 0:
            center, radius2 = cls._diameter(points[0], points[pin1])
            index = 1
        elif pin3 < 0:
            center, radius2 = cls._diameter(points[pin1], points[pin2])
        else:
            center, radius2 = cls._triangle(points[pin1], points[pin2], points[pin3])

        while index < end:
            if not cls._contains(points[index], center, radius2):
                if pin2 < 0:
                    center, radius2 = cls._solve_recursive(points, index, pin1, index)
                elif pin3 < 0:
                    center, radius2 = cls._solve_recursive(points, index, pin1, pin2, index)
                else:
                    center, radius2 = cls._circumsphere(points[pin1], points[pin2], points[pin3], points[index])
            index += 1
        return center, radius2


def box_vertices(low, high):
    """BoxBound.FillVertices (BoxBound.cs:140-156), in the reference's order."""
    (x0, y0, z0), (x1, y1, z1) = low, high
    return np.array([(x0, y0, z0), (x1, y1, z1), (x0, y0, z1), (x0, y1, z0), (x1, y0, z0), (x1, y1, z0), (x1, y0, z1), (x0, y1, z1)], dtype=np.float32)


def accelerator_sphere_bound(nodes):
    """Accelerator.SphereBound (Accelerator.cs:43-63) over QuadBoundingVolumeHierarchy.FillBounds(6, ...) (:62-118): the child
    bounds two quad levels below the root (leaf children met on the way are taken as they are), eight corners each."""
    stack0, stack1, boxes = [0], [], []

    def child_box(node, j):
        return (node["minX"][j], node["minY"][j], node["minZ"][j]), (node["maxX"][j], node["maxY"][j], node["maxZ"][j])

    for _ in range(6 // 2 - 1):
        while stack0:
            node = nodes[stack0.pop()]
            for j in range(4):
                child = int(node["token4"][j])
                if child == structs.TOKEN_EMPTY:
                    continue
                if structs.token_type(child) == structs.TOKEN_TYPE_NODE:
                    stack1.append(int(structs.token_index(child)))
                else:
                    boxes.append(child_box(node, j))
        stack0, stack1 = stack1, stack0

    while stack0:
        node = nodes[stack0.pop()]
        for j in range(4):
            if int(node["token4"][j]) != structs.TOKEN_EMPTY:
                boxes.append(child_box(node, j))

    points = np.concatenate([box_vertices(low, high) for low, high in boxes])
    return SphereBound(points)


def _root_bound_radius(nodes):
    """Accelerator.SphereBound.radius of the scene's accelerator (AmbientLight.cs:47, DirectionalLight.cs:72)."""
    return accelerator_sphere_bound(nodes).radius


def prepare(description, threads=0, tree=None, tree_builder=None, light_tree_builder=None):
    """ScenePreparer.Prepare -> PreparedScene constructor (PreparedScene.cs:26-40). `tree` = (nodes, max_depth) replaces the
    SweepBuilder mirror for a scene without instances; `tree_builder(triangles, spheres, instance_bounds) -> (nodes, max_depth)`
    replaces it for every pack of any scene (e.g. the device-side build: lambda t, s, b: scene.build_qbvh_device(t, s, instance_bounds=b));
    `light_tree_builder(description, instance_lights) -> (nodes, tokens, paths, power)` replaces build_light_tree the same way
    (the device-side build: scene.build_light_tree_device)."""
    lib = _library()
    d = description
    d.triangles = np.ascontiguousarray(d.triangles, dtype=structs.TRIANGLE)
    d.spheres = np.ascontiguousarray(d.spheres, dtype=structs.SPHERE)
    d.materials = np.ascontiguousarray(d.materials, dtype=structs.MATERIAL)
    d.point_lights = np.ascontiguousarray(d.point_lights, dtype=structs.POINT_LIGHT)
    d.infinite_lights = np.ascontiguousarray(d.infinite_lights, dtype=structs.INFINITE_LIGHT)

    instanced = bool(d.instances)
    if instanced:
        packs, instances, nodes, all_triangles, all_spheres, all_materials, max_depth, lights, all_slots = _prepare_packs(d, threads, tree_builder, light_tree_builder)
        light_nodes, tokens, paths, all_points, scene_power = lights
    else:
        nodes, max_depth = tree if tree is not None else (tree_builder(d.triangles, d.spheres, None) if tree_builder is not None else build_qbvh(d.triangles, d.spheres, threads))
        light_nodes, tokens, paths, scene_power = (light_tree_builder or build_light_tree)(d)

    # FilterLights / SumInfiniteLightsPower / CalculateThreshold (PreparedScene.cs:279-325)
    infinite_power = np.float32(0)
    keep = []
    distributions = []
    for i, light in enumerate(d.infinite_lights):
        if light["type"] == structs.INFINITE_CUBEMAP:
            first = int(light["texture"])
            mean = cubemap_average(d.textures[first:first + 6])
            luminance = lambda c: (np.float32(c[0]) * np.float32(0.212671) + np.float32(c[1]) * np.float32(0.715160)) + (np.float32(c[2]) * np.float32(0.072169) + np.float32(0))
            radius = max(np.float32(_root_bound_radius(nodes)), np.float32(1))
            power = float(np.float32(math.pi) * radius * radius * (luminance(mean) * luminance(light["radiance"])))
        elif light["type"] == structs.INFINITE_ENVIRONMENT:
            # AmbientLight.Prepare (AmbientLight.cs:36-46) over a CylindricalTexture: pi r^2 * Average.Luminance * Intensity.Luminance
            values, mean = build_environment(d.textures[int(light["texture"])])
            d.infinite_lights[i]["distribution"] = sum(len(v) for v in distributions)
            distributions.append(values)
            luminance = lambda c: (np.float32(c[0]) * np.float32(0.212671) + np.float32(c[1]) * np.float32(0.715160)) + (np.float32(c[2]) * np.float32(0.072169) + np.float32(0))
            radius = max(np.float32(_root_bound_radius(nodes)), np.float32(1))
            power = float(np.float32(math.pi) * radius * radius * (luminance(mean) * luminance(light["radiance"])))
        elif light["type"] == structs.INFINITE_DIRECTIONAL:
            # DirectionalLight.Prepare (DirectionalLight.cs:71-74): half of the scene's bounding disk area
            r, g, b = (np.float32(c) for c in light["intensity"])
            luminance = (r * np.float32(0.212671) + g * np.float32(0.715160)) + (b * np.float32(0.072169) + np.float32(0))
            radius = np.float32(_root_bound_radius(nodes))
            power = float(luminance * np.float32(math.pi * 0.5) * radius * radius)
        else:
            # AmbientLight.Prepare (AmbientLight.cs:42-51) over a constant texture: pi * max(r, 1)^2 * luminance
            r, g, b = (np.float32(c) for c in light["radiance"])
            luminance = (r * np.float32(0.212671) + g * np.float32(0.715160)) + (b * np.float32(0.072169) + np.float32(0))
            radius = max(np.float32(_root_bound_radius(nodes)), np.float32(1))
            power = float(np.float32(math.pi) * radius * radius * luminance)
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

    result = PreparedArrays(d, nodes, max_depth, light_nodes, tokens, paths, threshold, pdf, scene_power)
    if instanced:
        result.packs, result.instances = packs, instances
        result.all_triangles, result.all_spheres, result.all_materials = all_triangles, all_spheres, all_materials
        result.all_point_lights = all_points

    result.distributions = np.concatenate(distributions) if distributions else None

    if d.textures:
        records = np.zeros(len(d.textures), dtype=structs.TEXTURE)
        grids, offset = [], 0
        for record, texture in zip(records, d.textures):
            texels = np.ascontiguousarray(texture.texels, dtype=np.float32)
            height, width = texels.shape[:2]
            record["width"], record["height"], record["texelOffset"] = width, height, offset
            record["filter"], record["wrapper"] = texture.filter, texture.wrapper
            grids.append(texels.reshape(-1, 4))
            offset += width * height
        result.textures, result.texels = records, np.ascontiguousarray(np.concatenate(grids))
        if instanced:
            result.material_textures = all_slots
        else:
            given = d.material_textures
            result.material_textures = structs.material_textures(len(d.materials)) if given is None else np.ascontiguousarray(given, dtype=structs.MATERIAL_TEXTURES)
    return result
