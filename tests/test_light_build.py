"""The level-synchronous LightTree.Build (echorenderer_b200/csrc/echo_light_build.h: the passes and the driver of the device build,
lightbuild.cu) on a sequential CPU backend (tests/c_client/light_emulation.cpp): the tree and the emitter map it emits must be the
host mirror's — LightTree.cs:62-113 recursive, one node at a time — byte for byte: same nodes, same pre-order, same bit paths.
The GPU suite (test_gpu_build.py) asks the same of the CUDA backend. Also here: the accuracy of the pinned transcendentals the
builds share (the reference's come from the platform's C runtime), and the invariants of the emitted tree."""
import ctypes
import math
import os
import subprocess
from types import SimpleNamespace

import numpy as np
import pytest

from echorenderer_b200 import host, scenes, structs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def library(tmp_path_factory):
    path = tmp_path_factory.mktemp("light") / "liblight_emulation.so"
    subprocess.run(["g++", "-std=c++17", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC",
                    os.path.join(ROOT, "tests", "c_client", "light_emulation.cpp"), "-o", str(path)], check=True)
    lib = ctypes.CDLL(str(path))
    p, u32 = ctypes.c_void_p, ctypes.c_uint32
    lib.light_emulation_build.argtypes = [p, u32, p, u32, p, u32, p, u32, p, u32, ctypes.c_int32, p, ctypes.POINTER(u32), p, p, ctypes.POINTER(u32), ctypes.POINTER(u32)]
    lib.light_emulation_build.restype = ctypes.c_int32
    lib.light_emulation_acos_double.argtypes, lib.light_emulation_acos_double.restype = [ctypes.c_double], ctypes.c_double
    lib.light_emulation_acos.argtypes, lib.light_emulation_acos.restype = [ctypes.c_float], ctypes.c_float
    lib.light_emulation_cos.argtypes, lib.light_emulation_cos.restype = [ctypes.c_float], ctypes.c_float
    return lib


def describe(triangles=None, spheres=None, materials=None, point_lights=None):
    """what host.build_light_tree reads of a SceneDescription"""
    return SimpleNamespace(triangles=np.zeros(0, dtype=structs.TRIANGLE) if triangles is None else triangles,
                           spheres=np.zeros(0, dtype=structs.SPHERE) if spheres is None else spheres,
                           materials=np.zeros(0, dtype=structs.MATERIAL) if materials is None else materials,
                           point_lights=np.zeros(0, dtype=structs.POINT_LIGHT) if point_lights is None else point_lights)


def emulate(lib, description, instance_lights=None, reverse=False):
    d = description
    triangles = np.ascontiguousarray(d.triangles, dtype=structs.TRIANGLE)
    spheres = np.ascontiguousarray(d.spheres, dtype=structs.SPHERE)
    materials = np.ascontiguousarray(d.materials, dtype=structs.MATERIAL)
    points = np.ascontiguousarray(d.point_lights, dtype=structs.POINT_LIGHT)
    bounds = np.zeros((0, 12), dtype=np.float32) if instance_lights is None else np.ascontiguousarray(instance_lights, dtype=np.float32).reshape(-1, 12)
    candidates = len(triangles) + len(spheres) + len(points) + len(bounds)
    nodes = np.zeros(max(2 * candidates, 1), dtype=structs.LIGHT_NODE)
    tokens, paths = np.zeros(max(candidates, 1), dtype=np.uint32), np.zeros(max(candidates, 1), dtype=np.uint64)
    node_count, emitter_count, levels = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
    status = lib.light_emulation_build(triangles.ctypes.data, len(triangles), spheres.ctypes.data, len(spheres), materials.ctypes.data, len(materials),
                                       points.ctypes.data, len(points), bounds.ctypes.data, len(bounds), int(reverse), nodes.ctypes.data, ctypes.byref(node_count),
                                       tokens.ctypes.data, paths.ctypes.data, ctypes.byref(emitter_count), ctypes.byref(levels))
    return status, nodes[:node_count.value], tokens[:emitter_count.value], paths[:emitter_count.value], levels.value


def assert_same_tree(lib, description, instance_lights=None, reverse=False):
    expected_nodes, expected_tokens, expected_paths, _ = host.build_light_tree(description, instance_lights)
    status, nodes, tokens, paths, levels = emulate(lib, description, instance_lights, reverse)
    assert status == 0
    assert len(nodes) == len(expected_nodes) and len(tokens) == len(expected_tokens)
    assert nodes.tobytes() == expected_nodes.tobytes()
    assert tokens.tobytes() == expected_tokens.tobytes() and paths.tobytes() == expected_paths.tobytes()
    return nodes, tokens, paths, levels


def random_emitters(seed, triangle_count, sphere_count, point_count, scale, emissive_share=0.7):
    """a soup of small triangles and spheres, a share of them emissive (several emission colours, one too dim to count), and point lights"""
    rng = np.random.default_rng(seed)
    materials = np.concatenate([scenes.material(structs.MATERIAL_DIFFUSE, (0.7, 0.7, 0.7))]
                               + [scenes.material(structs.MATERIAL_EMISSIVE, tuple(rng.uniform(0.5, 40.0, 3))) for _ in range(5)]
                               + [scenes.material(structs.MATERIAL_EMISSIVE, (1e-9, 1e-9, 1e-9))])
    v0 = rng.uniform(-scale, scale, (triangle_count, 3))
    e1, e2 = rng.normal(0, scale * 0.02, (triangle_count, 3)), rng.normal(0, scale * 0.02, (triangle_count, 3))
    choice = np.where(rng.uniform(size=triangle_count) < emissive_share, rng.integers(1, 7, triangle_count), 0)
    triangles = scenes.make_triangles(v0, v0 + e1, v0 + e2, choice.astype(np.uint32))
    spheres = np.zeros(sphere_count, dtype=structs.SPHERE)
    spheres["position"] = rng.uniform(-scale, scale, (sphere_count, 3))
    spheres["radius"] = rng.uniform(0.01, 0.05, sphere_count) * scale
    spheres["material"] = np.where(rng.uniform(size=sphere_count) < emissive_share, rng.integers(1, 7, sphere_count), 0)
    points = np.zeros(point_count, dtype=structs.POINT_LIGHT)
    points["position"] = rng.uniform(-scale, scale, (point_count, 3))
    points["intensity"] = rng.uniform(0.1, 30.0, (point_count, 3))
    return describe(triangles, spheres, materials, points)


@pytest.mark.parametrize("fixture", ["cornell", "lights_small", "mixed_small"])
def test_emitted_tree_is_the_host_mirrors_byte_for_byte(library, fixture, request):
    prepared = request.getfixturevalue(fixture)
    description = describe(prepared.triangles, prepared.spheres, prepared.materials, prepared.point_lights)
    nodes, tokens, _, levels = assert_same_tree(library, description)
    assert len(nodes) == 2 * len(tokens) - 1 and levels >= 1
    assert nodes.tobytes() == prepared.light_nodes.tobytes()  # what the scene itself was prepared with
    assert_same_tree(library, description, reverse=True)  # no pass depends on the order of its indices


@pytest.mark.parametrize("seed,triangle_count,sphere_count,point_count,scale", [(1, 2, 0, 0, 1.0), (2, 0, 0, 2, 1.0), (3, 1, 1, 1, 5.0), (4, 40, 9, 3, 1.0),
                                                                               (5, 1500, 100, 20, 100.0), (6, 6000, 0, 0, 1e-2), (7, 500, 500, 500, 1e3)])
def test_random_emitters(library, seed, triangle_count, sphere_count, point_count, scale):
    description = random_emitters(seed, triangle_count, sphere_count, point_count, scale, emissive_share=1.0 if triangle_count + sphere_count < 5 else 0.7)
    nodes, tokens, paths, levels = assert_same_tree(library, description)
    assert len(tokens) >= 2 and len(np.unique(tokens)) == len(tokens) and len(np.unique(paths)) == len(paths)
    assert_same_tree(library, description, reverse=True)


def test_proper_cones(library):
    """Emitters facing almost the same way (a ceiling of lights, tilted by up to 1.5 degrees): their unions stay proper cones, so every sweep
    step runs all of ConeBound.Union — the binary64 angle, the rotation of the axis, the cosine — instead of the whole-sphere shortcut."""
    from tests.test_independent_kats import _tilted_ceiling
    description = _tilted_ceiling(scenes.many_lights_scene(light_count=700, rings=8, segments=8), 1.5)
    nodes, tokens, _, _ = assert_same_tree(library, description)
    branch = nodes["child0"] != 0xFFFFFFFF
    assert np.count_nonzero(branch & (nodes["cosOffset"] > -1) & (nodes["cosOffset"] < 1)) > 500
    assert_same_tree(library, description, reverse=True)


def test_placements_and_equal_centres(library):
    """PreparedInstance.LightBound rows join the list last (AddInstances); emitters with EQUAL centres keep their order (the stable sort)."""
    rng = np.random.default_rng(11)
    description = random_emitters(12, 60, 10, 4, 10.0)
    rows = np.zeros((9, 12), dtype=np.float32)
    low = rng.uniform(-10, 10, (9, 3))
    rows[:, 0:3], rows[:, 3:6] = low, low + rng.uniform(0.5, 3.0, (9, 3))
    axis = rng.normal(size=(9, 3))
    rows[:, 6:9] = axis / np.linalg.norm(axis, axis=1, keepdims=True)
    rows[:, 9], rows[:, 10], rows[:, 11] = rng.uniform(-1, 1, 9), rng.uniform(0, 1, 9), rng.uniform(0.5, 80.0, 9)
    rows[4, 11] = 0.0  # a placement without emitters is skipped
    nodes, tokens, _, _ = assert_same_tree(library, description, rows)
    assert np.count_nonzero(tokens >> 28 == 3) == 8

    # a lattice of identical emissive triangles: every sort meets ties, on every axis (also: -0 and +0 centres are one key)
    ix, iy = np.meshgrid(np.arange(-4, 4), np.arange(-2, 2), indexing="ij")
    v0 = np.stack([ix.reshape(-1), np.zeros(ix.size), iy.reshape(-1)], axis=-1).astype(np.float64)
    v0 = np.concatenate([v0, v0, -v0 * 0.0])
    materials = np.concatenate([scenes.material(structs.MATERIAL_EMISSIVE, (3.0, 2.0, 1.0))])
    lattice = describe(scenes.make_triangles(v0 - (0.25, 0, 0.25), v0 + (0.25, 0, -0.25), v0 + (-0.25, 0, 0.25), 0), materials=materials)
    assert_same_tree(library, lattice)
    assert_same_tree(library, lattice, reverse=True)


def test_degenerate_counts(library):
    """no emitter: an empty tree; one emitter: the root is its leaf (LightTree.cs:64-65)"""
    status, nodes, tokens, _, _ = emulate(library, describe(materials=np.concatenate([scenes.material(structs.MATERIAL_DIFFUSE)])))
    assert status == 0 and len(nodes) == 0 and len(tokens) == 0
    dark = random_emitters(3, 50, 5, 0, 1.0, emissive_share=0.0)
    status, nodes, tokens, _, _ = emulate(library, dark)
    assert status == 0 and len(nodes) == 0 and len(tokens) == 0

    single = describe(point_lights=np.array([((1.0, 2.0, 3.0), (0.5, -1.0, 4.0))], dtype=structs.POINT_LIGHT))
    nodes, tokens, paths, levels = assert_same_tree(library, single)
    assert len(nodes) == 1 and levels == 0 and nodes["child0"][0] == 0xFFFFFFFF and nodes["child1"][0] == tokens[0] and paths[0] == 0


def test_chain_deeper_than_a_path_is_refused(library):
    """LightTree.cs:29: 64 bits of path. Point lights on one line: every joint box has half area 0, every cut costs 0, the first one wins
    (`cost < minCost`) and the emitters peel off one at a time: depth = count - 1."""
    def chain(count):
        points = np.zeros(count, dtype=structs.POINT_LIGHT)
        points["position"][:, 0] = np.arange(count)
        points["intensity"] = (10.0 ** (0.5 * np.arange(count) - 18.0))[:, None]
        return describe(point_lights=points)

    *_, levels = assert_same_tree(library, chain(40))
    assert levels == 39
    *_, levels = assert_same_tree(library, chain(64))
    assert levels == 63  # the deepest tree a path can spell
    status, *_ = emulate(library, chain(65))
    assert status == 2
    with pytest.raises(ValueError):
        host.build_light_tree(chain(65))


def test_tree_invariants(library, lights_small):
    """structure of the emitted array: pre-order, tail subtree first, powers summed, boxes nested, paths spell the descent"""
    description = describe(lights_small.triangles, lights_small.spheres, lights_small.materials, lights_small.point_lights)
    _, nodes, tokens, paths, levels = emulate(library, description)
    leaf = nodes["child0"] == 0xFFFFFFFF
    assert np.count_nonzero(leaf) == len(tokens) == 300

    def leaves_below(index):
        return 1 if leaf[index] else leaves_below(nodes["child0"][index]) + leaves_below(nodes["child1"][index])

    depth_seen = 0
    for index in np.flatnonzero(~leaf):
        child0, child1 = int(nodes["child0"][index]), int(nodes["child1"][index])
        assert child0 == index + 1 and child1 == index + 2 * leaves_below(child0)
        assert nodes["power"][index] == np.float32(nodes["power"][child0]) + np.float32(nodes["power"][child1])
        for child in (child0, child1):
            assert np.all(nodes["boxMin"][index] <= nodes["boxMin"][child]) and np.all(nodes["boxMax"][index] >= nodes["boxMax"][child])

    for token, path in zip(tokens, paths):
        index, depth = 0, 0
        while not leaf[index]:
            index = int(nodes["child1"][index] if (int(path) >> depth) & 1 else nodes["child0"][index])
            depth += 1
        assert nodes["child1"][index] == token and int(path) >> depth == 0
        depth_seen = max(depth_seen, depth)
    assert depth_seen == levels


def test_pinned_transcendentals_are_the_runtime_functions_to_rounding(library):
    """MathF.Acos / MathF.Cos / Math.Acos belong to the C runtime in the reference; the pinned versions must be them up to the last bits"""
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-1, 1, 20000), 1 - 2.0 ** -rng.uniform(1, 50, 5000), -1 + 2.0 ** -rng.uniform(1, 50, 5000), [0.0, 1.0, -1.0, 0.5, -0.5, 1e-300, 2.0 ** -60]])
    mine = np.array([library.light_emulation_acos_double(float(v)) for v in x])
    assert np.max(np.abs(mine - np.arccos(x)) / np.maximum(np.arccos(x), 1e-300)) < 4.5e-16  # within two units of the last place of binary64
    assert np.array_equal(mine.astype(np.float32), np.arccos(x).astype(np.float32))        # Float3.Angle rounds it to binary32: the same float

    xf = np.concatenate([rng.uniform(-1, 1, 20000), [0.0, 1.0, -1.0, 0.5, -0.5]]).astype(np.float32)
    acos32 = np.array([library.light_emulation_acos(float(v)) for v in xf], dtype=np.float32)
    assert np.max(np.abs(acos32.astype(np.float64) - np.arccos(xf.astype(np.float64)))) < 4e-7
    angles = rng.uniform(-2 * np.pi, np.pi, 20000).astype(np.float32)  # ConeBound.RelativeArea's `offset - angle` and Union's `offset`
    cos32 = np.array([library.light_emulation_cos(float(v)) for v in angles], dtype=np.float32)
    assert np.max(np.abs(cos32.astype(np.float64) - np.cos(angles.astype(np.float64)))) < 2e-7


def test_whole_sphere_shortcut_of_the_cone_union_changes_nothing(library):
    """cone_union skips Float3.Angle when value0 already is the whole sphere (offset0 == pi): ConeBound.Union as written gives the same bits —
    for random cones, for cones at and next to cosOffset = -1, for parallel, opposite and zero axes (whose angle is 0 or NaN-free by the guards)."""
    library.light_emulation_cone_unions.argtypes = [ctypes.c_void_p] * 3
    library.light_emulation_acos_reaches_pi.restype = ctypes.c_uint32
    assert library.light_emulation_acos_reaches_pi() == 0x80000000  # the arc cosine is pi at -1 (and below, clamped) and nowhere else
    rng = np.random.default_rng(17)
    count = 20000
    cones = np.zeros((count, 10), dtype=np.float32)
    for base in (0, 5):
        axis = rng.normal(size=(count, 3))
        cones[:, base:base + 3] = axis / np.linalg.norm(axis, axis=1, keepdims=True)
        cones[:, base + 3] = rng.uniform(-1, 1, count)
        cones[:, base + 4] = rng.uniform(0, 1, count)
    cones[::3, 3] = -1.0                                             # value0 is the whole sphere
    cones[1::9, 3] = np.nextafter(np.float32(-1), np.float32(0))     # ... and just not
    cones[2::27, 3] = np.nextafter(np.float32(-1), np.float32(-2))   # below -1: the clamp makes it the whole sphere, the cosine is kept as it is
    cones[::7, 5:8] = cones[::7, 0:3]                                # parallel axes
    cones[::11, 5:8] = -cones[::11, 0:3]                             # opposite axes
    cones[::13, 5:8] = 0.0                                           # a zero axis (Float3.Angle returns 0)
    swap = cones[:, 8] < cones[:, 3]                                 # ConeBound.Encapsulate hands Union the wider cone (the lower cosOffset) first
    cones[swap] = np.concatenate([cones[swap, 5:], cones[swap, :5]], axis=1)
    a, b = np.zeros(5, dtype=np.float32), np.zeros(5, dtype=np.float32)
    saturated = 0
    for row in cones:
        library.light_emulation_cone_unions(row.ctypes.data, a.ctypes.data, b.ctypes.data)
        assert a.tobytes() == b.tobytes(), row
        saturated += row[3] == -1.0
    assert saturated > count // 4


def test_prepare_takes_a_light_tree_builder_for_every_pack(library):
    """host.prepare(light_tree_builder=...) — the hook the device build plugs into — hands every pack's emitters and its placements'
    PreparedInstance.LightBound rows to the builder: with the emulated level-synchronous build the prepared scene is the same scene."""
    calls = []

    def builder(description, instance_lights=None):
        status, nodes, tokens, paths, _ = emulate(library, description, instance_lights)
        assert status == 0
        calls.append(0 if instance_lights is None else len(instance_lights))
        return nodes.copy(), tokens.copy(), paths.copy(), float(nodes["power"][0]) if len(nodes) else 0.0

    expected = host.prepare(scenes.instanced_scene(grid=3, rings=8, segments=10))
    built = host.prepare(scenes.instanced_scene(grid=3, rings=8, segments=10), light_tree_builder=builder)
    assert len(calls) >= 2 and max(calls) > 0  # several packs, at least one with placements
    assert built.light_nodes.tobytes() == expected.light_nodes.tobytes()
    assert built.emitter_tokens.tobytes() == expected.emitter_tokens.tobytes() and built.emitter_bitpaths.tobytes() == expected.emitter_bitpaths.tobytes()
    assert built.infinite_threshold == expected.infinite_threshold and built.packs.tobytes() == expected.packs.tobytes()

    plain = host.prepare(scenes.many_lights_scene(light_count=64, rings=6, segments=6), light_tree_builder=builder)  # a scene without packs takes the hook too
    assert plain.light_nodes.tobytes() == host.prepare(scenes.many_lights_scene(light_count=64, rings=6, segments=6)).light_nodes.tobytes()


def test_cone_union_by_hand(library):
    """ConeBound.Union (ConeBound.cs:76-101) on cases worked out by hand from the C#, quirk included: `value0.axis.Angle(value1.axis)` is in
    DEGREES (Float3.cs:277-288) and is added to an offset in radians, and the rotation handed to `new Versor(cross, rotation)` is read as degrees.
      A  two direction cones (offset 0) whose axes are 4 degrees apart: max = 4 + 0; Min(4, pi) = pi > 0, no early exit; offset = (0 + 4) / 2 = 2
         < pi: a cone of 2 RADIANS half-angle, cosOffset = cos(2) = -0.41614684, its axis value0's turned by 2 DEGREES towards value1's
      B  the same, 8 degrees apart: offset = 4 >= pi: the whole sphere (axis Float3.Up, cosOffset -1)
      C  value0 with offset 1 rad (cosOffset cos 1), value1 a direction cone 0.5 degrees off: max = 0.5 <= 1: value0 unchanged
      D  the same axes: Angle = 0, max = offset1 = 0 <= offset0: value0 unchanged, whatever its offset
    cosExtend is the smaller of the two in every case."""
    library.light_emulation_cone_unions.argtypes = [ctypes.c_void_p] * 3

    def union(axis0, cos_offset0, cos_extend0, axis1, cos_offset1, cos_extend1):
        row = np.array([*axis0, cos_offset0, cos_extend0, *axis1, cos_offset1, cos_extend1], dtype=np.float32)
        a, b = np.zeros(5, dtype=np.float32), np.zeros(5, dtype=np.float32)
        library.light_emulation_cone_unions(row.ctypes.data, a.ctypes.data, b.ctypes.data)
        assert a.tobytes() == b.tobytes()
        return a

    def tilted(degrees):  # +Y turned towards +X
        return (math.sin(math.radians(degrees)), math.cos(math.radians(degrees)), 0.0)

    up = (0.0, 1.0, 0.0)
    a = union(up, 1.0, 0.25, tilted(4.0), 1.0, 0.5)
    assert a[3] == pytest.approx(math.cos(2.0), abs=1e-6) and a[4] == 0.25
    assert np.allclose(a[:3], tilted(2.0), atol=1e-6)

    b = union(up, 1.0, 0.25, tilted(8.0), 1.0, 0.5)
    assert b.tolist() == [0.0, 1.0, 0.0, -1.0, 0.25]

    c = union(up, math.cos(1.0), 0.75, tilted(0.5), 1.0, 0.5)
    assert c[:3].tolist() == [0.0, 1.0, 0.0] and c[3] == np.float32(math.cos(1.0)) and c[4] == 0.5

    for cos_offset in (1.0, 0.3, -0.9, -1.0):
        d = union(up, cos_offset, 0.1, up, 1.0, 0.2)
        assert d[:3].tolist() == [0.0, 1.0, 0.0] and d[3] == np.float32(cos_offset) and d[4] == np.float32(0.1)


def test_many_small_random_emitter_sets(library):
    """Two hundred small sets, built to make ties: few distinct positions (equal centres on every axis), repeated emitters, point lights among area
    lights (zero-size boxes: zero costs), emitters facing all ways or all one way, powers from 1e-3 to 1e3 — the level-synchronous passes must follow
    the recursion through every one of them."""
    rng = np.random.default_rng(2026)
    materials = np.concatenate([scenes.material(structs.MATERIAL_EMISSIVE, tuple(10.0 ** rng.uniform(-3, 3, 3))) for _ in range(6)])
    for trial in range(200):
        triangle_count, sphere_count, point_count = rng.integers(0, 14), rng.integers(0, 6), rng.integers(0, 6)
        if triangle_count + sphere_count + point_count < 2:
            triangle_count = 2
        lattice = rng.integers(1, 4)  # positions snapped to a lattice this coarse: 1 = everything in one spot
        snap = lambda count: np.round(rng.uniform(-2, 2, (count, 3)) * lattice) / lattice
        v0 = snap(triangle_count)
        if trial % 3 == 0:  # all facing one way: proper cones all the way up
            e1, e2 = np.tile((0.3, 0.0, 0.0), (triangle_count, 1)), np.tile((0.0, 0.0, 0.3), (triangle_count, 1))
        else:
            e1, e2 = rng.normal(0, 0.3, (triangle_count, 3)), rng.normal(0, 0.3, (triangle_count, 3))
        triangles = scenes.make_triangles(v0, v0 + e1, v0 + e2, rng.integers(0, 6, triangle_count).astype(np.uint32))
        spheres = np.zeros(sphere_count, dtype=structs.SPHERE)
        spheres["position"], spheres["radius"], spheres["material"] = snap(sphere_count), rng.choice([0.1, 0.25], sphere_count), rng.integers(0, 6, sphere_count)
        points = np.zeros(point_count, dtype=structs.POINT_LIGHT)
        points["position"], points["intensity"] = snap(point_count), 10.0 ** rng.uniform(-2, 2, (point_count, 3))
        assert_same_tree(library, describe(triangles, spheres, materials, points), reverse=bool(trial & 1))
