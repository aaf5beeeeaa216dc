"""The level-synchronous SweepBuilder (echorenderer_b200/csrc/echo_sweep.h: the passes and the driver of the device build, sweep.cu)
on a sequential CPU backend (tests/c_client/sweep_emulation.cpp): the QBVH it emits must be the host mirror's — SweepBuilder.cs +
the QuadBoundingVolumeHierarchy collapse, recursive, one node at a time — byte for byte: same nodes, same pre-order, same depth.
The GPU suite (test_gpu_build.py) asks the same of the CUDA backend."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from echorenderer_b200 import host, scenes, structs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emulation(tmp_path_factory):
    library = tmp_path_factory.mktemp("sweep") / "libsweep_emulation.so"
    subprocess.run(["g++", "-std=c++17", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC",
                    os.path.join(ROOT, "tests", "c_client", "sweep_emulation.cpp"), "-o", str(library)], check=True)
    lib = ctypes.CDLL(str(library))
    p, u32 = ctypes.c_void_p, ctypes.c_uint32
    lib.sweep_emulation_build.argtypes = [p, u32, p, u32, p, u32, ctypes.c_int32, p, ctypes.POINTER(u32), ctypes.POINTER(u32), ctypes.POINTER(u32)]
    lib.sweep_emulation_build.restype = ctypes.c_int32

    def build(triangles, spheres, reverse=False, instance_bounds=None):
        triangles = np.ascontiguousarray(triangles, dtype=structs.TRIANGLE)
        spheres = np.ascontiguousarray(spheres, dtype=structs.SPHERE)
        boxes = np.zeros((0, 6), dtype=np.float32) if instance_bounds is None else np.ascontiguousarray(instance_bounds, dtype=np.float32).reshape(-1, 6)
        nodes = np.zeros(max(len(triangles) + len(spheres) + len(boxes) - 1, 1), dtype=structs.QBVH_NODE)
        count, depth, levels = u32(), u32(), u32()
        status = lib.sweep_emulation_build(triangles.ctypes.data, len(triangles), spheres.ctypes.data, len(spheres), boxes.ctypes.data, len(boxes), int(reverse), nodes.ctypes.data,
                                           ctypes.byref(count), ctypes.byref(depth), ctypes.byref(levels))
        return status, nodes[:count.value], depth.value, levels.value

    return build


def assert_same_tree(emulation, triangles, spheres, reverse=False, instance_bounds=None):
    expected, expected_depth = host.build_qbvh(triangles, spheres, instance_bounds=instance_bounds)
    status, nodes, depth, levels = emulation(triangles, spheres, reverse, instance_bounds)
    assert status == 0
    assert len(nodes) == len(expected) and depth == expected_depth
    assert nodes.tobytes() == expected.tobytes()
    return levels


NO_SPHERES = np.zeros(0, dtype=structs.SPHERE)


@pytest.mark.parametrize("fixture", ["cornell", "terrain_small", "mixed_small", "lights_small"])
def test_emitted_tree_is_the_host_mirrors_byte_for_byte(emulation, fixture, request):
    prepared = request.getfixturevalue(fixture)
    levels = assert_same_tree(emulation, prepared.triangles, prepared.spheres)
    assert levels >= 2
    assert_same_tree(emulation, prepared.triangles, prepared.spheres, reverse=True)  # no pass depends on the order of its indices


def random_soup(seed, triangle_count, sphere_count, scale):
    rng = np.random.default_rng(seed)
    v0 = rng.uniform(-scale, scale, (triangle_count, 3))
    e1, e2 = rng.normal(0, scale * 0.02, (triangle_count, 3)), rng.normal(0, scale * 0.02, (triangle_count, 3))
    triangles = scenes.make_triangles(v0, v0 + e1, v0 + e2, 0)
    spheres = np.zeros(sphere_count, dtype=structs.SPHERE)
    spheres["position"] = rng.uniform(-scale, scale, (sphere_count, 3))
    spheres["radius"] = rng.uniform(0.01, 0.05, sphere_count) * scale
    return triangles, spheres


@pytest.mark.parametrize("seed,triangle_count,sphere_count,scale", [(1, 2, 0, 1.0), (2, 1, 1, 1.0), (3, 3, 0, 5.0), (4, 33, 7, 1.0), (5, 1000, 100, 100.0),
                                                                   (6, 20000, 500, 1e-3), (7, 5000, 5000, 1e4), (8, 0, 300, 2.0)])
def test_random_soups(emulation, seed, triangle_count, sphere_count, scale):
    triangles, spheres = random_soup(seed, triangle_count, sphere_count, scale)
    assert_same_tree(emulation, triangles, spheres)


def random_instance_bounds(seed, count, scale):
    rng = np.random.default_rng(seed)
    low = rng.uniform(-scale, scale, (count, 3))
    return np.concatenate([low, low + rng.uniform(0.0, 0.3 * scale, (count, 3))], axis=1).astype(np.float32)


def test_packs_with_placements(emulation):
    """GeometryCollection.CreateBounds: triangles, spheres, then the boxes of the pack's instances (TokenType.Instance leaves)."""
    triangles, spheres = random_soup(21, 400, 30, 10.0)
    assert_same_tree(emulation, triangles, spheres, instance_bounds=random_instance_bounds(22, 200, 10.0))
    assert_same_tree(emulation, triangles[:0], spheres[:0], instance_bounds=random_instance_bounds(23, 2304, 50.0))  # a pack of placements only
    prepared = host.prepare(scenes.instanced_scene(grid=4, rings=8, segments=8))
    assert np.any(structs.token_type(prepared.nodes["token4"].reshape(-1)) == structs.TOKEN_TYPE_INSTANCE)


def test_ties_and_degenerate_boxes(emulation):
    """Equal sort keys (the stable order decides), equal costs (the first cut wins), flat and point-sized boxes, axes that never change."""
    quad = scenes.plane(0, (2, 2))
    assert_same_tree(emulation, np.repeat(quad[:1], 50), NO_SPHERES)                       # fifty copies of one triangle
    assert_same_tree(emulation, np.concatenate([quad] * 40), NO_SPHERES)                   # forty copies of a two-triangle plane
    grid = scenes.terrain_triangles(24, 24, height=0.0)                                    # a flat, perfectly regular grid: ties everywhere
    assert_same_tree(emulation, grid, NO_SPHERES)
    row = np.concatenate([scenes.plane(0, (1, 1), position=(2.0 * k, 0, 0)) for k in range(70)])  # one axis only: nothing is ever re-sorted
    levels = assert_same_tree(emulation, row, NO_SPHERES)
    assert levels >= 7
    points = np.zeros(64, dtype=structs.SPHERE)                                             # zero-radius spheres on a lattice: zero-area boxes
    points["position"] = np.stack(np.meshgrid(np.arange(4.0), np.arange(4.0), np.arange(4.0)), axis=-1).reshape(-1, 3)
    assert_same_tree(emulation, quad[:0], points)


def test_a_chain_deeper_than_the_level_cap_gives_up(emulation):
    """Thousands of coincident primitives: every cost is equal, the first cut peels one primitive per level. The recursive reference
    would recurse as deep; the level-synchronous build stops at its cap and the caller falls back to another builder."""
    same = np.repeat(scenes.plane(0, (2, 2))[:1], 4000)
    status, _, _, _ = emulation(same, NO_SPHERES)
    assert status == 2
