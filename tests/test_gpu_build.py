"""The optional device-side tree build (echo_b200_build_qbvh). The default (sweep.cu) builds the reference's
SweepBuilder tree itself, level by level: the emitted nodes must equal the host mirror's byte for byte. The two other algorithms —
parallel locally-ordered clustering and a linear BVH, collapsed to the reference's QBVH node format — give other valid trees. All three
are checked for structural validity, against brute force through the oracle, and for device / oracle parity on the tree they produce."""
import time

import numpy as np
import pytest

from echorenderer_b200 import PreparedScene, _native, build_qbvh_device, host, scenes, structs
from tests import oracle_lib
from tests.test_gpu_trace import assert_hits_equal

pytestmark = pytest.mark.gpu

ALGORITHMS = {"sweep": 2, "ploc": 1, "lbvh": 0}
DEFAULT_ALGORITHM = 2


@pytest.fixture(params=list(ALGORITHMS))
def algorithm(request):
    _native.set_option("BUILD_ALGORITHM", ALGORITHMS[request.param])
    yield request.param
    _native.set_option("BUILD_ALGORITHM", DEFAULT_ALGORITHM)


def check_tree(nodes, max_depth, triangles, spheres):
    """Every primitive in exactly one slot, every slot box encloses its subtree, children ordered by their lower bound on the
    node's axes, [leaf, empty] pairs marked with axis 3, depth as CreateNode counts it."""
    seen = []
    depth = {}

    def box_of(node, k):
        return np.array([node["minX"][k], node["minY"][k], node["minZ"][k]]), np.array([node["maxX"][k], node["maxY"][k], node["maxZ"][k]])

    def primitive_box(token):
        index = structs.token_index(token)
        if structs.token_type(token) == structs.TOKEN_TYPE_TRIANGLE:
            t = triangles[index]
            points = np.stack([t["vertex0"], t["vertex0"] + t["edge1"], t["vertex0"] + t["edge2"]])
            return points.min(axis=0), points.max(axis=0)
        s = spheres[index]
        return s["position"] - s["radius"], s["position"] + s["radius"]

    def visit(index):
        node = nodes[index]
        low, high = np.full(3, np.inf), np.full(3, -np.inf)
        deepest = 0
        for pair, minor in ((0, int(node["axisMinor0"])), (1, int(node["axisMinor1"]))):
            a, b = int(node["token4"][pair * 2]), int(node["token4"][pair * 2 + 1])
            assert a != structs.TOKEN_EMPTY
            if minor == 3:
                assert b == structs.TOKEN_EMPTY and structs.token_type(a) != structs.TOKEN_TYPE_NODE
            else:
                assert b != structs.TOKEN_EMPTY and box_of(node, pair * 2)[0][minor] <= box_of(node, pair * 2 + 1)[0][minor]
        first, second = box_of(node, 0), box_of(node, 2)
        assert first[0][int(node["axisMajor"])] <= max(second[0][int(node["axisMajor"])], box_of(node, 1)[0][int(node["axisMajor"])] if node["token4"][1] != structs.TOKEN_EMPTY else -np.inf) \
            or np.minimum(first[0], box_of(node, 1)[0] if node["token4"][1] != structs.TOKEN_EMPTY else first[0])[int(node["axisMajor"])] <= second[0][int(node["axisMajor"])]
        for k in range(4):
            token = int(node["token4"][k])
            if token == structs.TOKEN_EMPTY:
                assert np.all(np.isposinf(box_of(node, k)[0])) and np.all(np.isposinf(box_of(node, k)[1]))
                continue
            slot_low, slot_high = box_of(node, k)
            if structs.token_type(token) == structs.TOKEN_TYPE_NODE:
                child_low, child_high, child_depth = visit(structs.token_index(token))
            else:
                seen.append(token)
                child_low, child_high = primitive_box(token)
                child_depth = 1
            assert np.all(slot_low <= child_low + 1e-6 * np.abs(child_low)) and np.all(slot_high >= child_high - 1e-6 * np.abs(child_high))
            low, high = np.minimum(low, slot_low), np.maximum(high, slot_high)
            deepest = max(deepest, child_depth)
        depth[index] = deepest + 1
        return low, high, deepest + 1

    import sys
    sys.setrecursionlimit(10000)
    visit(0)
    assert depth[0] == max_depth
    assert len(seen) == len(triangles) + len(spheres) and len(set(seen)) == len(seen)


@pytest.mark.parametrize("fixture", ["cornell", "terrain_small", "mixed_small"])
def test_device_tree_is_valid_and_traces_like_brute_force(fixture, request, algorithm):
    sah = request.getfixturevalue(fixture)
    nodes, depth = build_qbvh_device(sah.triangles, sah.spheres)
    assert len(nodes) <= len(sah.triangles) + len(sah.spheres) - 1
    check_tree(nodes, depth, sah.triangles, sah.spheres)

    prepared = host.prepare(sah.description, tree=(nodes, depth))
    oracle = oracle_lib.OracleScene(prepared)
    rays = scenes.random_rays(prepared.bounds, 50_000, seed=31)

    expected = oracle.trace(rays)
    reference = oracle_lib.OracleScene(sah).trace(rays)  # the SweepBuilder mirror's tree: same hits up to the order of exact ties
    same = expected["token"] == reference["token"]
    # coplanar overlapping faces (the Cornell boxes stand ON the floor) are ties to the last bit: which one wins depends on the
    # order of discovery and on whether the other one's node is culled by `entry >= distance` — i.e. on the tree
    assert same.mean() > 0.99
    hit = np.isfinite(reference["distance"])
    assert np.array_equal(hit, np.isfinite(expected["distance"]))
    assert np.allclose(expected["distance"][hit], reference["distance"][hit], rtol=2e-6)

    with PreparedScene(prepared) as scene:
        assert_hits_equal(scene.trace(rays), expected)
        shadow = scenes.random_rays(prepared.bounds, 50_000, seed=33, occlusion=True)
        assert np.array_equal(scene.occlude(shadow), oracle.occlude(shadow))


def test_tiny_and_degenerate_inputs(algorithm):
    two = scenes.plane(0, (2, 2))
    nodes, depth = build_qbvh_device(two, np.zeros(0, dtype=structs.SPHERE))
    assert len(nodes) == 1 and depth == 2
    check_tree(nodes, depth, two, np.zeros(0, dtype=structs.SPHERE))

    # many coincident primitives: equal Morton codes fall back to the sorted positions (Karras 2012, section 4)
    same = np.repeat(scenes.plane(0, (2, 2))[:1], 200)
    nodes, depth = build_qbvh_device(same, np.zeros(0, dtype=structs.SPHERE))
    check_tree(nodes, depth, same, np.zeros(0, dtype=structs.SPHERE))

    # thousands of coincident primitives: the clustering merges one pair per pass, stalls, and the build falls back to the Morton tree
    many = np.repeat(scenes.plane(0, (2, 2))[:1], 3000)
    nodes, depth = build_qbvh_device(many, np.zeros(0, dtype=structs.SPHERE))
    check_tree(nodes, depth, many, np.zeros(0, dtype=structs.SPHERE))

    from echorenderer_b200 import EchoNativeError
    with pytest.raises(EchoNativeError):
        build_qbvh_device(two[:1], np.zeros(0, dtype=structs.SPHERE))


def test_clustered_tree_is_cheaper_to_traverse_than_the_morton_tree(terrain_small):
    """PLOC's surface-area criterion buys tree quality: fewer node visits per query than the linear BVH on the same rays (oracle
    visit counters), within reach of the full-sweep SAH tree of the host mirror."""
    rays = scenes.random_rays(terrain_small.bounds, 100_000, seed=37)
    visits = {}
    for name, code in ALGORITHMS.items():
        _native.set_option("BUILD_ALGORITHM", code)
        nodes, depth = build_qbvh_device(terrain_small.triangles, terrain_small.spheres)
        prepared = host.prepare(terrain_small.description, tree=(nodes, depth))
        _, counters = oracle_lib.OracleScene(prepared).trace(rays, count_visits=True)
        visits[name] = counters[0] / len(rays)
    _native.set_option("BUILD_ALGORITHM", DEFAULT_ALGORITHM)
    _, counters = oracle_lib.OracleScene(terrain_small).trace(rays, count_visits=True)
    sah = counters[0] / len(rays)
    assert visits["ploc"] < visits["lbvh"]
    assert visits["ploc"] < 1.25 * sah
    assert visits["sweep"] == sah  # the same tree


def assert_device_tree_is_the_host_mirrors(triangles, spheres, instance_bounds=None):
    expected, expected_depth = host.build_qbvh(triangles, spheres, instance_bounds=instance_bounds)
    nodes, depth = build_qbvh_device(triangles, spheres, instance_bounds=instance_bounds)
    assert len(nodes) == len(expected) and depth == expected_depth
    assert nodes.tobytes() == expected.tobytes()


@pytest.mark.parametrize("fixture", ["cornell", "terrain_small", "mixed_small", "lights_small"])
def test_device_sweep_tree_is_the_host_mirrors_byte_for_byte(fixture, request):
    """BUILD_ALGORITHM 2 (the default): SweepBuilder.cs + the QuadBoundingVolumeHierarchy collapse, level-synchronous on the device —
    the same nodes in the same pre-order as the recursive build (tests/test_sweep_build.py asks the same of the CPU emulation)."""
    _native.set_option("BUILD_ALGORITHM", 2)
    prepared = request.getfixturevalue(fixture)
    assert_device_tree_is_the_host_mirrors(prepared.triangles, prepared.spheres)


def test_device_sweep_tree_on_soups_ties_and_tiny_inputs():
    from tests.test_sweep_build import NO_SPHERES, random_instance_bounds, random_soup
    _native.set_option("BUILD_ALGORITHM", 2)
    triangles, spheres = random_soup(21, 400, 30, 10.0)  # packs with placements: TokenType.Instance leaves after the spheres
    assert_device_tree_is_the_host_mirrors(triangles, spheres, random_instance_bounds(22, 200, 10.0))
    assert_device_tree_is_the_host_mirrors(triangles[:0], spheres[:0], random_instance_bounds(23, 2304, 50.0))
    for seed, triangle_count, sphere_count, scale in [(1, 2, 0, 1.0), (2, 1, 1, 1.0), (3, 3, 0, 5.0), (4, 33, 7, 1.0), (5, 1000, 100, 100.0), (6, 20000, 500, 1e-3),
                                                       (7, 5000, 5000, 1e4), (8, 0, 300, 2.0), (9, 300_000, 3000, 10.0)]:
        assert_device_tree_is_the_host_mirrors(*random_soup(seed, triangle_count, sphere_count, scale))
    quad = scenes.plane(0, (2, 2))
    assert_device_tree_is_the_host_mirrors(np.concatenate([quad] * 40), NO_SPHERES)                # equal keys: the stable order decides
    assert_device_tree_is_the_host_mirrors(scenes.terrain_triangles(24, 24, height=0.0), NO_SPHERES)  # a regular flat grid: equal costs everywhere
    assert_device_tree_is_the_host_mirrors(np.concatenate([scenes.plane(0, (1, 1), position=(2.0 * k, 0, 0)) for k in range(70)]), NO_SPHERES)


def test_device_sweep_tree_at_full_size():
    """C2's geometry (1 000 000 triangles + 10 000 spheres): the device-built tree is the host mirror's 565 329 nodes, byte for byte."""
    _native.set_option("BUILD_ALGORITHM", 2)
    description = scenes.terrain_scene()
    build_qbvh_device(description.triangles[:64], description.spheres[:0])  # context + module load
    started = time.perf_counter()
    nodes, depth = build_qbvh_device(description.triangles, description.spheres)
    device_seconds = time.perf_counter() - started
    started = time.perf_counter()
    expected, expected_depth = host.build_qbvh(description.triangles, description.spheres)
    host_seconds = time.perf_counter() - started
    print(f"sweep build of {len(description.triangles) + len(description.spheres)} primitives: device call {device_seconds * 1e3:.1f} ms, host mirror {host_seconds * 1e3:.1f} ms")
    assert len(nodes) == len(expected) == 565_329 and depth == expected_depth == 13
    assert nodes.tobytes() == expected.tobytes()



def test_instanced_scene_prepared_with_device_built_trees_is_the_same_scene():
    """Every pack of a nested instanced scene (placements of placements) built on the device: host.prepare with the device builder gives
    the very arrays the host mirror gives, so everything downstream — traversal, shading, light trees — is unchanged."""
    _native.set_option("BUILD_ALGORITHM", 2)
    expected = host.prepare(scenes.instanced_scene(grid=4, rings=12, segments=12))
    built = host.prepare(scenes.instanced_scene(grid=4, rings=12, segments=12), tree_builder=lambda t, s, b: build_qbvh_device(t, s, instance_bounds=b))
    assert built.max_depth == expected.max_depth and built.nodes.tobytes() == expected.nodes.tobytes()
    assert np.any(structs.token_type(built.nodes["token4"].reshape(-1)) == structs.TOKEN_TYPE_INSTANCE)


def test_scene_builds_its_own_accelerator(terrain_small, cornell):
    """echo_b200_scene_build_qbvh: a host that uploads only geometry gets the SweepBuilder's tree built on the device — the same node
    count and depth as the host mirror's, and hits, distances and barycentrics bit for bit those of the scene given the mirror's nodes."""
    _native.set_option("BUILD_ALGORITHM", 2)
    for prepared in (terrain_small, cornell):
        rays = scenes.random_rays(prepared.bounds, 50_000, seed=41)
        shadow = scenes.random_rays(prepared.bounds, 50_000, seed=43, occlusion=True)
        with PreparedScene(prepared) as given, PreparedScene(prepared, build_tree_on_device=True) as built:
            assert built.built_tree == (len(prepared.nodes), prepared.max_depth)
            assert_hits_equal(built.trace(rays), given.trace(rays))
            assert np.array_equal(built.occlude(shadow), given.occlude(shadow))
