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


# ---- the light tree (echo_b200_build_light_tree, csrc/lightbuild.cu): LightTree.Build level by level on the device ----

def assert_device_light_tree_is_the_host_mirrors(description, instance_lights=None):
    from echorenderer_b200 import build_light_tree_device
    expected_nodes, expected_tokens, expected_paths, expected_power = host.build_light_tree(description, instance_lights)
    nodes, tokens, paths, power = build_light_tree_device(description, instance_lights)
    assert len(nodes) == len(expected_nodes) and len(tokens) == len(expected_tokens)
    assert nodes.tobytes() == expected_nodes.tobytes()
    assert tokens.tobytes() == expected_tokens.tobytes() and paths.tobytes() == expected_paths.tobytes()
    assert np.float32(power) == np.float32(expected_power)
    return nodes, tokens


@pytest.mark.parametrize("fixture", ["cornell", "lights_small", "mixed_small"])
def test_device_light_tree_is_the_host_mirrors_byte_for_byte(fixture, request):
    """The same nodes in the same pre-order, the same emitter tokens and bit paths as the recursive build (tests/test_light_build.py asks
    the same of the CPU emulation of these passes) — and therefore what the scene was prepared with."""
    from tests.test_light_build import describe
    prepared = request.getfixturevalue(fixture)
    nodes, tokens = assert_device_light_tree_is_the_host_mirrors(describe(prepared.triangles, prepared.spheres, prepared.materials, prepared.point_lights))
    assert nodes.tobytes() == prepared.light_nodes.tobytes() and tokens.tobytes() == prepared.emitter_tokens.tobytes()


def test_device_light_tree_on_random_emitters_ties_placements_and_tiny_inputs():
    from echorenderer_b200 import build_light_tree_device
    from tests.test_light_build import describe, random_emitters
    for seed, triangle_count, sphere_count, point_count, scale in [(1, 2, 0, 0, 1.0), (2, 0, 0, 2, 1.0), (3, 1, 1, 1, 5.0), (4, 40, 9, 3, 1.0), (5, 1500, 100, 20, 100.0),
                                                                   (6, 6000, 0, 0, 1e-2), (7, 500, 500, 500, 1e3), (8, 40_000, 2000, 100, 30.0)]:
        assert_device_light_tree_is_the_host_mirrors(random_emitters(seed, triangle_count, sphere_count, point_count, scale, emissive_share=1.0 if triangle_count + sphere_count < 5 else 0.7))

    rng = np.random.default_rng(11)
    rows = np.zeros((9, 12), dtype=np.float32)  # PreparedInstance.LightBound rows: the placements join the list last
    low = rng.uniform(-10, 10, (9, 3))
    rows[:, 0:3], rows[:, 3:6] = low, low + rng.uniform(0.5, 3.0, (9, 3))
    axis = rng.normal(size=(9, 3))
    rows[:, 6:9] = axis / np.linalg.norm(axis, axis=1, keepdims=True)
    rows[:, 9], rows[:, 10], rows[:, 11] = rng.uniform(-1, 1, 9), rng.uniform(0, 1, 9), rng.uniform(0.5, 80.0, 9)
    rows[4, 11] = 0.0
    assert_device_light_tree_is_the_host_mirrors(random_emitters(12, 60, 10, 4, 10.0), rows)

    ix, iy = np.meshgrid(np.arange(-4, 4), np.arange(-2, 2), indexing="ij")  # a lattice of identical emitters, twice: every sort meets ties
    v0 = np.stack([ix.reshape(-1), np.zeros(ix.size), iy.reshape(-1)], axis=-1).astype(np.float64)
    v0 = np.concatenate([v0, v0, -v0 * 0.0])
    lattice = describe(scenes.make_triangles(v0 - (0.25, 0, 0.25), v0 + (0.25, 0, -0.25), v0 + (-0.25, 0, 0.25), 0), materials=np.concatenate([scenes.material(structs.MATERIAL_EMISSIVE, (3.0, 2.0, 1.0))]))
    assert_device_light_tree_is_the_host_mirrors(lattice)

    from tests.test_independent_kats import _tilted_ceiling  # emitters facing almost the same way: proper cones, every step runs the whole union
    nodes, _ = assert_device_light_tree_is_the_host_mirrors(_tilted_ceiling(scenes.many_lights_scene(light_count=3000, rings=8, segments=8), 1.5))
    assert np.count_nonzero((nodes["child0"] != 0xFFFFFFFF) & (nodes["cosOffset"] > -1) & (nodes["cosOffset"] < 1)) > 2000

    dark = random_emitters(3, 50, 5, 0, 1.0, emissive_share=0.0)  # no emitter: an empty tree
    nodes, tokens, paths, power = build_light_tree_device(dark)
    assert len(nodes) == 0 and len(tokens) == 0 and len(paths) == 0 and power == 0.0
    single = describe(point_lights=np.array([((1.0, 2.0, 3.0), (0.5, -1.0, 4.0))], dtype=structs.POINT_LIGHT))
    nodes, _ = assert_device_light_tree_is_the_host_mirrors(single)  # one emitter: the root is its leaf
    assert len(nodes) == 1

    def chain(count):  # point lights on a line: every cut costs 0, the first wins, depth = count - 1 (LightTree.cs:29: at most 63)
        points = np.zeros(count, dtype=structs.POINT_LIGHT)
        points["position"][:, 0] = np.arange(count)
        points["intensity"] = (10.0 ** (0.5 * np.arange(count) - 18.0))[:, None]
        return describe(point_lights=points)

    assert_device_light_tree_is_the_host_mirrors(chain(64))
    with pytest.raises(_native.EchoNativeError) as refused:
        build_light_tree_device(chain(65))
    assert refused.value.status == _native.ERR_UNSUPPORTED


def test_device_light_tree_at_full_size():
    """C4's emitters (10 000 emissive triangles among 309 502): the device-built light tree is the host mirror's 19 999 nodes, byte for byte."""
    from echorenderer_b200 import build_light_tree_device
    from tests.test_light_build import describe
    description = scenes.many_lights_scene()
    build_light_tree_device(describe(description.triangles[:64], materials=description.materials))  # context + module load
    started = time.perf_counter()
    nodes, tokens, paths, power = build_light_tree_device(description)
    device_seconds = time.perf_counter() - started
    phases = _native.last_light_build()
    started = time.perf_counter()
    expected_nodes, expected_tokens, expected_paths, expected_power = host.build_light_tree(description)
    host_seconds = time.perf_counter() - started
    print(f"light tree over {len(tokens)} emitters of {len(description.triangles)} triangles: device call {device_seconds * 1e3:.1f} ms {phases}, host mirror {host_seconds * 1e3:.1f} ms")
    assert len(nodes) == len(expected_nodes) == 19_999 and len(tokens) == 10_000
    assert nodes.tobytes() == expected_nodes.tobytes() and tokens.tobytes() == expected_tokens.tobytes() and paths.tobytes() == expected_paths.tobytes()
    assert np.float32(power) == np.float32(expected_power)


def test_scene_builds_its_own_light_tree(lights_small, cornell):
    """echo_b200_scene_build_light_tree: a host that uploads geometry and materials gets the reference's light tree built on the device —
    and every rendered tile is, bit for bit, the tile of the scene given the host mirror's tree and emitter map."""
    for prepared, size in ((lights_small, 64), (cornell, 32)):
        params = structs.render_params(size, size, 16, extend=4, seed=5, bounce_limit=16)
        tiles = scenes.tile_grid(size, size, 16)
        with PreparedScene(prepared) as given, PreparedScene(prepared, build_light_tree_on_device=True) as built:
            count, emitters, power = built.built_light_tree
            assert count == len(prepared.light_nodes) and emitters == len(prepared.emitter_tokens) and np.float32(power) == prepared.light_nodes["power"][0]
            expected, expected_stats = given.render_tiles(params, tiles)
            actual, stats = built.render_tiles(params, tiles)
            assert actual.tobytes() == expected.tobytes() and stats.tobytes() == expected_stats.tobytes()

    with pytest.raises(_native.EchoNativeError):  # scenes with packs build each pack's tree with echo_b200_build_light_tree
        PreparedScene(host.prepare(scenes.instanced_scene(grid=2, rings=6, segments=6)), build_light_tree_on_device=True)


def test_instanced_scene_prepared_with_device_built_light_trees_is_the_same_scene():
    """Every pack's light tree (placements included: PreparedInstance.LightBound rows) built on the device gives the arrays the host gives."""
    from echorenderer_b200 import build_light_tree_device
    expected = host.prepare(scenes.instanced_scene(grid=4, rings=12, segments=12))
    built = host.prepare(scenes.instanced_scene(grid=4, rings=12, segments=12), light_tree_builder=build_light_tree_device)
    assert len(expected.light_nodes) > 1
    assert built.light_nodes.tobytes() == expected.light_nodes.tobytes()
    assert built.emitter_tokens.tobytes() == expected.emitter_tokens.tobytes() and built.emitter_bitpaths.tobytes() == expected.emitter_bitpaths.tobytes()
    assert built.infinite_threshold == expected.infinite_threshold
