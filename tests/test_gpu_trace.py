"""Parity of the CUDA closest-hit / occlusion path with the oracle, through the C ABI (BASELINE config C2 and friends).

Bar (BASELINE.json north_star): hit primitive IDs bit-exact, distances within 4 ulp. The device follows the reference's
traversal order exactly, so the tests demand bit-exact tokens, distances AND barycentrics, ties included."""
import numpy as np
import pytest

from echorenderer_b200 import PreparedScene, scenes, structs
from tests import oracle_lib

pytestmark = pytest.mark.gpu


def assert_hits_equal(actual, expected):
    assert np.array_equal(actual["token"], expected["token"])
    assert np.array_equal(actual["distance"].view(np.uint32), expected["distance"].view(np.uint32))
    hit = expected["token"] != structs.TOKEN_EMPTY
    assert np.array_equal(actual["uv"][hit].view(np.uint32), expected["uv"][hit].view(np.uint32))


@pytest.mark.parametrize("fixture", ["cornell", "terrain_small", "mixed_small", "lights_small"])
def test_trace_matches_oracle(fixture, request):
    prepared = request.getfixturevalue(fixture)
    oracle = oracle_lib.OracleScene(prepared)
    rays = scenes.random_rays(prepared.bounds, 200_000, seed=11)

    with PreparedScene(prepared) as scene:
        hits = scene.trace(rays)
        expected = oracle.trace(rays)
        assert (expected["token"] != structs.TOKEN_EMPTY).mean() > 0.05
        assert_hits_equal(hits, expected)

        # secondary rays: origins on surfaces, ignore = hit token (TraceQuery.SpawnTrace)
        secondary = scenes.secondary_rays(prepared, rays, expected)
        assert_hits_equal(scene.trace(secondary), oracle.trace(secondary))


@pytest.mark.parametrize("fixture", ["cornell", "terrain_small", "mixed_small"])
def test_occlude_matches_oracle(fixture, request):
    prepared = request.getfixturevalue(fixture)
    oracle = oracle_lib.OracleScene(prepared)
    rays = scenes.random_rays(prepared.bounds, 200_000, seed=11, occlusion=True)

    with PreparedScene(prepared) as scene:
        occluded = scene.occlude(rays)
        expected = oracle.occlude(rays)
        assert 0.02 < expected.mean() < 0.98
        assert np.array_equal(occluded, expected)

        primary = scenes.random_rays(prepared.bounds, 100_000, seed=13)
        secondary = scenes.secondary_rays(prepared, primary, oracle.trace(primary))
        secondary["distance"] = 5.0
        assert np.array_equal(scene.occlude(secondary), oracle.occlude(secondary))


def test_edge_cases(terrain_small):
    """Empty batch, ragged (non multiple of the block) sizes, zero/negative/tiny limits, axis-aligned directions (1/0 = inf)."""
    oracle = oracle_lib.OracleScene(terrain_small)

    with PreparedScene(terrain_small) as scene:
        assert len(scene.trace(np.zeros(0, dtype=structs.RAY))) == 0
        assert len(scene.occlude(np.zeros(0, dtype=structs.RAY))) == 0

        for count in (1, 31, 129, 1000):
            rays = scenes.random_rays(terrain_small.bounds, count, seed=3)
            assert_hits_equal(scene.trace(rays), oracle.trace(rays))

        rays = scenes.random_rays(terrain_small.bounds, 4096, seed=5)
        rays["distance"][0::4] = 0.0
        rays["distance"][1::4] = -1.0
        rays["distance"][2::4] = 7e-7  # below FastMath.Epsilon: PreparedScene.Trace rejects the query
        rays["distance"][3::4] = 3.0
        assert_hits_equal(scene.trace(rays), oracle.trace(rays))
        assert np.array_equal(scene.occlude(rays), oracle.occlude(rays))

        axis = scenes.random_rays(terrain_small.bounds, 6000, seed=9)
        directions = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float32)
        axis["direction"] = directions[np.arange(6000) % 6]
        axis["origin"][:, 1] += 10.0
        assert_hits_equal(scene.trace(axis), oracle.trace(axis))
        assert np.array_equal(scene.occlude(axis), oracle.occlude(axis))


def test_cornell_ties(cornell):
    """The Cornell box has coplanar faces (box bottoms on the floor): equal-distance hits must resolve like the reference."""
    oracle = oracle_lib.OracleScene(cornell)
    rays = scenes.random_rays(cornell.bounds, 300_000, seed=17)
    rays["origin"][:, 1] = np.abs(rays["origin"][:, 1]) + 0.5
    rays["direction"][:, 1] = -np.abs(rays["direction"][:, 1])

    with PreparedScene(cornell) as scene:
        assert_hits_equal(scene.trace(rays), oracle.trace(rays))


def chain_scene(depth):
    """A hand-built, maximally unbalanced QBVH: node i = [triangle i, empty, node i + 1, empty] (axisMinor = 3 marks the
    [leaf, empty] pairs, QuadBoundingVolumeHierarchy.cs:531-533). Quad depth = `depth`, stack size = 3 * depth + 1 (:34)."""
    from echorenderer_b200 import host as host_module
    count = depth  # triangles; the last node holds the last two
    x = np.arange(count, dtype=np.float64) * 2.0
    v0 = np.stack([x, np.zeros(count), -np.ones(count)], axis=-1)
    v1 = np.stack([x + 1.0, np.zeros(count), -np.ones(count)], axis=-1)
    v2 = np.stack([x + 0.5, np.ones(count), np.ones(count)], axis=-1)
    triangles = scenes.make_triangles(v0, v1, v2, 0)
    low = np.minimum(np.minimum(v0, v1), v2).astype(np.float32)
    high = np.maximum(np.maximum(v0, v1), v2).astype(np.float32)

    nodes = np.zeros(count - 1, dtype=structs.QBVH_NODE)
    for key in ("minX", "minY", "minZ", "maxX", "maxY", "maxZ"):
        nodes[key] = np.inf
    nodes["token4"] = structs.TOKEN_EMPTY
    nodes["axisMajor"], nodes["axisMinor0"], nodes["axisMinor1"] = 0, 3, 3

    for i in range(count - 1):
        last = i == count - 2
        rest_low, rest_high = low[i + 1:].min(axis=0), high[i + 1:].max(axis=0)
        for slot, (lo, hi, token) in ((0, (low[i], high[i], structs.make_token(1, i))),
                                      (2, (rest_low, rest_high, structs.make_token(1, i + 1) if last else structs.make_token(0, i + 1)))):
            nodes["minX"][i, slot], nodes["minY"][i, slot], nodes["minZ"][i, slot] = lo
            nodes["maxX"][i, slot], nodes["maxY"][i, slot], nodes["maxZ"][i, slot] = hi
            nodes["token4"][i, slot] = token

    description = host_module.SceneDescription(triangles=triangles, materials=scenes.material(structs.MATERIAL_DIFFUSE),
                                               camera=scenes.perspective_camera((0, 5, -5)))
    empty_u32, empty_u64 = np.zeros(0, np.uint32), np.zeros(0, np.uint64)
    return host_module.PreparedArrays(description, nodes, count, np.zeros(0, structs.LIGHT_NODE), empty_u32, empty_u64, 0.0, 0.0, 0.0)


@pytest.mark.parametrize("depth", [12, 20, 40])
def test_deep_trees_use_the_larger_stack_classes(depth):
    """maxDepth 12 / 20 / 40 -> 37 / 61 / 121 stack entries -> the 48 / 96 / 192-entry kernels."""
    prepared = chain_scene(depth)
    oracle = oracle_lib.OracleScene(prepared)
    rays = scenes.random_rays(prepared.bounds, 50_000, seed=23)
    # every other ray is aimed at the centroid of some triangle of the chain, so it crosses many of the nested boxes
    target = prepared.triangles[np.arange(25_000) % len(prepared.triangles)]
    centroid = target["vertex0"] + (target["edge1"] + target["edge2"]) / np.float32(3)
    aim = (centroid - rays["origin"][::2]).astype(np.float64)
    rays["direction"][::2] = (aim / np.linalg.norm(aim, axis=1, keepdims=True)).astype(np.float32)
    shadow = rays.copy()
    shadow["distance"] = 7.0

    with PreparedScene(prepared) as scene:
        expected = oracle.trace(rays)
        assert (expected["token"] != structs.TOKEN_EMPTY).mean() > 0.2
        assert_hits_equal(scene.trace(rays), expected)
        assert np.array_equal(scene.occlude(shadow), oracle.occlude(shadow))


def test_too_deep_tree_is_rejected():
    from echorenderer_b200 import EchoNativeError
    with pytest.raises(EchoNativeError) as error:
        PreparedScene(chain_scene(70))
    assert error.value.status == 4  # ECHO_B200_ERR_UNSUPPORTED


def test_spheres_only_and_tiny_scenes():
    """A scene of spheres only (no triangle array at all) and the smallest legal hierarchy (2 primitives)."""
    from echorenderer_b200 import host as host_module
    rng = np.random.default_rng(5)
    spheres = np.zeros(700, dtype=structs.SPHERE)
    spheres["position"] = rng.uniform(-10, 10, size=(700, 3)).astype(np.float32)
    spheres["radius"] = rng.uniform(0.1, 1.0, size=700).astype(np.float32)
    description = host_module.SceneDescription(spheres=spheres, materials=scenes.material(structs.MATERIAL_DIFFUSE), camera=scenes.perspective_camera((0, 0, -30)))
    prepared = host_module.prepare(description)
    oracle = oracle_lib.OracleScene(prepared)
    rays = scenes.random_rays(prepared.bounds, 100_000, seed=31)

    with PreparedScene(prepared) as scene:
        expected = oracle.trace(rays)
        assert_hits_equal(scene.trace(rays), expected)
        inside = scenes.secondary_rays(prepared, rays, expected)   # leave the sphere surfaces: findFar path (SphereEntity.cs:106-109)
        assert_hits_equal(scene.trace(inside), oracle.trace(inside))
        inside["distance"] = 3.0
        assert np.array_equal(scene.occlude(inside), oracle.occlude(inside))

    description = host_module.SceneDescription(triangles=scenes.plane(0, (2, 2)), materials=scenes.material(structs.MATERIAL_DIFFUSE), camera=scenes.perspective_camera((0, 3, -3)))
    prepared = host_module.prepare(description)
    assert len(prepared.nodes) == 1
    oracle = oracle_lib.OracleScene(prepared)
    rays = scenes.random_rays((np.array([-1, -1, -1.0]), np.array([1, 1, 1.0])), 20_000, seed=33)

    with PreparedScene(prepared) as scene:
        assert_hits_equal(scene.trace(rays), oracle.trace(rays))
