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
