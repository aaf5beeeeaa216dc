"""Parity at BASELINE.json's FULL sizes (the other GPU tests run on small fixtures so the oracle finishes in seconds).

C2: the whole 1 000 000-triangle + 10 000-sphere scene and the whole 16 777 216-ray batch of `bench.py` — the oracle does the
33.5 M queries in a few seconds on the box's host cores, so tokens, distances and barycentrics are compared bit for bit on
every ray, followed by the size-independent properties of the domain (occlusion agrees with the closest hit on either side of
the hit distance; a hit never lies beyond the ray's limit; re-tracing is idempotent under the persistent scheduler).
C3 / C4: the full-resolution 1920 x 1080 frame of the full-size scenes, a spread of its tiles against the oracle.
C5: the 9 998 246-triangle scene (4.9 M nodes, quad depth 16) at 3840 x 2160: a 2 Mi-ray incoherent batch and a spread of tiles."""
import numpy as np
import pytest

from echorenderer_b200 import PreparedScene, hilbert_curve_pattern, host, scenes, structs
from tests import oracle_lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    return host.prepare(scenes.terrain_scene(1000, 500, 10000))


def test_c2_whole_batch_bit_exact(c2):
    oracle = oracle_lib.OracleScene(c2)
    count = 1 << 24
    rays = scenes.random_rays(c2.bounds, count, seed=11)
    low, high = c2.bounds
    diagonal = float(np.linalg.norm(np.asarray(high, np.float64) - np.asarray(low, np.float64)))
    shadow = rays.copy()
    shadow["distance"] = scenes.uniform(17, np.arange(count, dtype=np.uint64)) * np.float32(diagonal)

    with PreparedScene(c2) as scene:
        hits = scene.trace(rays)
        expected = oracle.trace(rays)
        assert np.array_equal(hits["token"], expected["token"])
        assert np.array_equal(hits["distance"].view(np.uint32), expected["distance"].view(np.uint32))
        hit = expected["token"] != structs.TOKEN_EMPTY
        assert 0.05 < hit.mean() < 0.95
        assert np.array_equal(hits["uv"][hit].view(np.uint32), expected["uv"][hit].view(np.uint32))
        kinds = structs.token_type(hits["token"][hit])
        assert (kinds == structs.TOKEN_TYPE_TRIANGLE).any() and (kinds == structs.TOKEN_TYPE_SPHERE).any()

        occluded = scene.occlude(shadow)
        assert np.array_equal(occluded, oracle.occlude(shadow))

        # occlusion vs closest hit: a ray is occluded within `travel` exactly when its closest hit is nearer than `travel`
        # (same intersection arithmetic on both paths; rays whose hit distance is within 1e-4 of travel are left out)
        distance = np.where(hit, hits["distance"], np.inf).astype(np.float64)
        travel = shadow["distance"].astype(np.float64)
        clear = np.abs(distance - travel) > 1e-4 * np.maximum(travel, 1.0)
        assert np.array_equal(occluded[clear].astype(bool), (distance < travel)[clear])

        # just short of / just beyond the hit: never / always occluded — up to grazing rays: the reference's occlusion test is
        # the division-free form of its closest-hit test (TriangleEntity.cs:236-258 `u > determinant` against :204-234
        # `u * determinantR > 1`), so on a shared edge one of them can accept a triangle the other one rounds out
        some = np.flatnonzero(hit)[:1 << 20]
        near, far = rays[some].copy(), rays[some].copy()
        near["distance"] = hits["distance"][some] * np.float32(0.999)
        far["distance"] = hits["distance"][some] * np.float32(1.001) + np.float32(1e-4)
        near_occluded, far_occluded = scene.occlude(near), scene.occlude(far)
        assert near_occluded.mean() < 1e-5 and far_occluded.mean() > 1 - 1e-5
        assert np.array_equal(near_occluded, oracle.occlude(near)) and np.array_equal(far_occluded, oracle.occlude(far))

        # a closest hit limited to just beyond the found distance finds the same primitive; limited to just short, nothing
        limited = rays[some].copy()
        limited["distance"] = far["distance"]
        again = scene.trace(limited)
        assert np.array_equal(again["token"], hits["token"][some]) and np.array_equal(again["distance"].view(np.uint32), hits["distance"][some].view(np.uint32))
        limited["distance"] = near["distance"]
        assert (scene.trace(limited)["token"] == structs.TOKEN_EMPTY).all()

        # idempotence: the persistent kernel hands rays to lanes in a timing-dependent order; results must not depend on it
        assert np.array_equal(scene.trace(rays).view(np.uint8), hits.view(np.uint8))


def relative_rmse(actual, expected):
    return float(np.sqrt(np.mean((actual - expected) ** 2)) / max(np.sqrt(np.mean(expected ** 2)), 1e-12))


@pytest.mark.parametrize("name,bounce_limit", [("mixed", 8), ("lights", 128)])
def test_full_resolution_tiles_match_oracle(name, bounce_limit):
    """C3 (mixed materials, ~300 k triangles, depth 8) and C4 (10 000 emissive triangles, light tree of 19 999 nodes) at
    1920 x 1080: every 127th tile of the Hilbert sequence (64 tiles over the whole frame), 16 spp, per-pixel means against
    the oracle with the north-star's image tolerance (rel. RMSE <= 1e-3; observed: bit-identical samples)."""
    description = scenes.mixed_material_scene() if name == "mixed" else scenes.many_lights_scene()
    prepared = host.prepare(description)
    oracle = oracle_lib.OracleScene(prepared)
    width, height, tile = 1920, 1080, 16
    tiles = np.ascontiguousarray(hilbert_curve_pattern(((width + tile - 1) // tile, (height + tile - 1) // tile))[::127][:64])
    params = structs.render_params(width, height, tile, extend=16, min_epoch=1, max_epoch=1, bounce_limit=bounce_limit, seed=1)

    with PreparedScene(prepared) as scene:
        actual, stats = scene.render_tiles(params, tiles)

    expected, expected_stats = oracle.render_tiles(params, tiles)
    assert np.isfinite(actual).all() and expected[..., :3].max() > 0
    assert relative_rmse(actual, expected) <= 1e-4
    different = np.any(actual.view(np.uint32) != expected.view(np.uint32), axis=-1)
    assert different.mean() <= 1e-3, f"{different.sum()} of {different.size} pixels are not bit-identical"
    inside = (np.minimum(tiles[:, 0] * tile + tile, width) - tiles[:, 0] * tile) * (np.minimum(tiles[:, 1] * tile + tile, height) - tiles[:, 1] * tile)
    assert int(stats["sampleEvaluated"][0]) == int(expected_stats["sampleEvaluated"][0]) == int(inside.sum()) * 16  # the top row of tiles is half outside
    for field in ("bounceCreated", "lightSampled", "lightOcclusionChecked", "traceQueries", "occludeQueries"):
        assert int(stats[field][0]) == int(expected_stats[field][0]) > 0, field


def test_c5_ten_million_triangles():
    """C5: 9 998 246 triangles, 4 948 274 QBVH nodes (633 MB), quad depth 16 — the largest scene of BASELINE.json. Closest hit and
    occlusion of 2 Mi incoherent rays bit for bit, then 64 tiles spread over the 3840 x 2160 frame at bounce limit 128."""
    prepared = host.prepare(scenes.large_scene())
    assert len(prepared.triangles) > 9_900_000 and prepared.max_depth >= 14
    oracle = oracle_lib.OracleScene(prepared)
    rays = scenes.random_rays(prepared.bounds, 1 << 21, seed=11)
    shadow = scenes.random_rays(prepared.bounds, 1 << 21, seed=11, occlusion=True)
    width, height, tile = 3840, 2160, 16
    tiles = np.ascontiguousarray(hilbert_curve_pattern((width // tile, height // tile))[::509][:64])
    params = structs.render_params(width, height, tile, extend=4, min_epoch=1, max_epoch=1, bounce_limit=128, seed=1)

    with PreparedScene(prepared) as scene:
        hits, expected = scene.trace(rays), oracle.trace(rays)
        assert np.array_equal(hits["token"], expected["token"])
        assert np.array_equal(hits["distance"].view(np.uint32), expected["distance"].view(np.uint32))
        hit = expected["token"] != structs.TOKEN_EMPTY
        assert hit.mean() > 0.05 and np.array_equal(hits["uv"][hit].view(np.uint32), expected["uv"][hit].view(np.uint32))
        assert np.array_equal(scene.occlude(shadow), oracle.occlude(shadow))
        secondary = scenes.secondary_rays(prepared, rays, expected)
        again, expected_again = scene.trace(secondary), oracle.trace(secondary)
        assert np.array_equal(again.view(np.uint8), expected_again.view(np.uint8))
        actual, stats = scene.render_tiles(params, tiles)

    expected, expected_stats = oracle.render_tiles(params, tiles)
    assert expected[..., :3].max() > 0 and relative_rmse(actual, expected) <= 1e-4
    assert np.any(actual.view(np.uint32) != expected.view(np.uint32), axis=-1).mean() <= 1e-3
    for field in ("sampleEvaluated", "bounceCreated", "lightSampled", "traceQueries", "occludeQueries"):
        assert int(stats[field][0]) == int(expected_stats[field][0]) > 0, field
