"""Parity of the CUDA traversal through instanced packs with the oracle, through the C ABI:
hit tokens, the instance layers of every hit, distances and barycentrics bit-exact, with and without ignore hierarchies."""
import numpy as np
import pytest

from echorenderer_b200 import EchoNativeError, PreparedScene, host, scenes, structs
from tests import oracle_lib
from tests.test_gpu_trace import assert_hits_equal
from tests.test_instancing import spawn_from_hits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def instanced():
    return host.prepare(scenes.instanced_scene())


def test_trace_hierarchy_matches_oracle(instanced):
    oracle = oracle_lib.OracleScene(instanced)
    rays = scenes.random_rays(instanced.bounds, 300_000, seed=11)

    with PreparedScene(instanced) as scene:
        hits, layers = scene.trace_hierarchy(rays)
        expected, expected_layers = oracle.trace_hierarchy(rays)
        hit = expected["token"] != structs.TOKEN_EMPTY
        assert hit.mean() > 0.1 and set(np.unique(expected_layers["instanceCount"][hit])) == {0, 1, 2}
        assert_hits_equal(hits, expected)
        assert np.array_equal(layers, expected_layers)

        # the plain batch call on an instanced scene: same hits, layers dropped
        assert_hits_equal(scene.trace(rays), expected)

        # rays leaving the hit points, ignoring the full hierarchy they start on
        spawned, ignore = spawn_from_hits(rays, expected, expected_layers)
        again, again_layers = scene.trace_hierarchy(spawned, ignore)
        expected_again, expected_again_layers = oracle.trace_hierarchy(spawned, ignore)
        assert_hits_equal(again, expected_again)
        assert np.array_equal(again_layers, expected_again_layers)

        # and with the layers left out (a different ignore hierarchy for hits inside placements)
        naive, naive_layers = scene.trace_hierarchy(spawned, None)
        expected_naive, expected_naive_layers = oracle.trace_hierarchy(spawned, None)
        assert_hits_equal(naive, expected_naive)
        assert np.array_equal(naive_layers, expected_naive_layers)


def test_occlude_hierarchy_matches_oracle(instanced):
    oracle = oracle_lib.OracleScene(instanced)
    rays = scenes.random_rays(instanced.bounds, 300_000, seed=13, occlusion=True)

    with PreparedScene(instanced) as scene:
        expected = oracle.occlude_hierarchy(rays)
        assert 0.02 < expected.mean() < 0.98
        assert np.array_equal(scene.occlude_hierarchy(rays), expected)
        assert np.array_equal(scene.occlude(rays), expected)

        primary = scenes.random_rays(instanced.bounds, 200_000, seed=17)
        hits, layers = oracle.trace_hierarchy(primary)
        spawned, ignore = spawn_from_hits(primary, hits, layers)
        spawned["distance"] = 4.0
        assert np.array_equal(scene.occlude_hierarchy(spawned, ignore), oracle.occlude_hierarchy(spawned, ignore))


def test_edge_cases(instanced):
    oracle = oracle_lib.OracleScene(instanced)

    with PreparedScene(instanced) as scene:
        empty = np.zeros(0, dtype=structs.RAY)
        hits, layers = scene.trace_hierarchy(empty)
        assert len(hits) == 0 and len(layers) == 0

        rays = scenes.random_rays(instanced.bounds, 1001, seed=19)  # ragged size, zero / negative / tiny limits
        rays["distance"][::5] = 0.0
        rays["distance"][1::5] = -1.0
        rays["distance"][2::5] = 3.0
        hits, layers = scene.trace_hierarchy(rays)
        expected, expected_layers = oracle.trace_hierarchy(rays)
        assert_hits_equal(hits, expected)
        assert np.array_equal(layers, expected_layers)
        assert np.array_equal(scene.occlude_hierarchy(rays), oracle.occlude_hierarchy(rays))


def test_commit_rejects_broken_packs(instanced):
    import copy
    broken = copy.copy(instanced)
    broken.instances = instanced.instances.copy()
    broken.instances["pack"][0] = 0  # a placement of the scene itself: cyclic
    with pytest.raises(EchoNativeError):
        PreparedScene(broken)

    broken = copy.copy(instanced)
    broken.packs = instanced.packs.copy()
    broken.packs["instanceCount"][0] += 100
    with pytest.raises(EchoNativeError):
        PreparedScene(broken)


def test_deep_nesting():
    """Five instance layers (TokenHierarchy.MaxLayer): a chain of packs each holding one placement of the next."""
    leaf = host.PackDescription(triangles=scenes.box(0, (1, 1, 1)), materials=scenes.material(structs.MATERIAL_DIFFUSE))
    packs = [leaf]
    for level in range(1, 5):
        packs.append(host.PackDescription(triangles=scenes.box(0, (0.3, 0.3, 0.3), (2.0, 0, 0)), materials=scenes.material(structs.MATERIAL_DIFFUSE),
                                          instances=[host.InstanceDescription(level - 1, (0.1 * level, 0.2, 0), (10 * level, 25, 0), 0.9)]))
    description = host.SceneDescription(triangles=scenes.plane(0, (20, 20), (0, -3, 0)), materials=scenes.material(structs.MATERIAL_DIFFUSE),
                                        instances=[host.InstanceDescription(4, (0, 0, 0), (0, 15, 0), 1.5)], packs=packs, camera=scenes.cornell_box().camera)
    prepared = host.prepare(description)
    oracle = oracle_lib.OracleScene(prepared)
    rays = scenes.random_rays(prepared.bounds, 100_000, seed=23)

    with PreparedScene(prepared) as scene:
        hits, layers = scene.trace_hierarchy(rays)
        expected, expected_layers = oracle.trace_hierarchy(rays)
        assert expected_layers["instanceCount"].max() == 5
        assert_hits_equal(hits, expected)
        assert np.array_equal(layers, expected_layers)


@pytest.mark.parametrize("nested", [False, True])
def test_samples_match_oracle(nested):
    """Per-sample radiance of the wavefront path tracer on an instanced scene: Interact through FindLayer, lights picked and
    sampled through the instance layers, shadow rays with full ignore hierarchies — bit-identical to the oracle."""
    from tests.test_gpu_render import relative_rmse, sample_grid
    prepared = host.prepare(scenes.instanced_scene(grid=4, rings=12, segments=14, nested=nested))
    oracle = oracle_lib.OracleScene(prepared)
    width, height = 96, 64
    params = structs.render_params(width, height, 16, extend=4, bounce_limit=16, seed=7)
    pixel_xy, sample_index = sample_grid(width, height, 4)

    with PreparedScene(prepared) as scene:
        actual = scene.evaluate_samples(params, pixel_xy, sample_index)

    expected = oracle.evaluate_samples(params, pixel_xy, sample_index)
    assert np.isfinite(actual).all() and expected.max() > 0
    different = np.any(actual.view(np.uint32) != expected.view(np.uint32), axis=1)
    assert different.mean() <= 1e-4, f"{different.sum()} of {len(different)} samples are not bit-identical"
    assert relative_rmse(actual, expected) <= 1e-4


def test_render_tiles_match_oracle(instanced):
    from tests.test_gpu_render import relative_rmse
    oracle = oracle_lib.OracleScene(instanced)
    width, height = 80, 48
    params = structs.render_params(width, height, 32, extend=8, min_epoch=2, max_epoch=2, bounce_limit=16, seed=3)
    tiles = scenes.tile_grid(width, height, 32)

    with PreparedScene(instanced) as scene:
        actual, stats = scene.render_tiles(params, tiles)

    expected, expected_stats = oracle.render_tiles(params, tiles)
    assert relative_rmse(actual, expected) <= 1e-4
    for name in structs.STATS_FIELDS[:12]:
        assert abs(int(stats[name][0]) - int(expected_stats[name][0])) <= 1e-4 * max(1, int(expected_stats[name][0])), name


def test_auxiliary_evaluators_match_oracle(instanced):
    from tests.test_gpu_render import sample_grid
    oracle = oracle_lib.OracleScene(instanced)
    width, height = 96, 64
    pixel_xy, sample_index = sample_grid(width, height, 2)

    with PreparedScene(instanced) as scene:
        for evaluator in (structs.EVALUATOR_ALBEDO | structs.EVALUATOR_DIVERGE_ONCE, structs.EVALUATOR_NORMAL_DEPTH):
            params = structs.render_params(width, height, 32, extend=2, seed=7, evaluator=evaluator)
            expected = np.zeros((len(sample_index), 4), dtype=np.float32)
            oracle.lib.oracle_evaluate_samples4(oracle.handle, oracle_lib.ptr(params), oracle_lib.ptr(pixel_xy), oracle_lib.ptr(sample_index), len(sample_index),
                                                oracle_lib.ptr(expected), 4, 0)
            actual = scene.evaluate_samples(params, pixel_xy, sample_index, channels=4)
            assert np.array_equal(actual.view(np.uint32), expected.view(np.uint32))
            assert len(np.unique(expected[:, :3], axis=0)) > 4


def test_textured_placements_match_oracle():
    """Image textures inside instanced packs: the pack's own swatch and a placement's replacement swatch bind different
    textures; the environment is an importance-sampled map. Per-sample radiance and the albedo pass against the oracle."""
    from tests.conftest import sky_texture
    from tests.test_gpu_render import sample_grid
    description = scenes.instanced_scene(grid=3, rings=10, segments=12)
    rng = np.random.default_rng(8)
    noise = np.concatenate([rng.uniform(0.1, 0.9, (16, 16, 3)), np.ones((16, 16, 1))], axis=-1)
    stripes = np.where((np.arange(16) % 2 == 0)[None, :, None], np.array([0.9, 0.2, 0.1]), np.array([0.1, 0.3, 0.9])) * np.ones((16, 1, 1))
    stripes = np.concatenate([stripes, np.ones((16, 16, 1))], axis=-1)
    description.textures = [host.TextureDescription(noise, structs.FILTER_BILINEAR, structs.WRAPPER_REPEAT),
                            host.TextureDescription(stripes, structs.FILTER_POINT, structs.WRAPPER_MIRROR),
                            host.TextureDescription(sky_texture(16, 32), structs.FILTER_BILINEAR, structs.WRAPPER_REPEAT)]
    blob = description.packs[0]
    blob.material_textures = structs.material_textures(len(blob.materials))
    blob.material_textures["albedo"][0] = 0
    for instance in description.instances:
        if instance.materials is not None:
            instance.material_textures = structs.material_textures(len(instance.materials))
            instance.material_textures["albedo"][1] = 1
            instance.material_textures["roughness"][0] = 0
    description.material_textures = structs.material_textures(len(description.materials))
    description.material_textures["albedo"][0] = 1
    description.infinite_lights = scenes.environment_light(2, (0.6, 0.6, 0.7), (0, 25, 0))

    prepared = host.prepare(description)
    assert (prepared.material_textures["albedo"] != structs.TEXTURE_NONE).sum() >= 3
    oracle = oracle_lib.OracleScene(prepared)
    width, height = 96, 64
    pixel_xy, sample_index = sample_grid(width, height, 2)

    with PreparedScene(prepared) as scene:
        for evaluator, channels in ((structs.EVALUATOR_PATH_TRACED, 3), (structs.EVALUATOR_ALBEDO | structs.EVALUATOR_DIVERGE_ONCE, 4)):
            params = structs.render_params(width, height, 32, extend=2, bounce_limit=12, seed=5, evaluator=evaluator)
            expected = np.zeros((len(sample_index), 4), dtype=np.float32)
            oracle.lib.oracle_evaluate_samples4(oracle.handle, oracle_lib.ptr(params), oracle_lib.ptr(pixel_xy), oracle_lib.ptr(sample_index), len(sample_index),
                                                oracle_lib.ptr(expected), 4, 0)
            actual = scene.evaluate_samples(params, pixel_xy, sample_index, channels=4)
            assert np.array_equal(actual.view(np.uint32), expected.view(np.uint32))
            assert len(np.unique(expected[:, :3].round(3), axis=0)) > 50


def test_naive_evaluator_matches_oracle(instanced):
    from tests.test_gpu_render import sample_grid
    oracle = oracle_lib.OracleScene(instanced)
    params = structs.render_params(64, 48, 16, extend=2, bounce_limit=24, seed=7, evaluator=structs.EVALUATOR_NAIVE)
    pixel_xy, sample_index = sample_grid(64, 48, 2)
    with PreparedScene(instanced) as scene:
        actual = scene.evaluate_samples(params, pixel_xy, sample_index)
    expected = oracle.evaluate_samples(params, pixel_xy, sample_index)
    assert np.array_equal(actual.view(np.uint32), expected.view(np.uint32))
