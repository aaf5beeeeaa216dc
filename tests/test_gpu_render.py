"""Parity of the wavefront path tracer with the oracle's per-sample PathTracedEvaluator, through the C ABI.

Device and oracle share the sample sequence and a deterministic sincos, and every other operation is an IEEE-rounded
+, -, *, /, sqrt or fma in the same order, so a sample's radiance is expected to be bit-identical. The tests allow a
tiny budget of differing samples (none observed so far) and bound the image error far below the north-star tolerance
(relative RMSE <= 1e-3 at matched spp)."""
import numpy as np
import pytest

from echorenderer_b200 import EvaluationOperation, EvaluationProfile, PathTracedEvaluator, PreparedScene, RenderTexture, scenes, structs
from tests import oracle_lib

pytestmark = pytest.mark.gpu


def sample_grid(width, height, spp, stride=1):
    ys, xs = np.meshgrid(np.arange(0, height, stride), np.arange(0, width, stride), indexing="ij")
    pixels = np.stack([xs.reshape(-1), ys.reshape(-1)], axis=-1).astype(np.int32)
    pixel_xy = np.repeat(pixels, spp, axis=0)
    sample_index = np.tile(np.arange(spp, dtype=np.uint32), len(pixels))
    return pixel_xy, sample_index


def relative_rmse(actual, expected):
    return float(np.sqrt(np.mean((actual - expected) ** 2)) / max(np.sqrt(np.mean(expected ** 2)), 1e-12))


@pytest.mark.parametrize("fixture,bounce_limit", [("cornell", 128), ("mixed_small", 8), ("lights_small", 128), ("terrain_small", 16), ("coated_small", 12),
                                                  ("directional_small", 12), ("textured_small", 12), ("environment_small", 12), ("cubemap_small", 12)])
def test_samples_match_oracle(fixture, bounce_limit, request):
    prepared = request.getfixturevalue(fixture)
    oracle = oracle_lib.OracleScene(prepared)
    width, height = 96, 64
    params = structs.render_params(width, height, 16, extend=4, bounce_limit=bounce_limit, seed=7)
    pixel_xy, sample_index = sample_grid(width, height, 4)

    with PreparedScene(prepared) as scene:
        actual = scene.evaluate_samples(params, pixel_xy, sample_index)

    expected = oracle.evaluate_samples(params, pixel_xy, sample_index)
    assert np.isfinite(actual).all()
    assert expected.max() > 0

    different = np.any(actual.view(np.uint32) != expected.view(np.uint32), axis=1)
    assert different.mean() <= 1e-4, f"{different.sum()} of {len(different)} samples are not bit-identical"
    assert relative_rmse(actual, expected) <= 1e-4


@pytest.mark.parametrize("fixture", ["cornell", "mixed_small", "directional_small", "textured_small", "environment_small", "cubemap_small"])
def test_render_tiles_match_oracle(fixture, request):
    prepared = request.getfixturevalue(fixture)
    oracle = oracle_lib.OracleScene(prepared)
    width, height = 80, 48  # ragged: partial tiles on both edges
    params = structs.render_params(width, height, 32, extend=8, min_epoch=2, max_epoch=2, bounce_limit=16, seed=3)
    tiles = scenes.tile_grid(width, height, 32)

    with PreparedScene(prepared) as scene:
        actual, stats = scene.render_tiles(params, tiles)

    expected, expected_stats = oracle.render_tiles(params, tiles)
    assert relative_rmse(actual, expected) <= 1e-4

    # the statistics rows of EvaluatorStatistics agree exactly when every sample takes the same decisions
    for name in structs.STATS_FIELDS[:12]:
        assert abs(int(stats[name][0]) - int(expected_stats[name][0])) <= 1e-4 * max(1, int(expected_stats[name][0])), name
    assert int(stats["sampleEvaluated"][0]) == width * height * 16
    assert int(stats["kernelLaunches"][0]) > 0


def test_adaptive_epochs_match_oracle(cornell):
    """MinEpoch < MaxEpoch: the noise-driven epoch loop (EvaluationOperation.cs:137) stops per pixel like the oracle's."""
    oracle = oracle_lib.OracleScene(cornell)
    width = height = 32
    params = structs.render_params(width, height, 16, extend=8, min_epoch=1, max_epoch=6, noise_threshold=0.3, seed=5)
    tiles = scenes.tile_grid(width, height, 16)

    with PreparedScene(cornell) as scene:
        actual, stats = scene.render_tiles(params, tiles)

    expected, expected_stats = oracle.render_tiles(params, tiles)
    assert int(stats["sampleEvaluated"][0]) == int(expected_stats["sampleEvaluated"][0])
    assert width * height * 8 < int(stats["sampleEvaluated"][0]) < width * height * 48
    assert relative_rmse(actual, expected) <= 1e-4


def test_evaluation_operation_interface(cornell):
    """The host-side mirror of EvaluationOperation: profile validation, tile application, TotalSamples, statistics labels."""
    with pytest.raises(ValueError):
        EvaluationProfile(extend=0).validate()
    with pytest.raises(ValueError):
        EvaluationProfile(min_epoch=3, max_epoch=2).validate()

    with PreparedScene(cornell) as scene:
        destination = RenderTexture(64, 64, 16)
        profile = EvaluationProfile(evaluator=PathTracedEvaluator(), extend=4, min_epoch=1, max_epoch=1)
        operation = EvaluationOperation(scene, profile, destination)
        operation.execute()

        assert operation.total_samples == 64 * 64 * 4
        report = operation.statistics_report()
        assert report["Pixel/Evaluated"] == 64 * 64
        assert report["Bounce/Created"] > 0
        image = destination.pixels
        assert np.isfinite(image).all() and image[..., :3].mean() > 0.01
        # the red wall is on the left, the green wall on the right (CornellBox.cs:44-45)
        assert image[24:40, 2:8, 0].mean() > 2 * image[24:40, 2:8, 1].mean()
        assert image[24:40, -8:-2, 1].mean() > 2 * image[24:40, -8:-2, 0].mean()


@pytest.mark.parametrize("fixture", ["cornell", "mixed_small", "coated_small", "textured_small", "environment_small"])
@pytest.mark.parametrize("evaluator", [structs.EVALUATOR_ALBEDO | structs.EVALUATOR_DIVERGE_ONCE, structs.EVALUATOR_ALBEDO,
                                       structs.EVALUATOR_NORMAL_DEPTH, structs.EVALUATOR_NORMAL_DEPTH | structs.EVALUATOR_DIVERGE_ONCE])
def test_auxiliary_evaluators_match_oracle(fixture, evaluator, request):
    """AlbedoEvaluator / NormalDepthEvaluator (AlbedoEvaluator.cs:18-55, NormalDepthEvaluator.cs:20-60): all four lanes of every
    sample bit-identical to the oracle, then the accumulated tiles."""
    prepared = request.getfixturevalue(fixture)
    oracle = oracle_lib.OracleScene(prepared)
    width, height = 96, 64
    params = structs.render_params(width, height, 32, extend=4, seed=7, evaluator=evaluator)
    pixel_xy, sample_index = sample_grid(width, height, 4)
    expected = np.zeros((len(sample_index), 4), dtype=np.float32)
    oracle.lib.oracle_evaluate_samples4(oracle.handle, oracle_lib.ptr(params), oracle_lib.ptr(pixel_xy), oracle_lib.ptr(sample_index), len(sample_index),
                                        oracle_lib.ptr(expected), 4, 0)
    tiles = scenes.tile_grid(width, height, 32)
    expected_tiles, expected_stats = oracle.render_tiles(params, tiles)

    with PreparedScene(prepared) as scene:
        actual = scene.evaluate_samples(params, pixel_xy, sample_index, channels=4)
        actual_tiles, stats = scene.render_tiles(params, tiles)

    assert np.array_equal(actual.view(np.uint32), expected.view(np.uint32))
    assert np.array_equal(actual_tiles.view(np.uint32), expected_tiles.view(np.uint32))
    assert int(stats["sampleEvaluated"][0]) == int(expected_stats["sampleEvaluated"][0]) == width * height * 4


def test_unknown_evaluator_is_rejected(cornell):
    from echorenderer_b200 import EchoNativeError
    with PreparedScene(cornell) as scene:
        with pytest.raises(EchoNativeError, match="evaluator"):
            scene.render_tiles(structs.render_params(32, 32, 16, evaluator=7), scenes.tile_grid(32, 32, 16))


def test_frame_sharding_paths(cornell):
    """The multi-device path of bench.py --workload render on one device: render_frame_device accumulates (mean * epochs,
    epochs) per pixel, frames of different "ranks" are summed (the all-reduce), frame_resolve divides. Tile sharding is
    bit-exact against a whole-frame render; sample sharding matches the all-epochs render to rounding."""
    import torch
    from echorenderer_b200 import shard_epochs, shard_tiles
    oracle = oracle_lib.OracleScene(cornell)
    width, height, tile = 80, 48, 16
    tiles = scenes.tile_grid(width, height, tile)
    world = 2

    with PreparedScene(cornell) as scene:
        stream = torch.cuda.current_stream().cuda_stream

        # tile sharding: disjoint tiles, sum with zeros
        params = structs.render_params(width, height, tile, extend=4, min_epoch=2, max_epoch=2, seed=6)
        frames = []
        for rank in range(world):
            frame = torch.zeros(height * width * 4, dtype=torch.float32, device="cuda")
            scene.render_frame_device(params, shard_tiles(tiles, rank, world), frame.data_ptr(), stream)
            frames.append(frame)
        total = frames[0] + frames[1]
        scene.frame_resolve_device(total.data_ptr(), width, height, stream)
        torch.cuda.synchronize()
        reduced = total.cpu().numpy().reshape(height, width, 4)
        expected, _ = oracle.render_tiles(params, tiles)
        whole = scenes.assemble_tiles(expected, tiles, width, height, tile)
        assert np.array_equal(reduced[..., :3].view(np.uint32), whole[..., :3].view(np.uint32))

        # sample sharding: every rank renders its block of epochs of all tiles
        epochs = 4
        frames = []
        for rank in range(world):
            first, count = shard_epochs(epochs, rank, world)
            block = structs.render_params(width, height, tile, extend=4, min_epoch=count, max_epoch=count, seed=6, epoch_offset=first)
            frame = torch.zeros(height * width * 4, dtype=torch.float32, device="cuda")
            scene.render_frame_device(block, tiles, frame.data_ptr(), stream)
            frames.append(frame)
        total = frames[0] + frames[1]
        torch.cuda.synchronize()
        assert np.all(total.cpu().numpy().reshape(height, width, 4)[..., 3] == epochs * 4)  # the weights add up to the sample count
        scene.frame_resolve_device(total.data_ptr(), width, height, stream)
        torch.cuda.synchronize()
        reduced = total.cpu().numpy().reshape(height, width, 4)
        everything = structs.render_params(width, height, tile, extend=4, min_epoch=epochs, max_epoch=epochs, seed=6)
        expected, _ = oracle.render_tiles(everything, tiles)
        whole = scenes.assemble_tiles(expected, tiles, width, height, tile)
        assert np.allclose(reduced[..., :3], whole[..., :3], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("fixture", ["cornell", "mixed_small", "textured_small"])
def test_naive_evaluator_matches_oracle(fixture, request):
    """StandardNaiveEvaluator (StandardNaiveEvaluator.cs:16-55): the recursion folded innermost-first, bit-identical per sample."""
    prepared = request.getfixturevalue(fixture)
    oracle = oracle_lib.OracleScene(prepared)
    width, height = 64, 48
    params = structs.render_params(width, height, 16, extend=4, bounce_limit=40, seed=7, evaluator=structs.EVALUATOR_NAIVE)
    pixel_xy, sample_index = sample_grid(width, height, 4)

    with PreparedScene(prepared) as scene:
        actual = scene.evaluate_samples(params, pixel_xy, sample_index)
        with pytest.raises(Exception, match="bounce limits"):
            scene.render_tiles(structs.render_params(32, 32, 16, bounce_limit=500, evaluator=structs.EVALUATOR_NAIVE), scenes.tile_grid(32, 32, 16))

    expected = oracle.evaluate_samples(params, pixel_xy, sample_index)
    assert expected.max() > 0
    assert np.array_equal(actual.view(np.uint32), expected.view(np.uint32))


@pytest.mark.parametrize("camera", ["orthographic", "cylindrical", "thin_lens"])
def test_cameras_match_oracle(camera):
    """OrthographicCamera, CylindricalCamera and the thin-lens PerspectiveCamera (Scenic/Cameras/*.cs) through the path tracer."""
    from echorenderer_b200 import host
    description = scenes.cornell_box()
    if camera == "orthographic":
        description.camera = scenes.orthographic_camera((0, 5, -4), (0, 0, 0), width=9.0)
    elif camera == "cylindrical":
        description.camera = scenes.cylindrical_camera((0, 5, 0), (10, 30, 0))
    else:
        description.camera = scenes.perspective_camera((0, 5, -18.025444), field_of_view=42.0, lens_radius=0.3, focal_distance=16.0)
    prepared = host.prepare(description)
    oracle = oracle_lib.OracleScene(prepared)
    width, height = 64, 32
    params = structs.render_params(width, height, 16, extend=4, bounce_limit=24, seed=9)
    pixel_xy, sample_index = sample_grid(width, height, 4)

    with PreparedScene(prepared) as scene:
        actual = scene.evaluate_samples(params, pixel_xy, sample_index)

    expected = oracle.evaluate_samples(params, pixel_xy, sample_index)
    assert expected.max() > 0 and (expected.sum(axis=1) > 0).mean() > 0.5
    assert np.array_equal(actual.view(np.uint32), expected.view(np.uint32))
