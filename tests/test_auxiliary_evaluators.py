"""AlbedoEvaluator / NormalDepthEvaluator (Evaluation/Evaluators/AlbedoEvaluator.cs:18-55,
NormalDepthEvaluator.cs:20-60) on the oracle: the auxiliary passes StandardPathTracedProfile renders for its denoiser."""
import numpy as np
import pytest

from echorenderer_b200 import scenes, structs

from . import oracle_lib

ALBEDO = structs.EVALUATOR_ALBEDO | structs.EVALUATOR_DIVERGE_ONCE        # AlbedoEvaluator default: DivergeOnce = true
NORMAL_DEPTH = structs.EVALUATOR_NORMAL_DEPTH                             # NormalDepthEvaluator default: DivergeOnce = false


def grid(width, height, stride=1):
    ys, xs = np.meshgrid(np.arange(0, height, stride), np.arange(0, width, stride), indexing="ij")
    pixels = np.stack([xs.reshape(-1), ys.reshape(-1)], axis=-1).astype(np.int32)
    return pixels, np.zeros(len(pixels), dtype=np.uint32)


def evaluate4(oracle, params, pixels, index):
    out = np.zeros((len(index), 4), dtype=np.float32)
    oracle.lib.oracle_evaluate_samples4(oracle.handle, oracle_lib.ptr(params), oracle_lib.ptr(pixels), oracle_lib.ptr(index), len(index), oracle_lib.ptr(out), 4, 0)
    return out


def first_hits(oracle, params, pixels, index):
    rays = oracle.spawn_rays(params, pixels, index)
    return rays, oracle.trace(rays)


def test_albedo_of_directly_visible_surfaces(cornell):
    """No specular surface in the Cornell box: the albedo pass is the albedo of the first hit (the emission colour on the light)."""
    oracle = oracle_lib.OracleScene(cornell)
    params = structs.render_params(64, 64, 16, extend=1, seed=2, evaluator=ALBEDO)
    pixels, index = grid(64, 64)
    value = evaluate4(oracle, params, pixels, index)
    rays, hits = first_hits(oracle, params, pixels, index)

    hit = hits["token"] != structs.TOKEN_EMPTY
    assert hit.mean() > 0.9 and np.all(value[:, 3] == 0)
    material = cornell.triangles["material"][structs.token_index(hits["token"][hit])]
    expected = cornell.materials["albedo"][material][:, :3]
    # the front wall is OneSided: seen from outside it is invisible and the pass reports what lies behind it
    one_sided = cornell.materials["type"][material] == structs.MATERIAL_ONESIDED
    assert np.array_equal(value[hit][~one_sided][:, :3], expected[~one_sided])
    assert len(np.unique(value[:, :3], axis=0)) >= 4  # white, red, green, the light


def test_normal_depth_of_directly_visible_surfaces(cornell):
    oracle = oracle_lib.OracleScene(cornell)
    params = structs.render_params(64, 64, 16, extend=1, seed=2, evaluator=NORMAL_DEPTH)
    pixels, index = grid(64, 64)
    value = evaluate4(oracle, params, pixels, index)
    rays, hits = first_hits(oracle, params, pixels, index)

    hit = hits["token"] != structs.TOKEN_EMPTY
    material = cornell.triangles["material"][structs.token_index(hits["token"][hit])]
    plain = cornell.materials["type"][material] != structs.MATERIAL_ONESIDED
    assert np.array_equal(value[hit][plain][:, 3], hits["distance"][hit][plain])           # depth = distance of the first hit
    assert np.allclose(np.linalg.norm(value[:, :3], axis=1), 1.0, atol=1e-5)               # unit normals (or -direction)
    triangles = cornell.triangles[structs.token_index(hits["token"][hit])][plain]
    assert np.allclose(value[hit][plain][:, :3], triangles["normal0"], atol=1e-6)          # flat shading normals

    escaped = ~hit
    if escaped.any():
        assert np.allclose(value[escaped][:, :3], -rays["direction"][escaped], atol=0)
        assert np.all(value[escaped][:, 3] == np.float32(cornell.bound_radius) * 2)


def test_specular_surfaces_are_followed(mixed_small):
    """Through the specular glass blob the passes report what is seen in or behind it; DivergeOnce decides how far they follow."""
    oracle = oracle_lib.OracleScene(mixed_small)
    width, height = 96, 54
    pixels, index = grid(width, height)
    _, hits = first_hits(oracle, structs.render_params(width, height, 16, extend=1, seed=2), pixels, index)
    hit = hits["token"] != structs.TOKEN_EMPTY
    first_material = np.full(len(hits), -1)
    triangle = hit & (structs.token_type(hits["token"]) == structs.TOKEN_TYPE_TRIANGLE)
    first_material[triangle] = mixed_small.triangles["material"][structs.token_index(hits["token"][triangle])]
    glass = first_material == 2  # specular Dielectric (scenes.mixed_materials)
    assert glass.sum() > 20

    once = evaluate4(oracle, structs.render_params(width, height, 16, extend=1, seed=2, evaluator=ALBEDO), pixels, index)
    never = evaluate4(oracle, structs.render_params(width, height, 16, extend=1, seed=2, evaluator=structs.EVALUATOR_ALBEDO), pixels, index)
    swatch = mixed_small.materials["albedo"][:, :3]
    ambient = mixed_small.description.infinite_lights["radiance"]
    allowed = np.concatenate([swatch, ambient, np.zeros((1, 3), dtype=np.float32)])

    for value in (once, never):
        distance = np.abs(value[:, None, :3] - allowed[None]).max(axis=2).min(axis=1)
        assert np.all(distance == 0)  # every value is a swatch albedo, or the ambient light of an escaped ray

    # DivergeOnce = false stops at the glass itself once the bounce leaves the camera direction: white glass albedo
    assert np.all(never[glass][:, :3] == 1.0)
    # DivergeOnce = true follows one diverging bounce and reports the next surface: the far side of the same glass blob for
    # refracted bounces (white again), something else for the reflected ones
    assert np.mean(np.all(once[glass][:, :3] == 1.0, axis=1)) < 1.0
    assert np.array_equal(once[~glass], never[~glass])

    depth_once = evaluate4(oracle, structs.render_params(width, height, 16, extend=1, seed=2, evaluator=NORMAL_DEPTH | structs.EVALUATOR_DIVERGE_ONCE), pixels, index)
    depth_never = evaluate4(oracle, structs.render_params(width, height, 16, extend=1, seed=2, evaluator=NORMAL_DEPTH), pixels, index)
    assert np.array_equal(depth_never[glass][:, 3], hits["distance"][glass])
    assert np.all(depth_once[glass][:, 3] >= hits["distance"][glass])


def test_render_tiles_accumulates_all_four_lanes(cornell):
    """EvaluationOperation accumulates the evaluator's Float4 whatever it means (EvaluationOperation.cs:124-131): the W lane of
    the NormalDepth pass is the mean depth of the pixel's samples."""
    oracle = oracle_lib.OracleScene(cornell)
    tiles = np.array([[1, 1]], dtype=np.int32)
    params = structs.render_params(48, 48, 16, extend=8, seed=4, evaluator=NORMAL_DEPTH)
    image, stats = oracle.render_tiles(params, tiles)
    assert int(stats["sampleEvaluated"][0]) == 16 * 16 * 8
    assert image[..., 3].min() > 5.0 and image[..., 3].max() < 40.0
    assert np.all(np.linalg.norm(image[..., :3], axis=-1) <= 1.0 + 1e-5)


def test_naive_evaluator_agrees_with_the_path_tracer(cornell):
    """StandardNaiveEvaluator (StandardNaiveEvaluator.cs:16-55) has no light sampling, no MIS and no roulette: it is an independent
    estimator of the image PathTracedEvaluator computes. Their block means agree within Monte-Carlo error."""
    oracle = oracle_lib.OracleScene(cornell)
    tiles = scenes.tile_grid(32, 32, 16)
    naive, _ = oracle.render_tiles(structs.render_params(32, 32, 16, extend=2048, seed=11, bounce_limit=48, evaluator=structs.EVALUATOR_NAIVE), tiles)
    traced, _ = oracle.render_tiles(structs.render_params(32, 32, 16, extend=256, seed=12, bounce_limit=128), tiles)
    a = scenes.assemble_tiles(naive, tiles, 32, 32, 16)[..., :3]
    b = scenes.assemble_tiles(traced, tiles, 32, 32, 16)[..., :3]
    assert np.allclose(a.mean(axis=(0, 1)), b.mean(axis=(0, 1)), rtol=0.05)
    blocks_a, blocks_b = a.reshape(4, 8, 4, 8, 3).mean(axis=(1, 3)), b.reshape(4, 8, 4, 8, 3).mean(axis=(1, 3))
    assert np.mean(np.abs(blocks_a - blocks_b)) < 0.08 * blocks_b.mean()


def test_orthographic_and_cylindrical_cameras(cornell):
    """OrthographicCamera.SpawnRay (OrthographicCamera.cs:33-38) and CylindricalCamera.SpawnRay (CylindricalCamera.cs:27-33)."""
    import copy
    from echorenderer_b200 import host
    ys, xs = np.meshgrid(np.arange(0, 16), np.arange(0, 32), indexing="ij")
    pixels = np.stack([xs.reshape(-1), ys.reshape(-1)], axis=-1).astype(np.int32)
    index = np.zeros(len(pixels), dtype=np.uint32)
    params = structs.render_params(32, 16, 16, extend=1, seed=2)

    description = scenes.cornell_box()
    description.camera = scenes.orthographic_camera((0, 5, -18), (0, 0, 0), width=8.0)
    rays = oracle_lib.OracleScene(host.prepare(description)).spawn_rays(params, pixels, index)
    assert np.all(rays["direction"] == np.array([0, 0, 1], dtype=np.float32))
    assert np.allclose(rays["origin"][:, 2], -18) and np.ptp(rays["origin"][:, 0]) == pytest.approx(8.0 * 31 / 32, rel=0.05)
    assert np.ptp(rays["origin"][:, 1]) == pytest.approx(8.0 * 15 / 32, rel=0.1)  # SpawnX scales both axes by 1 / width

    description = scenes.cornell_box()
    description.camera = scenes.cylindrical_camera((0, 5, 0), (0, 0, 0))
    rays = oracle_lib.OracleScene(host.prepare(description)).spawn_rays(params, pixels, index)
    assert np.allclose(np.linalg.norm(rays["direction"], axis=1), 1.0, atol=1e-6) and np.all(rays["origin"] == np.array([0, 5, 0], dtype=np.float32))
    bottom, top = rays["direction"][pixels[:, 1] == 0], rays["direction"][pixels[:, 1] == 15]
    assert bottom[:, 1].max() < -0.9 and top[:, 1].min() > 0.9  # v = 0 looks down, v = 1 looks up
    middle = rays["direction"][pixels[:, 1] == 8]
    azimuth = np.unwrap(np.arctan2(middle[:, 0], middle[:, 2]))
    assert abs(abs(azimuth[-1] - azimuth[0]) - 2 * np.pi * 31 / 32) < 0.05  # one full turn across the row


def _scene_for(fixture, request):
    if fixture == "directional_cone":
        from echorenderer_b200 import host
        description = scenes.mixed_material_scene(rings=12, segments=12)
        description.infinite_lights = np.concatenate([scenes.ambient_light((0.05, 0.05, 0.05)), scenes.directional_light((3.0, 2.8, 2.5), (55, 20, 0), angle=12.0)])
        return host.prepare(description)
    return request.getfixturevalue(fixture)


def _mean_radiance(oracle, evaluator, extend, seed, bounce_limit=12, **extra):
    width, height = 32, 18
    tiles = scenes.tile_grid(width, height, 16)
    if evaluator is not None:
        extra["evaluator"] = evaluator
    image, _ = oracle.render_tiles(structs.render_params(width, height, 16, extend=extend, seed=seed, bounce_limit=bounce_limit, **extra), tiles, threads=8)
    return scenes.assemble_tiles(image, tiles, width, height, 16)[..., :3].mean(axis=(0, 1)).astype(np.float64)


@pytest.fixture
def failed_pick_keeps_mis():
    """Test-only oracle switch (oracle/evaluation.hpp `failedPickKeepsMis`): see test_failed_light_pick_excess_of_the_reference."""
    lib = oracle_lib.library()
    lib.oracle_set_failed_pick_keeps_mis(1)
    yield
    lib.oracle_set_failed_pick_keeps_mis(0)


@pytest.mark.parametrize("fixture", ["mixed_small", "environment_small", "lights_small", "directional_cone"])
def test_naive_evaluator_agrees_on_other_light_transport(fixture, request, failed_pick_keeps_mis):
    """The same independent check on the other samplers of the unpinned integrator: rough / specular dielectrics and conductors
    with area lights (mixed), the importance-sampled environment map, the light tree over hundreds of oriented emitters, and a
    DirectionalLight cone. (A delta light cannot be found by brute force, so it is not in this list.) Run with the one
    statistically visible quirk of the reference neutralised (next test), so the tolerance is Monte-Carlo error only."""
    oracle = oracle_lib.OracleScene(_scene_for(fixture, request))
    # the directly visible pin-point emitters of lights_small are what is noisy (in both estimators): more seeds there
    seeds = 6 if fixture == "lights_small" else 2
    naive = np.mean([_mean_radiance(oracle, structs.EVALUATOR_NAIVE, 4096, seed) for seed in range(1, 1 + seeds)], axis=0)
    traced = np.mean([_mean_radiance(oracle, None, 512, seed, survivability=1e9) for seed in range(11, 11 + seeds)], axis=0)
    assert np.allclose(naive, traced, rtol=0.04), (naive, traced)


def _oriented_emitters(count, size, seed=23):
    """A diffuse ground plane under `count` one-sided emissive triangles of random orientation: big enough for brute force to
    converge, oriented so that light-tree branches whose two children both face away from a shading point are common."""
    from echorenderer_b200 import host
    from echorenderer_b200.scenes import F32, uniform, uniform_sphere_directions, make_triangles
    j = np.arange(count, dtype=np.uint64)
    centre = np.stack([(uniform(seed, j, 0) * 2 - 1) * F32(10.0), F32(1.0) + uniform(seed, j, 1) * F32(6.0), (uniform(seed, j, 2) * 2 - 1) * F32(10.0)], axis=-1).astype(np.float64)
    axis_a = uniform_sphere_directions(uniform(seed, j, 3), uniform(seed, j, 4)).astype(np.float64)
    axis_b = np.cross(axis_a, uniform_sphere_directions(uniform(seed, j, 5), uniform(seed, j, 6)).astype(np.float64))
    axis_b /= np.linalg.norm(axis_b, axis=1, keepdims=True)
    triangles = np.concatenate([scenes.plane(0, (60, 60)), make_triangles(centre, centre + axis_a * size, centre + axis_b * size, 1)])
    materials = np.concatenate([scenes.material(structs.MATERIAL_DIFFUSE, (0.75, 0.75, 0.75)), scenes.material(structs.MATERIAL_EMISSIVE, (10.0, 10.0, 10.0))])
    position = (0.0, 9.0, -22.0)
    camera = scenes.perspective_camera(position, scenes.look_rotation(position, (0, 2, 0)), field_of_view=45.0)
    return host.prepare(scenes.SceneDescription(triangles=triangles, materials=materials, camera=camera, name="oriented_emitters"))


def test_failed_light_pick_excess_of_the_reference():
    """A finding about the reference, kept on purpose. `ImportanceSampleRadiant` reports `mis = false` when `scene.Pick` fails
    (PathTracedEvaluator.cs:165-169) — which happens whenever the descent reaches a light-tree branch whose two children both
    have zero importance from the shading point (LightTree.cs:122). The loop then takes the no-MIS fallback (:138-149) and adds
    the emission the BSDF-sampled ray finds with weight 1, although those emitters are also reached by Pick on other draws:
    an excess of P(pick fails) x w_light. Direct light on a plane under 64 oriented emitters, against brute force
    (StandardNaiveEvaluator, one bounce deeper because its depth counts the camera ray): the restated reference is 3-5 % high,
    and exact within Monte-Carlo error once a failed pick keeps MIS. Oracle and device both keep the reference's behaviour."""
    oracle = oracle_lib.OracleScene(_oriented_emitters(64, 1.0))
    lib = oracle_lib.library()
    naive = np.mean([_mean_radiance(oracle, structs.EVALUATOR_NAIVE, 4096, seed, bounce_limit=2)[0] for seed in (1, 2, 3)])
    reference = np.mean([_mean_radiance(oracle, None, 256, seed, bounce_limit=1, survivability=1e9)[0] for seed in (11, 12, 13, 14)])
    lib.oracle_set_failed_pick_keeps_mis(1)
    try:
        unbiased = np.mean([_mean_radiance(oracle, None, 256, seed, bounce_limit=1, survivability=1e9)[0] for seed in (11, 12, 13, 14)])
    finally:
        lib.oracle_set_failed_pick_keeps_mis(0)
    assert 1.025 < reference / naive < 1.06, (reference, naive)   # measured 1.042 +- 0.004
    assert abs(unbiased / naive - 1) < 0.012, (unbiased, naive)   # measured 1.006 +- 0.004
