"""A known answer for the WHOLE integrator that needs no second implementation: the furnace. A convex Lambertian object of albedo rho in a
uniform environment of radiance L sends out exactly rho * L in every direction — its whole hemisphere sees the environment, the irradiance is
pi L, the radiosity rho pi L — and a ray that misses it sees L. PathTracedEvaluator (PathTracedEvaluator.cs:43-354) reaches that number through
everything it has: the camera sample, Interact, Material.Scatter, the infinite-light pick (PreparedScene.Pick's threshold, AmbientLight.Sample,
the occlusion query), BSDF.Sample / Evaluate / ProbabilityDensity, the power heuristic on both estimates, Russian roulette and the escape
through EvaluateInfinite. The restated `oracle/` must land on rho * L within Monte-Carlo error; rough (Oren-Nayar) and two-bounce variants
have no closed form and are left to the comparison with the naive evaluator (test_auxiliary_evaluators.py)."""
import numpy as np
import pytest

from echorenderer_b200 import SceneDescription, host, scenes, structs
from tests import oracle_lib

RHO = (0.8, 0.5, 0.3)
RADIANCE = (1.5, 1.0, 0.5)


def furnace(shape):
    materials = np.concatenate([scenes.material(structs.MATERIAL_DIFFUSE, RHO)])
    spheres = np.zeros(0, dtype=structs.SPHERE)
    triangles = np.zeros(0, dtype=structs.TRIANGLE)
    if shape == "sphere":
        spheres = np.zeros(2, dtype=structs.SPHERE)  # a tree needs two primitives: the second is a speck a kilometre behind the camera
        spheres["position"], spheres["radius"], spheres["material"] = [(0, 0, 0), (0, 0, -1000)], [1.0, 1e-3], 0
    else:
        triangles = scenes.box(0, (1.6, 1.2, 1.4), rotation=(20, 35, 10))
    position = (0.0, 0.0, -4.0)
    camera = scenes.perspective_camera(position, scenes.look_rotation(position, (0, 0, 0)), field_of_view=40.0, lens_radius=0.0)
    return SceneDescription(triangles=triangles, spheres=spheres, materials=materials, infinite_lights=scenes.ambient_light(RADIANCE), camera=camera, name="furnace")


@pytest.mark.parametrize("shape", ["sphere", "box"])
def test_convex_lambertian_object_in_a_uniform_environment(shape):
    oracle = oracle_lib.OracleScene(host.prepare(furnace(shape)))
    size, extend = 24, 48
    params = structs.render_params(size, size, 8, extend=extend, seed=6, bounce_limit=24)
    ys, xs = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    pixels = np.repeat(np.stack([xs.reshape(-1), ys.reshape(-1)], axis=-1), extend, axis=0).astype(np.int32)
    index = np.tile(np.arange(extend, dtype=np.uint32), size * size)

    radiance = oracle.evaluate_samples(params, pixels, index).astype(np.float64)
    hit = oracle.trace(oracle.spawn_rays(params, pixels, index))["token"] != structs.TOKEN_EMPTY
    assert 0.1 < hit.mean() < 0.6

    assert np.array_equal(radiance[~hit], np.broadcast_to(np.array(RADIANCE, dtype=np.float32).astype(np.float64), radiance[~hit].shape))  # a miss sees L itself
    expected = np.array(RHO) * np.array(RADIANCE)
    mean = radiance[hit].mean(axis=0)
    error = radiance[hit].std(axis=0) / np.sqrt(hit.sum())
    assert np.all(np.abs(mean - expected) < 4 * error + 1e-4), (mean, expected, error)
    assert np.all(error / expected < 0.01)  # thousands of samples: the check has teeth (1 % of rho L at most)


@pytest.mark.parametrize("kind,label", [(structs.MATERIAL_DIELECTRIC, "smooth dielectric"), (structs.MATERIAL_INVISIBLE, "invisible")])
def test_lossless_surfaces_in_a_uniform_environment(kind, label):
    """The white furnace: a surface that absorbs nothing leaves a uniform environment uniform. A smooth Dielectric picks reflection with probability
    F and refraction with 1 - F, each carrying the full throughput (Specular.cs), so EVERY sample through the glass sphere — whatever its chain of
    internal reflections, up to the bounce limit — returns L, not just their mean; an Invisible surface passes rays on unchanged."""
    description = furnace("sphere")
    description.materials = np.concatenate([scenes.material(kind, (1.0, 1.0, 1.0), roughness=(0.0, 0.0), ior=1.5)])
    oracle = oracle_lib.OracleScene(host.prepare(description))
    size, extend = 24, 8
    params = structs.render_params(size, size, 8, extend=extend, seed=8, bounce_limit=128)
    ys, xs = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    pixels = np.repeat(np.stack([xs.reshape(-1), ys.reshape(-1)], axis=-1), extend, axis=0).astype(np.int32)
    index = np.tile(np.arange(extend, dtype=np.uint32), size * size)

    radiance = oracle.evaluate_samples(params, pixels, index).astype(np.float64)
    hit = oracle.trace(oracle.spawn_rays(params, pixels, index))["token"] != structs.TOKEN_EMPTY
    assert hit.sum() > 1000
    assert np.allclose(radiance[hit], np.array(RADIANCE), rtol=2e-5, atol=0), label
    assert np.allclose(radiance[~hit], np.array(RADIANCE), rtol=1e-7, atol=0)
