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


def lit_plane(lights):
    """a Lambertian plane (it cannot see itself: no indirect light), point lights above it, nothing else"""
    points = np.zeros(len(lights), dtype=structs.POINT_LIGHT)
    for k, (intensity, where) in enumerate(lights):
        points["intensity"][k], points["position"][k] = intensity, where
    position = (0.0, 6.0, -7.0)
    camera = scenes.perspective_camera(position, scenes.look_rotation(position, (0, 0, 0)), field_of_view=50.0, lens_radius=0.0)
    return SceneDescription(triangles=scenes.plane(0, (20, 20)), materials=np.concatenate([scenes.material(structs.MATERIAL_DIFFUSE, RHO)]), point_lights=points, camera=camera)


def plane_samples(description, size, extend, seed):
    oracle = oracle_lib.OracleScene(host.prepare(description))
    params = structs.render_params(size, size, 8, extend=extend, seed=seed, bounce_limit=16)
    ys, xs = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    pixels = np.repeat(np.stack([xs.reshape(-1), ys.reshape(-1)], axis=-1), extend, axis=0).astype(np.int32)
    index = np.tile(np.arange(extend, dtype=np.uint32), size * size)
    radiance = oracle.evaluate_samples(params, pixels, index).astype(np.float64)
    rays = oracle.spawn_rays(params, pixels, index)
    hits = oracle.trace(rays)
    hit = hits["token"] != structs.TOKEN_EMPTY
    points = rays["origin"].astype(np.float64) + rays["direction"].astype(np.float64) * np.where(hit, hits["distance"], 0.0).astype(np.float64)[:, None]
    return radiance, hit, points


def point_light_radiance(points, intensity, position):
    """rho / pi * I cos(theta) / r^2 on the plane y = 0 with normal +Y: PreparedPointLight.Sample hands back intensity / distance^2 (PointLight.cs),
    the evaluator multiplies by the Lambertian BSDF rho / pi and |cos| and divides by the pick probability"""
    offset = np.asarray(position) - points
    squared = (offset ** 2).sum(axis=1)
    return np.array(RHO) / np.pi * np.asarray(intensity) * (offset[:, 1] / np.sqrt(squared) / squared)[:, None]


def test_one_point_light_over_a_plane_is_exact_per_sample():
    """One delta light: the light tree is its leaf (pick probability 1), no MIS partner, the bounce off the plane escapes into black — so every
    sample IS the closed form at its own hit point, to rounding."""
    light = ((30.0, 20.0, 10.0), (1.0, 3.0, -0.5))
    radiance, hit, points = plane_samples(lit_plane([light]), 16, 2, seed=8)
    assert hit.mean() > 0.8 and np.all(radiance[~hit] == 0)
    assert np.allclose(radiance[hit], point_light_radiance(points[hit], *light), rtol=5e-6, atol=0)


def test_two_point_lights_over_a_plane_converge_to_their_sum():
    """Two lights of different power: LightTree.Pick chooses one per sample by LightBound.Importance and divides by that probability — each sample is
    one light's share scaled up, the mean over many samples of a pixel neighbourhood the sum of both closed forms."""
    lights = [((30.0, 20.0, 10.0), (1.5, 3.0, -0.5)), ((4.0, 8.0, 16.0), (-3.0, 1.5, 1.0))]
    radiance, hit, points = plane_samples(lit_plane(lights), 12, 256, seed=9)
    expected = sum(point_light_radiance(points[hit], *light) for light in lights)
    mean, truth = radiance[hit].mean(axis=0), expected.mean(axis=0)
    error = radiance[hit].std(axis=0) / np.sqrt(hit.sum())
    assert np.all(np.abs(mean - truth) < 4 * error), (mean, truth, error)
    assert np.all(error / truth < 0.01)
    only_first = point_light_radiance(points[hit], *lights[0]).mean(axis=0)
    assert np.all(np.abs(mean - only_first) > 10 * error)  # and the second light is in there


def polygon_irradiance(points, vertices):
    """Lambert's formula: the irradiance on a surface with normal +Y at `points` from a polygon of unit radiance, sum over the edges of the angle an
    edge subtends times the cosine between the surface normal and the normal of the plane through the point and the edge, halved"""
    total = np.zeros(len(points))
    for a, b in zip(vertices, np.roll(vertices, -1, axis=0)):
        va, vb = a - points, b - points
        va, vb = va / np.linalg.norm(va, axis=1, keepdims=True), vb / np.linalg.norm(vb, axis=1, keepdims=True)
        gamma = np.arccos(np.clip((va * vb).sum(axis=1), -1.0, 1.0))
        normal = np.cross(va, vb)
        total += gamma * normal[:, 1] / np.linalg.norm(normal, axis=1)
    return np.abs(total) / 2.0


def test_emissive_triangle_over_a_plane_converges_to_lamberts_formula():
    """An area light: a one-sided Emissive triangle (Emissive.Emit: the side its normal faces, Emissive.cs:63) facing a Lambertian plane. The plane's
    radiance is rho / pi times the irradiance Lambert's polygon formula gives in closed form; the evaluator gets there by sampling the triangle's
    area (PreparedTriangle.Sample / ProbabilityDensity), by BSDF samples that happen to reach it, and by the power heuristic between the two."""
    emission = np.array((10.0, 8.0, 6.0))
    corner, side = np.array((-1.0, 3.0, -1.0)), 2.0
    vertices = np.stack([corner, corner + (side, 0, 0), corner + (0, 0, side)])  # edge1 x edge2 = -Y: it faces the plane
    materials = np.concatenate([scenes.material(structs.MATERIAL_DIFFUSE, RHO), scenes.material(structs.MATERIAL_EMISSIVE, tuple(emission))])
    triangles = np.concatenate([scenes.plane(0, (20, 20)), scenes.make_triangles(vertices[0:1], vertices[1:2], vertices[2:3], 1)])
    description = lit_plane([])
    description.triangles, description.materials = triangles, materials

    radiance, hit, points = plane_samples(description, 12, 256, seed=10)
    on_plane = hit & (points[:, 1] < 1e-3)  # the samples that met the plane first (not the light, not the void)
    assert on_plane.mean() > 0.7

    expected = np.array(RHO) / np.pi * emission * polygon_irradiance(points[on_plane], vertices)[:, None]
    mean, truth = radiance[on_plane].mean(axis=0), expected.mean(axis=0)
    error = radiance[on_plane].std(axis=0) / np.sqrt(on_plane.sum())
    assert np.all(np.abs(mean - truth) < 4 * error), (mean, truth, error)
    assert np.all(error / truth < 0.01)


def test_emissive_sphere_over_a_plane_converges_to_its_closed_form():
    """A sphere of radiance Le and radius R whose centre is at distance d, wholly above the horizon of the lit point, gives the irradiance
    pi Le (R / d)^2 cos(theta): PreparedSphere.Sample's cone sampling and its pdf (SphereEntity.cs), MIS against the BSDF samples that reach it."""
    emission, centre, radius = np.array((12.0, 9.0, 5.0)), np.array((0.5, 3.0, 0.0)), 0.75
    description = lit_plane([])
    description.materials = np.concatenate([scenes.material(structs.MATERIAL_DIFFUSE, RHO), scenes.material(structs.MATERIAL_EMISSIVE, tuple(emission))])
    description.spheres = np.zeros(1, dtype=structs.SPHERE)
    description.spheres["position"], description.spheres["radius"], description.spheres["material"] = centre, radius, 1

    radiance, hit, points = plane_samples(description, 12, 256, seed=11)
    on_plane = hit & (np.abs(points[:, 1]) < 1e-3)
    assert on_plane.mean() > 0.7

    offset = centre - points[on_plane]
    squared = (offset ** 2).sum(axis=1)
    irradiance = np.pi * radius * radius / squared * offset[:, 1] / np.sqrt(squared)
    expected = np.array(RHO) / np.pi * emission * irradiance[:, None]
    mean, truth = radiance[on_plane].mean(axis=0), expected.mean(axis=0)
    error = radiance[on_plane].std(axis=0) / np.sqrt(on_plane.sum())
    assert np.all(np.abs(mean - truth) < 4 * error), (mean, truth, error)
    assert np.all(error / truth < 0.01)


# ---- absolute BSDF values through the whole pipeline: with ONE delta light and a flat surface a sample's radiance is tint * f(outgoing, incident) *
# I cos(theta) / r^2, nothing else — so radiance / (I cos / r^2) IS Material.Scatter + BSDF.Evaluate at that pair of directions, and can be set
# beside the formula written down from the C# (BxDFTests.cs pins the lobes' sample / evaluate / pdf CONSISTENCY, not their absolute values) ----

def bsdf_through_the_pipeline(plane_material, seed):
    light = ((30.0, 20.0, 10.0), (1.0, 3.0, -0.5))
    description = lit_plane([light])
    description.materials = plane_material
    radiance, hit, points = plane_samples(description, 16, 2, seed=seed)
    offset = np.asarray(light[1]) - points[hit]
    squared = (offset ** 2).sum(axis=1)
    incident = offset / np.sqrt(squared)[:, None]
    received = np.asarray(light[0]) * (incident[:, 1] / squared)[:, None]  # I cos(theta) / r^2 on the plane y = 0
    camera = np.array((0.0, 6.0, -7.0))
    outgoing = camera - points[hit]
    outgoing /= np.linalg.norm(outgoing, axis=1, keepdims=True)
    return radiance[hit] / received, outgoing, incident


def test_oren_nayar_value_through_the_pipeline():
    """Diffuse with a Roughness texture value above zero scatters as OrenNayar (Diffuse.cs): a = 1 / (pi + (pi / 2 - 2 / 3) sigma), b = a sigma,
    f = a + b s with s = o . i - cos_o cos_i, divided by max(cos_o, cos_i) when positive (Lambertian.cs, OrenNayar.Evaluate)."""
    sigma = 0.7
    value, outgoing, incident = bsdf_through_the_pipeline(np.concatenate([scenes.material(structs.MATERIAL_DIFFUSE, RHO, roughness=(sigma, sigma))]), seed=12)
    a = 1.0 / (np.pi + (np.pi / 2.0 - 2.0 / 3.0) * sigma)
    b = a * sigma
    cos_o, cos_i = outgoing[:, 1], incident[:, 1]
    s = (outgoing * incident).sum(axis=1) - cos_o * cos_i
    s = np.where(s > 8e-7, s / np.maximum(cos_o, cos_i), s)
    expected = np.array(RHO) * (a + b * s)[:, None]
    assert len(value) > 300 and np.ptp(s) > 0.3  # both signs of s, a real range of angles
    assert np.allclose(value, expected, rtol=3e-5, atol=0)


def test_rough_dielectric_reflection_value_through_the_pipeline():
    """Dielectric with roughness scatters as GlossyReflection + GlossyTransmission over a TrowbridgeReitzMicrofacet and RealFresnel(1, eta)
    (Dielectric.cs); light from the camera's side meets the reflection lobe only: f = F(o . h) D(h) G(o, i) / (4 cos_o cos_i) with alpha =
    clamp01(0.75 roughness)^2 (IMicrofacet.GetAlpha), D = 1 / (pi alpha^2 (cos^2 + sin^2 / alpha^2)^2), G = 1 / (1 + Lambda(o) + Lambda(i)),
    Lambda = sqrt(1 + alpha^2 tan^2) / 2 - 1 / 2 (IMicrofacet.cs), F the unpolarised Fresnel reflectance with Snell's law (Fresnel.cs)."""
    roughness, eta, tint = 0.6, 1.5, (0.9, 0.8, 0.7)
    value, outgoing, incident = bsdf_through_the_pipeline(np.concatenate([scenes.material(structs.MATERIAL_DIELECTRIC, tint, roughness=(roughness, roughness), ior=eta)]), seed=13)
    alpha = min(1.0, max(0.0, roughness * 0.75)) ** 2
    half = outgoing + incident
    half /= np.linalg.norm(half, axis=1, keepdims=True)
    cos_o, cos_i, cos_h = outgoing[:, 1], incident[:, 1], half[:, 1]

    d = 1.0 / (np.pi * alpha * alpha * (cos_h ** 2 + (1.0 - cos_h ** 2) / (alpha * alpha)) ** 2)
    shadow = lambda c: np.sqrt(1.0 + alpha * alpha * (1.0 - c * c) / (c * c)) / 2.0 - 0.5
    g = 1.0 / (1.0 + shadow(cos_o) + shadow(cos_i))
    cos_oh = (outgoing * half).sum(axis=1)
    cos_t = np.sqrt(1.0 - (1.0 / eta) ** 2 * (1.0 - cos_oh ** 2))
    parallel = (eta * cos_oh - cos_t) / (eta * cos_oh + cos_t)
    perpendicular = (cos_oh - eta * cos_t) / (cos_oh + eta * cos_t)
    fresnel = (parallel ** 2 + perpendicular ** 2) / 2.0

    expected = np.array(tint) * (fresnel * d * g / (4.0 * cos_o * cos_i))[:, None]
    assert len(value) > 300 and expected.max() / expected.min() > 5  # on and off the highlight

    # PathTracedEvaluator draws the bounce FIRST and skips light sampling altogether when that sample is impossible (`!Positive(bounce.scatterPdf)`,
    # PathTracedEvaluator.cs:68): a microfacet normal whose reflection or refraction leaves on the wrong side. Those samples are black in the
    # reference too; every other one carries the closed form.
    lit = value.max(axis=1) > 0
    assert 0.8 < lit.mean() < 1.0
    assert np.allclose(value[lit], expected[lit], rtol=2e-5, atol=0)


def test_rough_conductor_value_through_the_pipeline():
    """Conductor with physical parameters (Artistic off: RefractiveIndex n and Extinction k per channel) and roughness scatters as GlossyReflection over
    the same microfacet with ComplexFresnel(1, n, k) (Conductor.cs): with t = n^2 - k^2 - sin^2, a2b2 = sqrt(t^2 + 4 n^2 k^2), Rs = (a2b2 + cos^2 -
    cos sqrt(2 (a2b2 + t))) / (a2b2 + cos^2 + cos sqrt(2 (a2b2 + t))), Rp = Rs (cos^2 a2b2 + sin^4 - cos sqrt(2 (a2b2 + t)) sin^2) / (... + ...),
    F = (Rs + Rp) / 2 (Fresnel.cs, ComplexFresnel.Evaluate)."""
    roughness, n, k = 0.5, np.array((0.18, 0.42, 1.37)), np.array((3.42, 2.35, 1.77))
    material = scenes.material(structs.MATERIAL_CONDUCTOR, (1.0, 1.0, 1.0), roughness=(roughness, roughness), param_a=tuple(n), param_b=tuple(k), flags=0)
    value, outgoing, incident = bsdf_through_the_pipeline(np.concatenate([material]), seed=14)
    alpha = min(1.0, max(0.0, roughness * 0.75)) ** 2
    half = outgoing + incident
    half /= np.linalg.norm(half, axis=1, keepdims=True)
    cos_o, cos_i, cos_h = outgoing[:, 1], incident[:, 1], half[:, 1]

    d = 1.0 / (np.pi * alpha * alpha * (cos_h ** 2 + (1.0 - cos_h ** 2) / (alpha * alpha)) ** 2)
    shadow = lambda c: np.sqrt(1.0 + alpha * alpha * (1.0 - c * c) / (c * c)) / 2.0 - 0.5
    g = 1.0 / (1.0 + shadow(cos_o) + shadow(cos_i))

    cos = np.clip(np.abs((outgoing * half).sum(axis=1)), 0.0, 1.0)[:, None]
    cos2, sin2 = cos * cos, 1.0 - cos * cos
    term = n * n - k * k - sin2
    a2b2 = np.sqrt(term * term + 4.0 * n * n * k * k)
    para0, para1 = a2b2 + cos2, cos * np.sqrt(2.0) * np.sqrt(a2b2 + term)
    perp0, perp1 = cos2 * a2b2 + sin2 * sin2, para1 * sin2
    para, perp = (para0 - para1) / (para0 + para1), (perp0 - perp1) / (perp0 + perp1)
    fresnel = (para * perp + para) / 2.0

    expected = fresnel * (d * g / (4.0 * cos_o * cos_i))[:, None]
    lit = value.max(axis=1) > 0  # see the dielectric: an impossible bounce sample switches light sampling off for that sample
    assert len(value) > 300 and 0.8 < lit.mean() <= 1.0
    assert np.allclose(value[lit], expected[lit], rtol=3e-5, atol=0)


def test_rough_dielectric_transmission_value_through_the_pipeline():
    """The light BELOW the glass sheet, the camera above: only the GlossyTransmission lobe answers (Glossy.cs, GlossyTransmission.Evaluate). With
    etaR = eta_incident / eta_outgoing = eta for a camera on the outside, h = normalize(o + etaR i) turned to the upper side, f = (1 - F(o . h)) D(h)
    G(o, i) |etaR^2 (o . h)(i . h)| / ((etaR (i . h) + o . h)^2 |cos_o cos_i|), black when o . h and i . h have the same sign."""
    roughness, eta, tint = 0.6, 1.5, (0.9, 0.8, 0.7)
    light = ((30.0, 20.0, 10.0), (0.5, -2.5, 0.5))
    description = lit_plane([light])
    description.materials = np.concatenate([scenes.material(structs.MATERIAL_DIELECTRIC, tint, roughness=(roughness, roughness), ior=eta)])
    radiance, hit, points = plane_samples(description, 16, 2, seed=15)

    offset = np.asarray(light[1]) - points[hit]
    squared = (offset ** 2).sum(axis=1)
    incident = offset / np.sqrt(squared)[:, None]
    received = np.asarray(light[0]) * (np.abs(incident[:, 1]) / squared)[:, None]
    outgoing = np.array((0.0, 6.0, -7.0)) - points[hit]
    outgoing /= np.linalg.norm(outgoing, axis=1, keepdims=True)
    value = radiance[hit] / received

    alpha = min(1.0, max(0.0, roughness * 0.75)) ** 2
    half = outgoing + eta * incident
    half /= np.linalg.norm(half, axis=1, keepdims=True)
    half *= np.sign(half[:, 1])[:, None]
    cos_o, cos_i, cos_h = outgoing[:, 1], incident[:, 1], half[:, 1]
    dot_o, dot_i = (outgoing * half).sum(axis=1), (incident * half).sum(axis=1)

    d = 1.0 / (np.pi * alpha * alpha * (cos_h ** 2 + (1.0 - cos_h ** 2) / (alpha * alpha)) ** 2)
    shadow = lambda c: np.sqrt(1.0 + alpha * alpha * (1.0 - c * c) / (c * c)) / 2.0 - 0.5
    g = 1.0 / (1.0 + shadow(cos_o) + shadow(cos_i))
    cos_t = np.sqrt(np.maximum(0.0, 1.0 - (1.0 / eta) ** 2 * (1.0 - dot_o ** 2)))
    parallel = (eta * dot_o - cos_t) / (eta * dot_o + cos_t)
    perpendicular = (dot_o - eta * cos_t) / (dot_o + eta * cos_t)
    fresnel = (parallel ** 2 + perpendicular ** 2) / 2.0

    f = (1.0 - fresnel) * d * g * np.abs(eta * eta * dot_o * dot_i / ((eta * dot_i + dot_o) ** 2 * cos_o * cos_i))
    f = np.where(dot_o * dot_i > 0, 0.0, f)
    expected = np.array(tint) * f[:, None]
    lit = value.max(axis=1) > 0
    # black where the formula is black (o . h and i . h on the same side: no refraction connects the two directions through that microfacet),
    # and on the few samples whose bounce was impossible (see the reflection lobe above)
    assert len(value) > 300 and lit.sum() > 60 and np.all(expected[lit].min(axis=1) > 0)
    dark_where_the_formula_is_lit = (~lit) & (expected.max(axis=1) > 0)
    assert dark_where_the_formula_is_lit.sum() < 0.25 * (expected.max(axis=1) > 0).sum()
    assert np.allclose(value[lit], expected[lit], rtol=5e-5, atol=0)


def test_coated_diffuse_value_through_the_pipeline():
    """CoatedDiffuse (CoatedDiffuse.cs:37-55) = CoatedLambertianReflection + GlossyReflection over RealFresnel(1, eta), both under the BSDF's albedo tint:
    the coated lobe is eta_r^2 / pi / (1 - albedo R) (1 - F(cos_o)) (1 - F(cos_i)) with eta_r = 1 / eta and R = FresnelDiffuseReflectance(1 / eta)
    (Lambertian.cs:131-199: 1 - eta_r^2 (1 - E(eta)), E the closed form with the two logarithms), the coat the GGX reflection checked above."""
    roughness, eta, albedo = 0.5, 1.5, np.array((0.8, 0.5, 0.3))
    value, outgoing, incident = bsdf_through_the_pipeline(scenes.coated_diffuse(tuple(albedo), roughness=(roughness, roughness), ior=eta), seed=16)

    def fresnel(cos):  # RealFresnel(1, eta).Evaluate from the outside
        cos_t = np.sqrt(1.0 - (1.0 / eta) ** 2 * (1.0 - cos ** 2))
        parallel = (eta * cos - cos_t) / (eta * cos + cos_t)
        perpendicular = (cos - eta * cos_t) / (cos + eta * cos_t)
        return (parallel ** 2 + perpendicular ** 2) / 2.0

    def entrance(e):  # CoatedLambertianReflection.FresnelDiffuseReflectance's EntranceReflectance, e >= 1
        e2, e4 = e * e, e ** 4
        q0 = (e - 1) * (3 * e + 1) / (6 * (e + 1) ** 2)
        q1 = e2 * (e2 - 1) ** 2 / (e2 + 1) ** 3
        q2 = -2 * e2 * e * (e2 + e + e - 1) / ((e2 + 1) * (e4 - 1))
        q3 = 8 * e4 * (e4 + 1) / ((e2 + 1) * (e4 - 1) ** 2)
        return 0.5 + (np.log((e - 1) / (e + 1)) * q1 + q0) + (np.log(e) * q3 + q2)

    eta_r = 1.0 / eta
    reflectance = 1.0 - eta_r * eta_r * (1.0 - entrance(eta))
    cos_o, cos_i = outgoing[:, 1], incident[:, 1]
    coated = (eta_r * eta_r / np.pi / (1.0 - albedo * reflectance)) * ((1.0 - fresnel(cos_o)) * (1.0 - fresnel(cos_i)))[:, None]

    alpha = min(1.0, max(0.0, roughness * 0.75)) ** 2
    half = outgoing + incident
    half /= np.linalg.norm(half, axis=1, keepdims=True)
    cos_h = half[:, 1]
    d = 1.0 / (np.pi * alpha * alpha * (cos_h ** 2 + (1.0 - cos_h ** 2) / (alpha * alpha)) ** 2)
    shadow = lambda c: np.sqrt(1.0 + alpha * alpha * (1.0 - c * c) / (c * c)) / 2.0 - 0.5
    coat = fresnel((outgoing * half).sum(axis=1)) * d / (1.0 + shadow(cos_o) + shadow(cos_i)) / (4.0 * cos_o * cos_i)

    expected = albedo * (coated + coat[:, None])
    lit = value.max(axis=1) > 0
    assert len(value) > 300 and 0.8 < lit.mean() <= 1.0 and 0.3 < reflectance < 0.7
    assert np.allclose(value[lit], expected[lit], rtol=5e-5, atol=0)


def test_ambient_and_point_light_together_carry_the_references_excess():
    """PreparedScene.Pick first chooses between the infinite lights and the light tree by the power threshold t (PreparedScene.cs:279-325), then within;
    on a plane (no interreflection) an unbiased estimator would return the sum of both closed forms, rho L + rho / pi I cos / r^2. The reference
    returns MORE, and exactly how much follows from its code: when the picked light is a delta light, `mis = light.TopToken.IsAreaLight()` is false and
    the bounce that escapes adds the WHOLE environment (PathTracedEvaluator.cs:136-143, "Fallback with no MIS") — although the environment's light-sampling
    share is also collected, MIS-weighted, in the samples that picked the environment. Per sample the environment part is therefore counted
    (1 - t) times too often by its light-sampling share: with the environment sampled uniformly over the sphere (pdf 1 / 4 pi, IDirectionalTexture.cs)
    against the cosine lobe (cos / pi) under the power heuristic, that share of rho L is 2 * Integral_0^1 c a^2 / (a^2 + c^2 / pi^2) dc = (a pi)^2 ln(1 +
    1 / (a pi)^2) with a = t / (4 pi). The oracle keeps the reference's behaviour, and must land on the sum PLUS that excess."""
    light = ((30.0, 20.0, 10.0), (1.0, 3.0, -0.5))
    description = lit_plane([light])
    description.infinite_lights = scenes.ambient_light(RADIANCE)
    prepared = host.prepare(description)
    radiance, hit, points = plane_samples(description, 12, 256, seed=17)

    threshold = float(prepared.infinite_threshold)
    a_pi = threshold / 4.0
    excess = (1.0 - threshold) * a_pi * a_pi * np.log(1.0 + 1.0 / (a_pi * a_pi)) * np.array(RHO) * np.array(RADIANCE)
    unbiased = (point_light_radiance(points[hit], *light) + np.array(RHO) * np.array(RADIANCE)).mean(axis=0)
    mean = radiance[hit].mean(axis=0)
    error = radiance[hit].std(axis=0) / np.sqrt(hit.sum())
    assert 0.3 < threshold < 0.7 and np.all(excess > 10 * error)             # the excess is far outside the noise ...
    assert np.all(np.abs(mean - (unbiased + excess)) < 4 * error), (mean, unbiased, excess, error)  # ... and it is the one the code implies
    assert np.all(mean - unbiased > 10 * error)


def test_sky_and_an_area_light_together_converge_to_their_sum():
    """Two non-delta lights, so every pick keeps MIS and nothing is counted twice: a plane under a uniform sky L and a one-sided emissive triangle Le
    facing it. The triangle also hides its part of the sky, and that part is the same projected solid angle Lambert's formula gives:
    radiance = rho / pi (L (pi - E) + Le E) with E the polygon irradiance for unit radiance."""
    emission = np.array((10.0, 8.0, 6.0))
    corner, side = np.array((-1.0, 3.0, -1.0)), 2.0
    vertices = np.stack([corner, corner + (side, 0, 0), corner + (0, 0, side)])
    description = lit_plane([])
    description.materials = np.concatenate([scenes.material(structs.MATERIAL_DIFFUSE, RHO), scenes.material(structs.MATERIAL_EMISSIVE, tuple(emission))])
    description.triangles = np.concatenate([scenes.plane(0, (20, 20)), scenes.make_triangles(vertices[0:1], vertices[1:2], vertices[2:3], 1)])
    description.infinite_lights = scenes.ambient_light(RADIANCE)

    radiance, hit, points = plane_samples(description, 12, 256, seed=18)
    on_plane = hit & (points[:, 1] < 1e-3)
    covered = polygon_irradiance(points[on_plane], vertices)[:, None]
    expected = np.array(RHO) / np.pi * (np.array(RADIANCE) * (np.pi - covered) + emission * covered)
    mean, truth = radiance[on_plane].mean(axis=0), expected.mean(axis=0)
    error = radiance[on_plane].std(axis=0) / np.sqrt(on_plane.sum())
    assert np.all(np.abs(mean - truth) < 4 * error), (mean, truth, error)
    assert np.all(error / truth < 0.01)


# ---- specular surfaces: a flat sheet under a sky that is bright above the horizon and black below it. A ray reflected off the sheet sees the
# sky, a refracted one the black below: the mean radiance of the samples of a view direction is the Fresnel reflectance times the sky — the
# probability SpecularFresnel picks reflection with (Specular.cs), or the weight SpecularReflection<ComplexFresnel> carries. ----

def half_sky(material):
    from echorenderer_b200 import TextureDescription
    texels = np.zeros((32, 64, 4), dtype=np.float32)
    texels[16:, :, :3] = 1.0  # rows grow upward (v = 0 looks down): the upper hemisphere is white
    texels[..., 3] = 1.0
    description = lit_plane([])
    description.materials = material
    description.textures = [TextureDescription(texels)]
    description.infinite_lights = scenes.environment_light(0, intensity=RADIANCE)
    return description


def view_directions(points, hit):
    outgoing = np.array((0.0, 6.0, -7.0)) - points[hit]
    return outgoing / np.linalg.norm(outgoing, axis=1, keepdims=True)


def test_smooth_dielectric_sheet_reflects_the_fresnel_share_of_the_sky():
    eta = 1.5
    radiance, hit, points = plane_samples(half_sky(np.concatenate([scenes.material(structs.MATERIAL_DIELECTRIC, (1.0, 1.0, 1.0), ior=eta)])), 12, 256, seed=19)
    cos = view_directions(points, hit)[:, 1]
    cos_t = np.sqrt(1.0 - (1.0 / eta) ** 2 * (1.0 - cos ** 2))
    fresnel = (((eta * cos - cos_t) / (eta * cos + cos_t)) ** 2 + ((cos - eta * cos_t) / (cos + eta * cos_t)) ** 2) / 2.0

    values = radiance[hit] / np.array(RADIANCE)
    assert np.all((np.abs(values - 1.0) < 1e-5) | (values == 0.0))  # every sample is the sky or the dark: the choice is the whole estimator
    mean, truth = values[:, 0].mean(), fresnel.mean()
    error = values[:, 0].std() / np.sqrt(hit.sum())
    assert abs(mean - truth) < 4 * error and error / truth < 0.02 and 0.03 < truth < 0.2, (mean, truth, error)


def test_smooth_conductor_sheet_reflects_its_complex_fresnel_reflectance():
    """Physical parameters (n, k) and the artistic ones (Conductor.cs:66-96, after Gulbrandsen 2014: main colour r clamped below 1, edge colour g;
    n = g (1 - r) / (1 + r) + (1 - g) (1 + sqrt r) / (1 - sqrt r), k = sqrt(max(0, (r (n + 1)^2 - (n - 1)^2) / (1 - r)))) through the same mirror:
    every sample carries ComplexFresnel(1, n, k) at the view angle."""
    def reflectance(n, k, cos):
        cos = cos[:, None]
        cos2, sin2 = cos * cos, 1.0 - cos * cos
        term = n * n - k * k - sin2
        a2b2 = np.sqrt(term * term + 4.0 * n * n * k * k)
        para0, para1 = a2b2 + cos2, cos * np.sqrt(2.0) * np.sqrt(a2b2 + term)
        perp0, perp1 = cos2 * a2b2 + sin2 * sin2, para1 * sin2
        para, perp = (para0 - para1) / (para0 + para1), (perp0 - perp1) / (perp0 + perp1)
        return (para * perp + para) / 2.0

    n, k = np.array((0.18, 0.42, 1.37)), np.array((3.42, 2.35, 1.77))
    physical = scenes.material(structs.MATERIAL_CONDUCTOR, (1.0, 1.0, 1.0), param_a=tuple(n), param_b=tuple(k), flags=0)
    radiance, hit, points = plane_samples(half_sky(np.concatenate([physical])), 12, 4, seed=20)
    expected = reflectance(n, k, view_directions(points, hit)[:, 1]) * np.array(RADIANCE)
    assert hit.sum() > 400 and np.allclose(radiance[hit], expected, rtol=3e-5, atol=0)

    main, edge = np.array((0.6, 0.7, 0.9)), np.array((0.0, 1.0, 0.5))  # the C3 conductor (bunny.echo)
    artistic = scenes.material(structs.MATERIAL_CONDUCTOR, (1.0, 1.0, 1.0), param_a=tuple(main), param_b=tuple(edge), flags=structs.MATERIAL_FLAG_ARTISTIC)
    radiance, hit, points = plane_samples(half_sky(np.concatenate([artistic])), 12, 4, seed=21)
    root = np.sqrt(main)
    n = edge * (1.0 - main) / (1.0 + main) + (1.0 - edge) * (1.0 + root) / (1.0 - root)
    k = np.sqrt(np.maximum(0.0, (main * (n + 1.0) ** 2 - (n - 1.0) ** 2) / (1.0 - main)))
    expected = reflectance(n, k, view_directions(points, hit)[:, 1]) * np.array(RADIANCE)
    assert hit.sum() > 400 and np.allclose(radiance[hit], expected, rtol=5e-5, atol=0)


@pytest.mark.parametrize("backface", [True, False])
def test_one_sided_sheet_under_a_uniform_sky(backface):
    """OneSided (OneSided.cs): `Cull = Positive(outgoing . normal) != Backface` — with Backface (the default) the sheet scatters as its Base when seen
    from the side its normal points to and is Invisible from behind; without it the other way round. Under a uniform sky L a Lambertian base
    makes the scattering side rho L (the furnace again: a plane is convex) and the culled side L, sample for sample."""
    description = lit_plane([])
    description.materials = np.concatenate([scenes.material(structs.MATERIAL_DIFFUSE, RHO),
                                            scenes.material(structs.MATERIAL_ONESIDED, flags=structs.MATERIAL_FLAG_BACKFACE if backface else 0, base=0)])
    description.triangles["material"] = 1
    description.infinite_lights = scenes.ambient_light(RADIANCE)

    for camera_height in (6.0, -6.0):  # above the sheet (its normal is +Y), then below it
        position = (0.0, camera_height, -7.0)
        description.camera = scenes.perspective_camera(position, scenes.look_rotation(position, (0, 0, 0)), field_of_view=50.0, lens_radius=0.0)
        radiance, hit, _ = plane_samples(description, 12, 64, seed=22)
        scatters = (camera_height > 0) == backface
        assert hit.mean() > 0.8
        if scatters:
            mean = radiance[hit].mean(axis=0)
            error = radiance[hit].std(axis=0) / np.sqrt(hit.sum())
            assert np.all(np.abs(mean - np.array(RHO) * np.array(RADIANCE)) < 4 * error + 1e-4) and np.all(error < 0.01)
        else:
            assert np.allclose(radiance[hit], np.array(RADIANCE), rtol=1e-6, atol=0)


def test_transmissive_diffuse_sheet_and_a_visible_emitter():
    """Diffuse with Transmissive scatters as the two-sided Lambertian (Diffuse.cs, Lambertian.cs: f = 1 / 2 pi into both hemispheres): under a uniform
    sky the sheet returns rho L seen from either side. And an Emissive surface shows its emission to the side its normal faces and nothing of its
    own to the other (Emissive.Emit): there the path goes on with the black BSDF and ends — the sample is black even under a sky."""
    description = lit_plane([])
    description.materials = np.concatenate([scenes.material(structs.MATERIAL_DIFFUSE, RHO, flags=structs.MATERIAL_FLAG_TRANSMISSIVE)])
    description.infinite_lights = scenes.ambient_light(RADIANCE)
    for camera_height in (6.0, -6.0):
        position = (0.0, camera_height, -7.0)
        description.camera = scenes.perspective_camera(position, scenes.look_rotation(position, (0, 0, 0)), field_of_view=50.0, lens_radius=0.0)
        radiance, hit, _ = plane_samples(description, 12, 128, seed=23)
        mean = radiance[hit].mean(axis=0)
        error = radiance[hit].std(axis=0) / np.sqrt(hit.sum())
        assert np.all(np.abs(mean - np.array(RHO) * np.array(RADIANCE)) < 4 * error) and np.all(error / mean < 0.01), (mean, error)

    emission = (7.0, 5.0, 3.0)
    description.materials = np.concatenate([scenes.material(structs.MATERIAL_EMISSIVE, emission)])
    for camera_height, expected in ((6.0, emission), (-6.0, (0.0, 0.0, 0.0))):
        position = (0.0, camera_height, -7.0)
        description.camera = scenes.perspective_camera(position, scenes.look_rotation(position, (0, 0, 0)), field_of_view=50.0, lens_radius=0.0)
        radiance, hit, _ = plane_samples(description, 12, 4, seed=24)
        assert hit.sum() > 400 and np.allclose(radiance[hit], np.array(expected, dtype=np.float32).astype(np.float64), rtol=1e-6, atol=0)
