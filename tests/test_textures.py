"""Image textures on the oracle: TextureGrid filters and wrappers (Textures/Grids/IFilter.cs, IWrapper.cs),
the textured material slots (Material.Sample, Material.cs:100-102), normal mapping (Material.cs:77-98), the pinned inverse
trigonometry behind sphere texture coordinates (SphereEntity.cs:236-245)."""
import numpy as np
import pytest

from echorenderer_b200 import host, scenes, structs

from . import oracle_lib


@pytest.fixture(scope="module")
def textured():
    prepared = host.prepare(scenes.textured_scene())
    return prepared, oracle_lib.OracleScene(prepared)


def sample(oracle, texture, uv):
    uv = np.ascontiguousarray(uv, dtype=np.float32)
    out = np.zeros((len(uv), 4), dtype=np.float32)
    oracle.lib.oracle_texture_sample(oracle.handle, texture, oracle_lib.ptr(uv), len(uv), oracle_lib.ptr(out))
    return out


def wrap(index, size, wrapper):
    if wrapper == structs.WRAPPER_CLAMP:
        return np.clip(index, 0, size - 1)
    if wrapper == structs.WRAPPER_REPEAT:
        return np.mod(index, size)
    folded = np.mod(index, 2 * size)
    return np.minimum(folded, 2 * size - 1 - folded)


def test_pinned_inverse_trigonometry():
    lib = oracle_lib.library()
    rng = np.random.default_rng(3)
    values = np.concatenate([rng.uniform(-1, 1, 4000), [-1.0, 1.0, 0.0, 0.5, -0.5, 1e-5, -1e-5]]).astype(np.float32)
    asin = np.array([lib.oracle_asin(float(v)) for v in values])
    acos = np.array([lib.oracle_acos(float(v)) for v in values])
    assert np.abs(asin - np.arcsin(values.astype(np.float64))).max() < 4e-7
    assert np.abs(acos - np.arccos(values.astype(np.float64))).max() < 6e-7

    y, x = rng.normal(size=4000).astype(np.float32), rng.normal(size=4000).astype(np.float32)
    atan2 = np.array([lib.oracle_atan2(float(a), float(b)) for a, b in zip(y, x)])
    assert np.abs(atan2 - np.arctan2(y.astype(np.float64), x.astype(np.float64))).max() < 6e-7
    assert lib.oracle_atan2(0.0, 0.0) == 0.0 and lib.oracle_atan2(0.0, -1.0) == pytest.approx(np.pi)
    assert lib.oracle_atan2(1.0, 0.0) == pytest.approx(np.pi / 2) and lib.oracle_atan2(-1.0, 0.0) == pytest.approx(-np.pi / 2)


@pytest.mark.parametrize("texture", range(7))
def test_filters_and_wrappers(textured, texture):
    """Against an independent float64 restatement: point = the texel the coordinate falls in; bilinear = the usual half-texel
    interpolation; both through clamp / repeat / mirror, inside and far outside [0, 1]."""
    prepared, oracle = textured
    record = prepared.textures[texture]
    width, height = int(record["width"]), int(record["height"])
    texels = prepared.texels[record["texelOffset"]:record["texelOffset"] + width * height].reshape(height, width, 4).astype(np.float64)
    wrapper = int(record["wrapper"])

    rng = np.random.default_rng(texture)
    uv = np.concatenate([rng.uniform(0, 1, (3000, 2)), rng.uniform(-3.5, 4.5, (3000, 2))]).astype(np.float32)
    value = sample(oracle, texture, uv)
    scaled = uv.astype(np.float64) * [width, height]

    if record["filter"] == structs.FILTER_POINT:
        x, y = wrap(np.floor(scaled[:, 0]).astype(int), width, wrapper), wrap(np.floor(scaled[:, 1]).astype(int), height, wrapper)
        # float32 products may land on the other side of a texel border than the float64 ones for a handful of coordinates
        assert np.mean(np.all(value == texels[y, x].astype(np.float32), axis=1)) > 0.999
        return

    base = np.floor(scaled - 0.5).astype(int)
    time = scaled - 0.5 - base
    x0, x1 = wrap(base[:, 0], width, wrapper), wrap(base[:, 0] + 1, width, wrapper)
    y0, y1 = wrap(base[:, 1], height, wrapper), wrap(base[:, 1] + 1, height, wrapper)
    low = texels[y0, x0] * (1 - time[:, :1]) + texels[y0, x1] * time[:, :1]
    high = texels[y1, x0] * (1 - time[:, :1]) + texels[y1, x1] * time[:, :1]
    expected = low * (1 - time[:, 1:]) + high * time[:, 1:]
    assert np.abs(value - expected).max() < 2e-5  # float32 coordinates: the weights carry ~1e-6 * |uv * size|

    # texel centres reproduce the texel exactly, whichever way cvtps2dq rounds the half (IFilter.cs:41-63)
    ys, xs = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    centres = np.stack([(xs.reshape(-1) + 0.5) / width, (ys.reshape(-1) + 0.5) / height], axis=-1).astype(np.float32)
    exact = (centres[:, 0] * np.float32(width) == xs.reshape(-1) + np.float32(0.5)) & (centres[:, 1] * np.float32(height) == ys.reshape(-1) + np.float32(0.5))
    assert exact.sum() > 10
    assert np.array_equal(sample(oracle, texture, centres)[exact], texels.reshape(-1, 4)[exact].astype(np.float32))


def albedo_pass(oracle, width, height, evaluator=structs.EVALUATOR_ALBEDO | structs.EVALUATOR_DIVERGE_ONCE):
    params = structs.render_params(width, height, 16, extend=1, seed=2, evaluator=evaluator)
    ys, xs = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    pixels = np.stack([xs.reshape(-1), ys.reshape(-1)], axis=-1).astype(np.int32)
    index = np.zeros(len(pixels), dtype=np.uint32)
    out = np.zeros((len(pixels), 4), dtype=np.float32)
    oracle.lib.oracle_evaluate_samples4(oracle.handle, oracle_lib.ptr(params), oracle_lib.ptr(pixels), oracle_lib.ptr(index), len(index), oracle_lib.ptr(out), 4, 0)
    rays = oracle.spawn_rays(params, pixels, index)
    return out, rays, oracle.trace(rays)


def test_albedo_follows_the_texture_coordinates(textured):
    """Ground: planar texture coordinates 0.5 per unit over a repeating 8 x 8 checker, point filter — the albedo pass must
    show exactly the checker texel under every hit point (PreparedTriangle.GetTexcoord, TriangleEntity.cs:188)."""
    prepared, oracle = textured
    value, rays, hits = albedo_pass(oracle, 96, 54)
    ground = (hits["token"] != structs.TOKEN_EMPTY) & (structs.token_type(hits["token"]) == structs.TOKEN_TYPE_TRIANGLE)
    ground &= prepared.triangles["material"][np.where(ground, structs.token_index(hits["token"]), 0)] == 0
    assert ground.sum() > 500

    position = rays["origin"][ground].astype(np.float64) + rays["direction"][ground].astype(np.float64) * hits["distance"][ground][:, None]
    cell = np.floor(position[:, [0, 2]] * 0.5 * 8).astype(int) % 8
    checker = prepared.texels[:64].reshape(8, 8, 4)
    expected = checker[cell[:, 1], cell[:, 0]][:, :3]
    border = np.abs(position[:, [0, 2]] * 4 - np.round(position[:, [0, 2]] * 4)).min(axis=1) < 1e-3  # hits on a texel border may go either way
    assert np.array_equal(value[ground][~border][:, :3], expected[~border])


def test_sphere_texture_coordinates_and_alpha_cutout(textured):
    prepared, oracle = textured
    value, rays, hits = albedo_pass(oracle, 96, 54)
    sphere = (hits["token"] != structs.TOKEN_EMPTY) & (structs.token_type(hits["token"]) == structs.TOKEN_TYPE_SPHERE)
    assert sphere.sum() > 50
    # stripes: R is 0.9 everywhere, G varies with the latitude
    assert np.allclose(value[sphere][:, 0], 0.9, atol=1e-6) and value[sphere][:, 1].std() > 0.05

    quad = (hits["token"] != structs.TOKEN_EMPTY) & (structs.token_type(hits["token"]) == structs.TOKEN_TYPE_TRIANGLE)
    quad &= prepared.triangles["material"][np.where(quad, structs.token_index(hits["token"]), 0)] == 5
    assert quad.sum() > 20
    record = prepared.textures[6]
    cutout = prepared.texels[record["texelOffset"]:record["texelOffset"] + 256]
    opaque = cutout[cutout[:, 3] >= 0.5][:, :3]
    shown = value[quad][:, :3]
    is_cutout_colour = (np.abs(shown[:, None, :] - opaque[None]).max(axis=2) == 0).any(axis=1)
    assert 0.2 < is_cutout_colour.mean() < 0.9  # the rest looks through the holes (alpha < 0.5 -> Invisible, Material.cs:67-72)


def test_normal_mapping(textured):
    prepared, oracle = textured
    mapped, _, hits = albedo_pass(oracle, 96, 54, structs.EVALUATOR_NORMAL_DEPTH)
    flat_prepared = host.prepare(scenes.textured_scene())
    flat_prepared.material_textures["normalIntensity"][:] = 0.0  # zeroNormal, Material.cs:58
    flat, _, _ = albedo_pass(oracle_lib.OracleScene(flat_prepared), 96, 54, structs.EVALUATOR_NORMAL_DEPTH)

    triangle = (hits["token"] != structs.TOKEN_EMPTY) & (structs.token_type(hits["token"]) == structs.TOKEN_TYPE_TRIANGLE)
    material = np.where(triangle, prepared.triangles["material"][np.where(triangle, structs.token_index(hits["token"]), 0)], -1)
    marble = material == 1
    assert marble.sum() > 100
    assert np.allclose(np.linalg.norm(mapped[marble][:, :3], axis=1), 1.0, atol=1e-5)
    assert np.abs(mapped[marble][:, :3] - flat[marble][:, :3]).max(axis=1).mean() > 0.02  # the map bends the normals
    assert np.array_equal(mapped[material == 0], flat[material == 0])                     # only where a normal map is bound
    assert np.array_equal(mapped[:, 3], flat[:, 3])                                       # depth is untouched


def infinite_light(oracle, index, sample, direction):
    out = np.zeros(11, dtype=np.float32)
    sample, direction = np.ascontiguousarray(sample, dtype=np.float32), np.ascontiguousarray(direction, dtype=np.float32)
    oracle.lib.oracle_infinite_light(oracle.handle, index, oracle_lib.ptr(sample), oracle_lib.ptr(direction), oracle_lib.ptr(out))
    return out


def test_environment_light_sampling_is_consistent(environment_small):
    """CylindricalTexture (CylindricalTexture.cs:98-140): Sample, Evaluate and ProbabilityDensity agree on the same direction,
    the density integrates to one over the sphere, and the sun texels are drawn far more often than their solid angle."""
    oracle = oracle_lib.OracleScene(environment_small)
    rng = np.random.default_rng(9)
    inverse, ratios, bright = [], [], 0

    for _ in range(3000):
        out = infinite_light(oracle, 0, rng.uniform(0, 1, 2), (0, 1, 0))
        radiance, pdf, incident = out[0:3], out[3], out[4:7]
        assert pdf > 0 and np.linalg.norm(incident) == pytest.approx(1.0, abs=1e-5)
        again = infinite_light(oracle, 0, (0.5, 0.5), incident)
        assert again[10] == pytest.approx(pdf, rel=2e-3)           # ProbabilityDensity(Sample().incident) == Sample().pdf
        assert np.allclose(again[7:10], radiance, rtol=2e-2, atol=1e-3)  # Evaluate(incident) == sampled radiance (bilinear, re-derived uv)
        inverse.append(1.0 / pdf)
        bright += radiance[0] > 5.0

    assert np.mean(inverse) == pytest.approx(4 * np.pi, rel=0.05)  # E[1 / pdf] = the sphere's solid angle
    assert bright / 3000 > 0.15                                       # the sun covers ~0.3 % of the map


def test_environment_light_irradiance():
    """A lone Lambertian plane under the sky: albedo / pi * the cosine-weighted integral of the map over the upper hemisphere,
    computed by quadrature with the texture's own direction convention (CylindricalTexture.ToDirection, :153-164)."""
    from tests.conftest import sky_texture
    albedo = 0.7
    texture = host.TextureDescription(sky_texture(), structs.FILTER_BILINEAR, structs.WRAPPER_REPEAT)
    description = host.SceneDescription(triangles=scenes.plane(0, (400, 400)), materials=scenes.material(structs.MATERIAL_DIFFUSE, (albedo,) * 3),
                                        textures=[texture], infinite_lights=scenes.environment_light(0, (1, 1, 1), directly_visible=False))
    position = (0.0, 4.0, -2.0)
    description.camera = scenes.perspective_camera(position, scenes.look_rotation(position, (0, 0, 0)), field_of_view=20.0)
    prepared = host.prepare(description)
    image, _ = oracle_lib.OracleScene(prepared).render_tiles(structs.render_params(32, 32, 16, extend=512, seed=3, bounce_limit=3), scenes.tile_grid(32, 32, 16))

    n = 1024
    v, u = np.meshgrid((np.arange(n // 2) + 0.5) / (n // 2), (np.arange(n) + 0.5) / n, indexing="ij")
    radiance = host.sample_texture(texture, np.stack([u, v], axis=-1))[..., :3]
    phi = v * np.pi
    cosine = np.maximum(-np.cos(phi), 0.0)  # direction.y = -cos(phi): the upper hemisphere is v > 0.5
    solid_angle = np.sin(phi) * (np.pi / (n // 2)) * (2 * np.pi / n)
    expected = albedo / np.pi * (radiance * (cosine * solid_angle)[..., None]).sum(axis=(0, 1))
    assert np.allclose(image[..., :3].mean(axis=(0, 1, 2)), expected, rtol=0.03)


def test_cubemap_faces(cubemap_small):
    """Cubemap.Evaluate (Cubemap.cs:62-82): the face is the major axis of the direction, px nx py ny pz nz; sampling is
    uniform over the sphere (IDirectionalTexture's defaults)."""
    oracle = oracle_lib.OracleScene(cubemap_small)
    light = cubemap_small.description.infinite_lights[0]
    to_world = light["rotation"].reshape(3, 3).astype(np.float64)
    colours = np.array([(1.0, 0.3, 0.2), (0.2, 1.0, 0.3), (0.5, 0.7, 1.6), (0.4, 0.3, 0.2), (0.9, 0.9, 0.3), (0.8, 0.3, 0.9)])

    for face, local in enumerate([(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]):
        world = to_world @ np.array(local, dtype=np.float64)
        out = infinite_light(oracle, 0, (0.5, 0.5), world / np.linalg.norm(world))
        ratio = out[7:10] / (colours[face] * 0.8)
        assert np.allclose(ratio, ratio[0], rtol=1e-4) and 0.3 < ratio[0] < 1.3  # the face's colour, shaded by its pattern
        assert out[10] == pytest.approx(1 / (4 * np.pi), rel=1e-6)

    rng = np.random.default_rng(4)
    for _ in range(200):
        out = infinite_light(oracle, 0, rng.uniform(0, 1, 2), (0, 1, 0))
        again = infinite_light(oracle, 0, (0.5, 0.5), out[4:7])
        assert out[3] == pytest.approx(1 / (4 * np.pi), rel=1e-6)
        assert np.allclose(again[7:10], out[0:3], rtol=5e-3, atol=1e-4)  # Evaluate(LocalToWorld * d) == the value sampled for d


@pytest.mark.parametrize("fixture", ["environment_small", "cubemap_small"])
def test_directional_texture_average(fixture, request):
    """DirectionalTextureBaseTests.Average (src/Echo.UnitTests/Textures/DirectionalTextureTests.cs:95-101): `AverageConverge` — the mean
    of Evaluate over stratified uniform-sphere directions (IDirectionalTexture.cs:59-88) — equals the `Average` the texture's
    Prepare computed, which is what the light's power is built from. Here: the oracle's Evaluate against the host mirror's
    averages (CylindricalTexture.Prepare :93-95; the cubemap's stand-in), on a 160 x 160 stratification instead of 1000 x 1000,
    so the tolerance is that sampling's error instead of the reference's AlmostZero."""
    prepared = request.getfixturevalue(fixture)
    oracle = oracle_lib.OracleScene(prepared)
    light = prepared.description.infinite_lights[0]
    textures = prepared.description.textures
    if fixture == "environment_small":
        _, expected = host.build_environment(textures[int(light["texture"])])
    else:
        first = int(light["texture"])
        expected = host.cubemap_average(textures[first:first + 6])

    size = 160
    rng = np.random.default_rng(5)
    total = np.zeros(3)
    for y in range(size):
        for x in range(size):
            u, v = (x + rng.random()) / size, (y + rng.random()) / size
            z = 1 - 2 * u  # Sample2D.UniformSphere, Sample2D.cs:35
            r = np.sqrt(max(0.0, 1 - z * z))
            direction = (r * np.cos(2 * np.pi * v), r * np.sin(2 * np.pi * v), z)
            total += infinite_light(oracle, 0, (0.5, 0.5), direction)[7:10]
    average = total / size ** 2 / np.asarray(light["radiance"], dtype=np.float64)
    assert np.allclose(average, expected, rtol=0.01), (average, expected)
