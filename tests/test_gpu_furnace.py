"""The closed-form known answers of tests/test_furnace.py asked of the DEVICE directly (no oracle in the comparison): the per-sample exact
ones — a point light over a Lambertian plane, a smooth conductor sheet under a half sky, a glass sphere in a uniform environment."""
import numpy as np
import pytest

from echorenderer_b200 import PreparedScene, host, scenes, structs
from tests import oracle_lib
from tests.test_furnace import RADIANCE, furnace, half_sky, lit_plane, point_light_radiance

pytestmark = pytest.mark.gpu


def device_samples(description, size, extend, seed, bounce_limit=16):
    prepared = host.prepare(description)
    params = structs.render_params(size, size, 8, extend=extend, seed=seed, bounce_limit=bounce_limit)
    ys, xs = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    pixels = np.repeat(np.stack([xs.reshape(-1), ys.reshape(-1)], axis=-1), extend, axis=0).astype(np.int32)
    index = np.tile(np.arange(extend, dtype=np.uint32), size * size)
    rays = oracle_lib.OracleScene(prepared).spawn_rays(params, pixels, index)  # the camera rays (device == oracle: test_gpu_render.py)
    with PreparedScene(prepared) as scene:
        radiance = scene.evaluate_samples(params, pixels, index).astype(np.float64)
        hits = scene.trace(rays)
    hit = hits["token"] != structs.TOKEN_EMPTY
    points = rays["origin"].astype(np.float64) + rays["direction"].astype(np.float64) * np.where(hit, hits["distance"], 0.0).astype(np.float64)[:, None]
    return radiance[:, :3], hit, points


def test_device_point_light_over_a_plane_is_the_closed_form_per_sample():
    light = ((30.0, 20.0, 10.0), (1.0, 3.0, -0.5))
    radiance, hit, points = device_samples(lit_plane([light]), 16, 2, seed=8)
    assert hit.mean() > 0.8 and np.all(radiance[~hit] == 0)
    assert np.allclose(radiance[hit], point_light_radiance(points[hit], *light), rtol=5e-6, atol=0)


def test_device_conductor_mirror_and_glass_furnace():
    n, k = np.array((0.18, 0.42, 1.37)), np.array((3.42, 2.35, 1.77))
    physical = scenes.material(structs.MATERIAL_CONDUCTOR, (1.0, 1.0, 1.0), param_a=tuple(n), param_b=tuple(k), flags=0)
    radiance, hit, points = device_samples(half_sky(np.concatenate([physical])), 12, 4, seed=20)
    outgoing = np.array((0.0, 6.0, -7.0)) - points[hit]
    cos = (outgoing[:, 1] / np.linalg.norm(outgoing, axis=1))[:, None]
    cos2, sin2 = cos * cos, 1.0 - cos * cos
    term = n * n - k * k - sin2
    a2b2 = np.sqrt(term * term + 4.0 * n * n * k * k)
    para0, para1 = a2b2 + cos2, cos * np.sqrt(2.0) * np.sqrt(a2b2 + term)
    perp0, perp1 = cos2 * a2b2 + sin2 * sin2, para1 * sin2
    para, perp = (para0 - para1) / (para0 + para1), (perp0 - perp1) / (perp0 + perp1)
    assert hit.sum() > 400 and np.allclose(radiance[hit], (para * perp + para) / 2.0 * np.array(RADIANCE), rtol=3e-5, atol=0)

    description = furnace("sphere")
    description.materials = np.concatenate([scenes.material(structs.MATERIAL_DIELECTRIC, (1.0, 1.0, 1.0), roughness=(0.0, 0.0), ior=1.5)])
    radiance, hit, _ = device_samples(description, 24, 8, seed=8, bounce_limit=128)
    assert hit.sum() > 1000 and np.allclose(radiance, np.array(RADIANCE), rtol=2e-5, atol=0)
