"""Randomised parity: many small scenes with degenerate content (zero-area and duplicated triangles, zero-radius and nested
spheres, coincident geometry, tiny and huge scales) queried with hostile rays (axis-aligned, starting on vertices and
surfaces, zero / denormal / infinite / NaN components, all limits) — closest hit and occlusion must match the oracle bit
for bit, through the persistent kernels and through the one-thread-per-ray kernels (visit counters)."""
import numpy as np
import pytest

from echorenderer_b200 import PreparedScene, host, scenes, structs
from tests import oracle_lib
pytestmark = pytest.mark.gpu


def assert_hits_equal(actual, expected):
    """Bit-exact tokens and distances; barycentrics bit-exact too, except that two NaNs (an overflow on absurd coordinates)
    count as equal whatever their payload."""
    assert np.array_equal(actual["token"], expected["token"])
    assert np.array_equal(actual["distance"].view(np.uint32), expected["distance"].view(np.uint32))
    hit = expected["token"] != structs.TOKEN_EMPTY
    a, e = actual["uv"][hit], expected["uv"][hit]
    both_nan = np.isnan(a) & np.isnan(e)
    assert np.array_equal(a.view(np.uint32)[~both_nan], e.view(np.uint32)[~both_nan])


def random_scene(rng):
    scale = float(10.0 ** rng.integers(-3, 4))
    count = int(rng.integers(2, 400))
    v0 = rng.normal(size=(count, 3)) * scale
    v1 = v0 + rng.normal(size=(count, 3)) * scale * rng.choice([1e-4, 0.1, 1.0, 5.0], size=(count, 1))
    v2 = v0 + rng.normal(size=(count, 3)) * scale * rng.choice([1e-4, 0.1, 1.0, 5.0], size=(count, 1))

    degenerate = rng.random(count) < 0.05
    v2[degenerate] = v1[degenerate]                        # zero area
    copies = rng.random(count) < 0.05
    source = rng.integers(0, count, size=count)
    v0[copies], v1[copies], v2[copies] = v0[source[copies]], v1[source[copies]], v2[source[copies]]  # coincident triangles (exact ties)
    snapped = rng.random(count) < 0.2
    v0[snapped], v1[snapped], v2[snapped] = np.round(v0[snapped] / scale) * scale, np.round(v1[snapped] / scale) * scale, np.round(v2[snapped] / scale) * scale

    triangles = scenes.make_triangles(v0, v1, v2, 0)
    spheres = np.zeros(int(rng.integers(0, 40)), dtype=structs.SPHERE)
    spheres["position"] = rng.normal(size=(len(spheres), 3)) * scale
    spheres["radius"] = np.abs(rng.normal(size=len(spheres))) * scale * rng.choice([0.0, 0.01, 1.0, 3.0], size=len(spheres))
    description = host.SceneDescription(triangles=triangles, spheres=spheres, materials=scenes.material(structs.MATERIAL_DIFFUSE), camera=scenes.cornell_box().camera)
    return host.prepare(description), scale


def hostile_rays(prepared, rng, scale, count=4000):
    rays = scenes.random_rays(prepared.bounds, count, seed=int(rng.integers(1, 1 << 30)))
    kind = rng.integers(0, 10, size=count)

    axis = np.eye(3, dtype=np.float32)[rng.integers(0, 3, size=count)] * rng.choice([-1.0, 1.0], size=(count, 1)).astype(np.float32)
    rays["direction"][kind == 0] = axis[kind == 0]                                         # axis-aligned: two reciprocals are +-inf
    plane = rays["direction"].copy()
    plane[:, 1] = 0
    plane /= np.maximum(np.linalg.norm(plane, axis=1, keepdims=True), 1e-30)
    rays["direction"][kind == 1] = plane[kind == 1]                                        # one zero component

    triangles = prepared.triangles
    pick = rng.integers(0, len(triangles), size=count)
    on_vertex = triangles["vertex0"][pick]
    on_surface = on_vertex + 0.3 * triangles["edge1"][pick] + 0.3 * triangles["edge2"][pick]
    rays["origin"][kind == 2] = on_vertex[kind == 2]                                       # starting exactly on a vertex
    rays["origin"][kind == 3] = on_surface[kind == 3]                                      # starting on a surface, ignoring it
    rays["ignore"][kind == 3] = (structs.TOKEN_TYPE_TRIANGLE << structs.TOKEN_INDEX_BITS) | pick[kind == 3].astype(np.uint32)
    toward = on_surface - rays["origin"]
    toward /= np.maximum(np.linalg.norm(toward, axis=1, keepdims=True), 1e-30)
    rays["direction"][kind == 4] = toward[kind == 4].astype(np.float32)                    # aimed at geometry

    limits = np.array([0.0, -1.0, 1e-30, 8e-7, 7.9e-7, np.inf, scale, scale * 1e-3], dtype=np.float32)
    rays["distance"][kind == 5] = limits[rng.integers(0, len(limits), size=int((kind == 5).sum()))]
    rays["distance"][kind == 6] = (np.abs(rng.normal(size=int((kind == 6).sum()))) * scale * 3).astype(np.float32)

    special = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1e-40, 3e38], dtype=np.float32)
    rows = np.flatnonzero(kind == 7)
    rays["origin"][rows, rng.integers(0, 3, size=len(rows))] = special[rng.integers(0, len(special), size=len(rows))]
    rows = np.flatnonzero(kind == 8)
    rays["direction"][rows, rng.integers(0, 3, size=len(rows))] = special[rng.integers(0, len(special), size=len(rows))]
    return rays


@pytest.mark.parametrize("seed", range(24))
def test_random_scene_parity(seed):
    rng = np.random.default_rng(1000 + seed)
    prepared, scale = random_scene(rng)
    oracle = oracle_lib.OracleScene(prepared)
    rays = hostile_rays(prepared, rng, scale)

    with PreparedScene(prepared) as scene:
        expected = oracle.trace(rays)
        assert_hits_equal(scene.trace(rays), expected)

        shadow = rays.copy()
        finite = np.isfinite(shadow["distance"]) & (shadow["distance"] > 0)
        shadow["distance"][~finite & (rng.random(len(rays)) < 0.5)] = np.float32(scale)
        assert np.array_equal(scene.occlude(shadow), oracle.occlude(shadow))

        # the one-thread-per-ray kernels (the ones behind the visit counters) walk the same tree in the same order
        import torch
        d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
        d_hits = torch.empty(len(rays) * 16, dtype=torch.uint8, device="cuda")
        d_counts = torch.zeros(3, dtype=torch.int64, device="cuda")
        scene.trace_device(d_rays.data_ptr(), len(rays), d_hits.data_ptr(), 0, d_counts.data_ptr())
        torch.cuda.synchronize()
        assert_hits_equal(d_hits.cpu().numpy().view(structs.HIT), expected)
        _, counters = oracle.trace(rays, count_visits=True)
        assert np.array_equal(d_counts.cpu().numpy().astype(np.uint64), counters)


def random_materials(rng, count):
    """A swatch of every material type with parameters drawn from the corners of their ranges."""
    edge = lambda: float(rng.choice([0.0, 1e-6, 0.01, 0.3, 1.0, 1.5]))
    records = []
    for k in range(count):
        kind = int(rng.choice([structs.MATERIAL_DIFFUSE, structs.MATERIAL_DIELECTRIC, structs.MATERIAL_CONDUCTOR, structs.MATERIAL_COATED_DIFFUSE,
                               structs.MATERIAL_EMISSIVE, structs.MATERIAL_ONESIDED, structs.MATERIAL_INVISIBLE]))
        albedo = tuple(float(v) for v in rng.choice([0.0, 0.05, 0.5, 0.9, 1.0], size=3))
        alpha = float(rng.choice([1.0, 1.0, 1.0, 0.49, 0.5]))
        roughness = (edge(), edge())
        ior = float(rng.choice([1.0, 1.0001, 1.33, 1.5, 2.4, 0.7]))
        flags = 0
        if kind == structs.MATERIAL_DIFFUSE and rng.random() < 0.3:
            flags |= structs.MATERIAL_FLAG_TRANSMISSIVE
        if kind == structs.MATERIAL_CONDUCTOR and rng.random() < 0.6:
            flags |= structs.MATERIAL_FLAG_ARTISTIC
        if kind == structs.MATERIAL_ONESIDED and rng.random() < 0.5:
            flags |= structs.MATERIAL_FLAG_BACKFACE
        param_a = tuple(float(v) for v in rng.choice([0.0, 0.2, 0.9, 1.0, 3.0], size=3))
        param_b = tuple(float(v) for v in rng.choice([0.0, 0.5, 1.0, 4.0], size=3))
        if kind == structs.MATERIAL_EMISSIVE:
            albedo = tuple(float(v) for v in rng.choice([0.0, 2.0, 9.0], size=3))
        record = scenes.material(kind, albedo, alpha, roughness, ior, param_a, param_b, flags, base=int(rng.integers(0, max(k, 1))))
        if kind == structs.MATERIAL_COATED_DIFFUSE:
            record["paramA"][0][0] = host.fresnel_diffuse_reflectance(np.float32(1.0) / np.float32(ior))
        if kind == structs.MATERIAL_ONESIDED and k == 0:
            record["type"] = structs.MATERIAL_DIFFUSE  # a OneSided needs an earlier material to wrap
        records.append(record)
    return np.concatenate(records)


@pytest.mark.parametrize("seed", range(10))
def test_random_materials_render_parity(seed):
    """Per-sample radiance with random swatches (specular and rough limits, index 1, alpha cut-outs, two-sided diffuse, OneSided
    chains, black and bright emitters) against the oracle."""
    from tests.test_gpu_render import relative_rmse, sample_grid
    rng = np.random.default_rng(500 + seed)
    description = scenes.mixed_material_scene(rings=12, segments=14)
    count = 12
    description.materials = random_materials(rng, count)
    description.triangles["material"] = rng.integers(0, count, size=len(description.triangles)) if seed % 2 else description.triangles["material"] % count
    description.spheres["material"] = rng.integers(0, count, size=len(description.spheres))
    description.point_lights = np.zeros(1, dtype=structs.POINT_LIGHT)
    description.point_lights["intensity"], description.point_lights["position"] = (30.0, 28.0, 25.0), (2.0, 9.0, -6.0)
    prepared = host.prepare(description)
    oracle = oracle_lib.OracleScene(prepared)

    width, height = 64, 40
    params = structs.render_params(width, height, 16, extend=4, bounce_limit=24, seed=3 + seed)
    pixel_xy, sample_index = sample_grid(width, height, 4)

    with PreparedScene(prepared) as scene:
        actual = scene.evaluate_samples(params, pixel_xy, sample_index)

    expected = oracle.evaluate_samples(params, pixel_xy, sample_index)
    both_nan = np.isnan(actual) & np.isnan(expected)
    different = np.any((actual.view(np.uint32) != expected.view(np.uint32)) & ~both_nan, axis=1)
    assert different.mean() <= 1e-4, f"{different.sum()} of {len(different)} samples are not bit-identical"


def random_instanced_scene(rng):
    """Random packs placed with random rotations, uniform scales over six decades and random nesting (up to four layers)."""
    def random_pack(instances):
        count = int(rng.integers(2, 60))
        v0 = rng.normal(size=(count, 3))
        triangles = scenes.make_triangles(v0, v0 + rng.normal(size=(count, 3)) * 0.6, v0 + rng.normal(size=(count, 3)) * 0.6, 0)
        spheres = np.zeros(int(rng.integers(0, 6)), dtype=structs.SPHERE)
        spheres["position"], spheres["radius"] = rng.normal(size=(len(spheres), 3)), np.abs(rng.normal(size=len(spheres))) * 0.5
        return host.PackDescription(triangles=triangles, spheres=spheres, materials=scenes.material(structs.MATERIAL_DIFFUSE), instances=instances)

    def placement(pack):
        scale = float(10.0 ** rng.uniform(-2, 2))
        return host.InstanceDescription(pack, tuple(rng.normal(size=3) * 3), tuple(rng.uniform(0, 360, size=3)), scale)

    packs = [random_pack([])]
    for level in range(int(rng.integers(1, 4))):  # pack k + 1 places one to three copies of earlier packs
        packs.append(random_pack([placement(int(rng.integers(0, len(packs)))) for _ in range(int(rng.integers(1, 4)))]))

    root = random_pack([placement(int(rng.integers(0, len(packs)))) for _ in range(int(rng.integers(1, 12)))])
    description = host.SceneDescription(triangles=root.triangles, spheres=root.spheres, materials=root.materials, instances=root.instances, packs=packs,
                                        camera=scenes.cornell_box().camera)
    return host.prepare(description)


@pytest.mark.parametrize("seed", range(12))
def test_random_instanced_parity(seed):
    from tests.test_instancing import spawn_from_hits
    rng = np.random.default_rng(2000 + seed)
    prepared = random_instanced_scene(rng)
    oracle = oracle_lib.OracleScene(prepared)
    rays = hostile_rays(prepared, rng, 1.0, count=6000)
    rays["ignore"] = structs.TOKEN_EMPTY

    with PreparedScene(prepared) as scene:
        expected, expected_layers = oracle.trace_hierarchy(rays)
        hits, layers = scene.trace_hierarchy(rays)
        assert_hits_equal(hits, expected)
        assert np.array_equal(layers, expected_layers)
        assert np.array_equal(scene.occlude_hierarchy(rays), oracle.occlude_hierarchy(rays))

        spawned, ignore = spawn_from_hits(rays, expected, expected_layers)
        sane = np.isfinite(spawned["origin"]).all(axis=1)
        spawned, ignore = spawned[sane], ignore[sane]
        again, again_layers = scene.trace_hierarchy(spawned, ignore)
        expected_again, expected_again_layers = oracle.trace_hierarchy(spawned, ignore)
        assert_hits_equal(again, expected_again)
        assert np.array_equal(again_layers, expected_again_layers)
