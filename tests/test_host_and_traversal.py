"""CPU checks of the host-side preparation (libecho_host.so) and of the oracle's traversal / light tree / integrator.

The reference has no tests for QuadBoundingVolumeHierarchy, BoxBound4, PreparedTriangle.Intersect, LightTree or
PathTracedEvaluator, so these parts are "parity unpinned" against the reference itself; they are pinned here
by the cross-checks SURVEY.md asks for: brute-force linear intersection, structural invariants, probability-mass
consistency and the committed regression fixtures under tests/golden/.
"""
import os

import numpy as np
import pytest

from echorenderer_b200 import host, scenes, structs
from tests import oracle_lib as ol

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def walk(prepared):
    """Depth-first walk of the node array; returns (leaf tokens in order, max quad depth)."""
    nodes = prepared.nodes
    tokens, depth_max = [], 0
    stack = [(0, 1)]
    while stack:
        index, depth = stack.pop()
        depth_max = max(depth_max, depth)
        for token in nodes[index]["token4"]:
            if token == structs.TOKEN_EMPTY:
                continue
            if structs.token_type(token) == structs.TOKEN_TYPE_NODE:
                stack.append((int(structs.token_index(token)), depth + 1))
            else:
                tokens.append(int(token))
                depth_max = max(depth_max, depth + 1)
    return tokens, depth_max


@pytest.mark.parametrize("fixture", ["cornell", "terrain_small", "mixed_small"])
def test_qbvh_structure(fixture, request):
    """QuadBoundingVolumeHierarchy constructor invariants (QuadBoundingVolumeHierarchy.cs:24-36,363-565)."""
    prepared = request.getfixturevalue(fixture)
    nodes = prepared.nodes
    count = len(prepared.triangles) + len(prepared.spheres)

    tokens, depth = walk(prepared)
    assert sorted(tokens) == sorted([int(structs.make_token(1, i)) for i in range(len(prepared.triangles))] +
                                    [int(structs.make_token(2, i)) for i in range(len(prepared.spheres))])  # every primitive exactly once
    assert depth == prepared.max_depth                      # stackSize = maxDepth * 3 + 1 relies on it (:34)
    assert count / 3 - 1 <= len(nodes) <= count             # every other binary level collapses

    # pre-order layout: children of node i have larger indices, the first child node is i + 1
    for i, node in enumerate(nodes):
        children = [int(structs.token_index(t)) for t in node["token4"] if t != structs.TOKEN_EMPTY and structs.token_type(t) == 0]
        assert all(c > i for c in children)
        if children:
            assert min(children) == i + 1
        assert 0 <= node["axisMajor"] <= 2 and 0 <= node["axisMinor0"] <= 3 and 0 <= node["axisMinor1"] <= 3
        # axis 3 marks a [leaf, empty] pair (:531-533); empty slots carry BoxBound.None = (+inf, +inf)
        for pair, axis in ((0, node["axisMinor0"]), (2, node["axisMinor1"])):
            if axis == 3:
                assert node["token4"][pair + 1] == structs.TOKEN_EMPTY and structs.token_type(node["token4"][pair]) != 0
        empty = node["token4"] == structs.TOKEN_EMPTY
        assert np.all(np.isposinf(node["minX"][empty])) and np.all(np.isposinf(node["maxX"][empty]))
        # children sorted by their min along the split axis (GetChildrenSorted, :551-563)
        for pair, axis in ((0, node["axisMinor0"]), (2, node["axisMinor1"])):
            if axis < 3:
                key = ("minX", "minY", "minZ")[axis]
                assert node[key][pair] <= node[key][pair + 1]

    # every child box lies inside its parent's slot box
    for i, node in enumerate(nodes):
        for slot, token in enumerate(node["token4"]):
            if token == structs.TOKEN_EMPTY or structs.token_type(token) != 0:
                continue
            child = nodes[int(structs.token_index(token))]
            valid = child["token4"] != structs.TOKEN_EMPTY
            for low, high in (("minX", "maxX"), ("minY", "maxY"), ("minZ", "maxZ")):
                assert child[low][valid].min() >= node[low][slot] and child[high][valid].max() <= node[high][slot]


@pytest.mark.parametrize("fixture", ["cornell", "terrain_small", "mixed_small", "lights_small"])
def test_qbvh_trace_matches_linear(fixture, request):
    """Traversal against brute force over every primitive (the LinearAccelerator cross-check). Near-ties (hits closer than
    4 ulp, e.g. Cornell's coplanar box bottoms) may resolve to either primitive: traversal order decides, like the reference."""
    prepared = request.getfixturevalue(fixture)
    oracle = ol.OracleScene(prepared)
    rays = scenes.random_rays(prepared.bounds, 30_000, seed=5)
    fast, slow = oracle.trace(rays), oracle.trace_linear(rays)

    hit_fast, hit_slow = fast["token"] != structs.TOKEN_EMPTY, slow["token"] != structs.TOKEN_EMPTY
    assert np.array_equal(hit_fast, hit_slow)
    both = hit_fast
    ulp = np.abs(fast["distance"][both].view(np.int32).astype(np.int64) - slow["distance"][both].view(np.int32).astype(np.int64))
    assert ulp.max() <= 4
    different = fast["token"][both] != slow["token"][both]
    assert np.all(ulp[different] <= 4) and different.mean() < 0.01

    secondary = scenes.secondary_rays(prepared, rays, fast)[:8000]
    fast2, slow2 = oracle.trace(secondary), oracle.trace_linear(secondary)
    ulp = np.abs(fast2["distance"].view(np.int32).astype(np.int64) - slow2["distance"].view(np.int32).astype(np.int64))
    assert (ulp > 4).mean() < 1e-3  # self-hit rule is token based: an "ignore" token never hits itself
    assert not np.any(fast2["token"] == secondary["ignore"]) or fixture != "cornell"

    shadow = scenes.random_rays(prepared.bounds, 30_000, seed=5, occlusion=True)
    a, b = oracle.occlude(shadow), oracle.occlude_linear(shadow)
    assert (a != b).mean() < 1e-3


def test_trace_guards(terrain_small):
    """PreparedScene.Trace / Occlude reject queries whose distance is not Positive (PreparedScene.cs:66-86)."""
    oracle = ol.OracleScene(terrain_small)
    rays = scenes.random_rays(terrain_small.bounds, 512, seed=3)
    for limit in (0.0, -1.0, 7e-7, float("nan")):
        rays["distance"] = limit
        hits = oracle.trace(rays)
        assert np.all(hits["token"] == structs.TOKEN_EMPTY)
        assert np.all(oracle.occlude(rays) == 0)


def test_light_tree_structure_and_mass(lights_small):
    """LightTree (LightTree.cs:21-154): every emitter appears once with its branch bit path; Pick's pdf equals ProbabilityMass;
    the masses of all emitters sum to one wherever some importance is positive."""
    prepared = lights_small
    nodes = prepared.light_nodes
    emitters = len(prepared.emitter_tokens)
    assert emitters == 300 and len(nodes) == 2 * emitters - 1

    # re-derive every bit path by walking from the root
    paths = {}
    stack = [(0, 0, 0)]
    while stack:
        index, depth, bits = stack.pop()
        node = nodes[index]
        if node["child0"] == structs.TOKEN_EMPTY:
            paths[int(node["child1"])] = bits
            continue
        assert depth < 64
        stack.append((int(node["child0"]), depth + 1, bits))
        stack.append((int(node["child1"]), depth + 1, bits | (1 << depth)))
        # a branch's power is the sum of its children's (LightBound.Encapsulate)
        assert np.isclose(node["power"], nodes[int(node["child0"])]["power"] + nodes[int(node["child1"])]["power"], rtol=1e-5)
    assert paths == {int(t): int(p) for t, p in zip(prepared.emitter_tokens, prepared.emitter_bitpaths)}

    oracle = ol.OracleScene(prepared)
    rng = np.random.default_rng(9)

    for _ in range(40):
        position = rng.uniform([-25, 0.5, -25], [25, 15, 25])
        normal = rng.normal(size=3)
        normal /= np.linalg.norm(normal)

        # a branch whose two children both have zero importance yields 0 / 0 = NaN in ProbabilityMass (LightTree.cs:145-147 has
        # no guard, unlike Pick :122); the evaluator drops such masses with !Positive(pmf). The finite masses sum to <= 1.
        masses = np.array([oracle.light_mass(token, position, normal) for token in prepared.emitter_tokens])
        total = np.nansum(masses)
        assert 0.5 < total <= 1 - prepared.infinite_threshold + 1e-3
        if not np.isnan(masses).any():
            assert abs(total - (1 - prepared.infinite_threshold)) < 1e-3

        for sample in rng.random(25):
            token, pdf = oracle.light_pick(position, normal, sample)
            if pdf == 0:
                continue
            assert token in paths
            mass = oracle.light_mass(token, position, normal)
            assert abs(mass - pdf) <= 2e-5 * pdf + 1e-9  # same factors, multiplied root-to-leaf vs leaf-to-root


def test_light_pick_distribution(lights_small):
    """Pick draws emitters proportionally to ProbabilityMass (stratified 1-D samples, chi-square-free bound)."""
    oracle = ol.OracleScene(lights_small)
    position, normal = np.array([0.0, 6.0, 0.0]), np.array([0.0, 1.0, 0.0])
    n = 200_000
    picks = {}
    for sample in (np.arange(n) + 0.5) / n:
        token, pdf = oracle.light_pick(position, normal, sample)
        if pdf > 0:
            picks[token] = picks.get(token, 0) + 1
    for token, count in sorted(picks.items(), key=lambda kv: -kv[1])[:20]:
        mass = oracle.light_mass(token, position, normal)
        assert abs(count / n - mass) <= 0.02 * mass + 2 / n


def test_triangle_sample_pdf_consistency(lights_small):
    """PreparedTriangle.Sample / ProbabilityDensity (TriangleEntity.cs:166-185) agree through the oracle's hooks."""
    oracle = ol.OracleScene(lights_small)
    rng = np.random.default_rng(4)
    origin = np.array([1.0, 5.0, -2.0])
    checked = 0
    for token in lights_small.emitter_tokens[:60]:
        for _ in range(5):
            ok, point, normal, pdf = oracle.geometry_sample(int(token), origin, rng.random(2))
            if not ok or pdf < 1e-6:
                continue
            delta = point.astype(np.float64) - origin
            incident = (delta / np.linalg.norm(delta)).astype(np.float32)
            density = oracle.geometry_pdf(int(token), origin, incident)
            if density == 0:
                continue  # the re-intersection can miss at the very edge of the triangle
            assert abs(density - pdf) <= 2e-3 * pdf
            checked += 1
    assert checked > 100


def test_render_converges_to_brute_force_lighting(cornell):
    """The integrator (NEE + MIS + RR) is unbiased: a many-sample Cornell pixel block matches the same block rendered with
    twice the samples and a different seed within Monte-Carlo error, and the image has the Cornell colour layout."""
    oracle = ol.OracleScene(cornell)
    tiles = np.array([[1, 1], [2, 1]], dtype=np.int32)
    a, _ = oracle.render_tiles(structs.render_params(64, 64, 16, extend=256, seed=1), tiles)
    b, _ = oracle.render_tiles(structs.render_params(64, 64, 16, extend=512, seed=2), tiles)
    mean_a, mean_b = a[..., :3].mean(axis=(0, 1, 2)), b[..., :3].mean(axis=(0, 1, 2))
    assert np.all(np.abs(mean_a - mean_b) <= 0.02 * mean_b)


def test_golden_fixtures(cornell, terrain_small):
    """Regression fixtures generated by tests/golden/make_golden.py from the oracle (NOT from the reference, which cannot run
    here): pins the oracle's hit records and per-sample radiance bit for bit against accidental edits."""
    data = np.load(os.path.join(GOLDEN, "oracle_regression.npz"))

    for name, prepared in (("cornell", cornell), ("terrain", terrain_small)):
        oracle = ol.OracleScene(prepared)
        rays = scenes.random_rays(prepared.bounds, 4096, seed=21)
        hits = oracle.trace(rays)
        assert np.array_equal(hits["token"], data[f"{name}_token"])
        assert np.array_equal(hits["distance"].view(np.uint32), data[f"{name}_distance_bits"])

    oracle = ol.OracleScene(cornell)
    params = structs.render_params(32, 32, 16, extend=4, seed=9)
    ys, xs = np.meshgrid(np.arange(0, 32, 2), np.arange(0, 32, 2), indexing="ij")
    pixels = np.repeat(np.stack([xs.reshape(-1), ys.reshape(-1)], axis=-1), 2, axis=0).astype(np.int32)
    index = np.tile(np.arange(2, dtype=np.uint32), len(pixels) // 2)
    radiance = oracle.evaluate_samples(params, pixels, index)
    assert np.array_equal(radiance.view(np.uint32), data["cornell_radiance_bits"])

    # the rows added after the hot path: instanced packs, image textures, the auxiliary passes
    instanced = host.prepare(scenes.instanced_scene(grid=3, rings=8, segments=10))
    oracle = ol.OracleScene(instanced)
    hits, layers = oracle.trace_hierarchy(scenes.random_rays(instanced.bounds, 4096, seed=21))
    assert np.array_equal(hits["token"], data["instanced_token"]) and np.array_equal(hits["distance"].view(np.uint32), data["instanced_distance_bits"])
    assert np.array_equal(np.concatenate([layers["instanceCount"][:, None], layers["instances"]], axis=1), data["instanced_layers"])
    params = structs.render_params(32, 32, 16, extend=2, seed=9, bounce_limit=12)
    half_pixels, half_index = np.ascontiguousarray(pixels[::2]), np.ascontiguousarray(index[::2])
    assert np.array_equal(oracle.evaluate_samples(params, half_pixels, half_index).view(np.uint32), data["instanced_radiance_bits"])

    textured = host.prepare(scenes.textured_scene(rings=8, segments=10))
    oracle = ol.OracleScene(textured)
    assert np.array_equal(oracle.evaluate_samples(params, half_pixels, half_index).view(np.uint32), data["textured_radiance_bits"])
    for name, code in (("albedo", structs.EVALUATOR_ALBEDO | structs.EVALUATOR_DIVERGE_ONCE), ("normal_depth", structs.EVALUATOR_NORMAL_DEPTH)):
        aux = structs.render_params(32, 32, 16, extend=2, seed=9, evaluator=code)
        value = np.zeros((len(half_index), 4), dtype=np.float32)
        oracle.lib.oracle_evaluate_samples4(oracle.handle, ol.ptr(aux), ol.ptr(half_pixels), ol.ptr(half_index), len(value), ol.ptr(value), 4, 0)
        assert np.array_equal(value.view(np.uint32), data[f"textured_{name}_bits"])


@pytest.mark.parametrize("angle", [0.0, 2.0])
def test_directional_light_irradiance(angle):
    """DirectionalLight (Scenic/Lights/DirectionalLight.cs:52-108): a lone Lambertian plane under a light tilted by 60 degrees
    shows albedo / pi * Intensity * cos(60) whether the light is a delta light or a narrow cone (":66 Maintain consistent
    intensity regardless of the angle"), and the light is picked, sampled and weighed as InfiniteDelta / Infinite."""
    albedo, intensity = 0.6, 3.0
    description = host.SceneDescription(triangles=scenes.plane(0, (200, 200)), materials=scenes.material(structs.MATERIAL_DIFFUSE, (albedo,) * 3),
                                        infinite_lights=scenes.directional_light((intensity,) * 3, (90 - 60, 0, 0), angle=angle))
    position = (0.0, 5.0, -3.0)
    description.camera = scenes.perspective_camera(position, scenes.look_rotation(position, (0, 0, 0)), field_of_view=30.0)
    prepared = host.prepare(description)
    light = prepared.description.infinite_lights[0]
    assert bool(light["isDelta"]) == (angle == 0.0) and prepared.infinite_threshold == 1.0  # no other light: always picked
    assert np.allclose(light["direction"], [0, 0.5, -np.cos(np.radians(30))], atol=1e-6)  # incidentDirection points TOWARDS the light

    oracle = ol.OracleScene(prepared)
    params = structs.render_params(32, 32, 16, extend=256, seed=5, bounce_limit=4)
    image, stats = oracle.render_tiles(params, scenes.tile_grid(32, 32, 16))
    cosine = abs(float(light["direction"][1]))  # the plane's normal is +Y
    expected = albedo / np.pi * intensity * cosine
    assert image[..., :3].mean() == pytest.approx(expected, rel=0.02)
    checked, passed = int(stats["lightOcclusionChecked"][0]), int(stats["lightOcclusionPassed"][0])
    assert checked > 0 and passed >= 0.999 * checked  # nothing shadows the plane (but its own other triangle, along the diagonal seam)


@pytest.mark.parametrize("seed", [0, 7, 22])
def test_hostile_rays_on_degenerate_scenes(seed):
    """The oracle on the fuzz inputs of tests/test_gpu_fuzz.py: queries with infinite / NaN / overflowing components must not
    walk off the node array (an accepted NaN hit distance lets empty children through the distance test; the reference would
    index out of range there, QuadBoundingVolumeHierarchy.cs:211), and for well-formed rays the tree agrees with brute force."""
    from tests.test_gpu_fuzz import hostile_rays, random_scene
    rng = np.random.default_rng(1000 + seed)
    prepared, scale = random_scene(rng)
    oracle = ol.OracleScene(prepared)
    rays = hostile_rays(prepared, rng, scale)
    hits = oracle.trace(rays)
    oracle.occlude(rays)

    sane = np.isfinite(rays["origin"]).all(axis=1) & np.isfinite(rays["direction"]).all(axis=1) & (np.abs(rays["direction"]).max(axis=1) < 2) \
        & (np.abs(rays["origin"]).max(axis=1) < 1e30)
    linear = oracle.trace_linear(rays[sane])
    same = hits["token"][sane] == linear["token"]
    # exact ties between coincident triangles resolve by visit order, which differs between the tree and the brute-force loop
    assert same.mean() > 0.97
    assert np.array_equal(hits["distance"][sane][same].view(np.uint32), linear["distance"][same].view(np.uint32))
    tied = ~same
    assert np.array_equal(hits["distance"][sane][tied], linear["distance"][tied])


@pytest.mark.parametrize("size", [(10, 20), (31, 13), (1, 3), (1, 1), (123, 456)])
@pytest.mark.parametrize("pattern", ["ordered", "ordered_vertical", "hilbert"])
def test_tile_patterns(pattern, size):
    """TilePatternTests.CreateSequence (src/Echo.UnitTests/Processes/TilePatternTests.cs:18-31, the reference's own sizes): every
    position of the grid exactly once, all inside it. Plus what makes the Hilbert pattern the default of EvaluationProfile: the
    sequence starts at the centre and the four interlaced quadrant curves each move one tile at a time."""
    from echorenderer_b200 import hilbert_curve_pattern, ordered_pattern
    positions = {"ordered": lambda s: ordered_pattern(s), "ordered_vertical": lambda s: ordered_pattern(s, False), "hilbert": hilbert_curve_pattern}[pattern](size)
    assert positions.shape == (size[0] * size[1], 2)
    assert np.all(positions >= 0) and np.all(positions < np.array(size))
    assert len({(int(x), int(y)) for x, y in positions}) == len(positions)
    if pattern == "hilbert" and min(size) >= 10:
        centre = np.array(size) // 2
        assert np.abs(positions[:4] - centre).max() <= 1
        head = positions[: 4 * (min(size) // 2) ** 2 // 4 * 4].reshape(-1, 4, 2)  # while all four quadrant curves are still running
        steps = np.abs(np.diff(head, axis=0)).sum(axis=2)
        assert steps.max() <= 2 and (steps == 1).mean() > 0.95  # the generalised curve takes a diagonal step on odd sizes


def test_render_texture_uses_the_default_pattern():
    """EvaluationOperation.cs:174: `profile.Pattern.CreateSequence(size.CeiledDivide(tileSize))`, HilbertCurvePattern by default."""
    from echorenderer_b200 import RenderTexture, hilbert_curve_pattern, shard_tiles
    texture = RenderTexture(100, 50, 16)
    assert texture.tile_count == (7, 4)
    assert np.array_equal(texture.tile_positions, hilbert_curve_pattern((7, 4)))
    shards = [shard_tiles(texture.tile_positions, rank, 4) for rank in range(4)]
    assert sorted(map(tuple, np.concatenate(shards).tolist())) == sorted(map(tuple, texture.tile_positions.tolist()))
