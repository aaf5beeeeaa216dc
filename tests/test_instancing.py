"""Instancing on the CPU: the oracle's restatement of GeometryCollection.Trace / Occlude on
TokenType.Instance leaves + PreparedInstance.Trace / Occlude (GeometryCollection.cs:123-131,160-168, PreparedInstance.cs:47-111)
against (a) its own brute-force root loop and (b) the same scene with every placement baked into world-space primitives."""
import numpy as np
import pytest

from echorenderer_b200 import host, scenes, structs

from . import oracle_lib

EMPTY = structs.TOKEN_EMPTY


@pytest.fixture(scope="module")
def instanced():
    prepared = host.prepare(scenes.instanced_scene())
    return prepared, oracle_lib.OracleScene(prepared)


def spawn_from_hits(rays, hits, layers, seed=41):
    """Rays leaving the hit points in random directions, ignoring the full hierarchy they start on (TraceQuery.SpawnTrace)."""
    hit = hits["token"] != EMPTY
    rays, hits, layers = rays[hit], hits[hit], layers[hit]
    index = np.arange(len(rays), dtype=np.uint64)
    spawned = np.zeros(len(rays), dtype=structs.RAY)
    spawned["origin"] = rays["direction"] * np.maximum(hits["distance"], np.float32(8e-7))[:, None] + rays["origin"]
    spawned["direction"] = scenes.uniform_sphere_directions(scenes.uniform(seed, index, 0), scenes.uniform(seed, index, 1))
    spawned["distance"] = np.inf
    spawned["ignore"] = hits["token"]
    return spawned, layers.copy()


def bake(prepared):
    """Every placement flattened into world-space triangles, spheres and point lights (float64 composition of the instance
    matrices), with each placement's swatch resolved to absolute material indices."""
    triangles, spheres, points = [], [], []

    def visit(pack_index, matrix, scale, material_offset):
        pack = prepared.packs[pack_index]
        t = prepared.triangles[pack["triangleOffset"]:pack["triangleOffset"] + pack["triangleCount"]]
        if len(t):
            v0 = t["vertex0"].astype(np.float64)
            v1, v2 = v0 + t["edge1"], v0 + t["edge2"]
            world = [v @ matrix[:3, :3].T + matrix[:3, 3] for v in (v0, v1, v2)]
            normals = [t[k].astype(np.float64) @ matrix[:3, :3].T for k in ("normal0", "normal1", "normal2")]
            triangles.append(scenes.make_triangles(*world, t["material"] + material_offset, *normals))
        s = prepared.spheres[pack["sphereOffset"]:pack["sphereOffset"] + pack["sphereCount"]].copy()
        if len(s):
            s["position"] = s["position"].astype(np.float64) @ matrix[:3, :3].T + matrix[:3, 3]
            s["radius"] = s["radius"] * scale
            s["material"] += material_offset
            spheres.append(s)
        p = prepared.point_lights[pack["pointLightOffset"]:pack["pointLightOffset"] + pack["pointLightCount"]].copy()
        if len(p):
            p["position"] = p["position"].astype(np.float64) @ matrix[:3, :3].T + matrix[:3, 3]
            p["intensity"] = p["intensity"] * np.float32(scale * scale)  # same radiance at the same world distance
            points.append(p)
        for k in range(pack["instanceCount"]):
            instance = prepared.instances[pack["instanceOffset"] + k]
            local = np.eye(4)
            local[:3] = instance["inverse"].reshape(3, 4)
            visit(int(instance["pack"]), matrix @ local, scale * float(instance["inverseScale"]), int(instance["materialOffset"]))

    visit(0, np.eye(4), 1.0, 0)
    d = prepared.description
    description = host.SceneDescription(triangles=np.concatenate(triangles), spheres=np.concatenate(spheres), materials=prepared.materials,
                                        point_lights=np.concatenate(points) if points else np.zeros(0, dtype=structs.POINT_LIGHT),
                                        infinite_lights=d.infinite_lights, camera=d.camera)
    return host.prepare(description)


def test_pack_layout(instanced):
    prepared, _ = instanced
    packs, instances = prepared.packs, prepared.instances
    assert len(packs) == 3 and len(instances) == 39  # 36 placements in the scene + 3 inside the cluster pack
    assert packs[0]["nodeOffset"] == 0 and packs[0]["instanceCount"] == 36
    assert int(packs["nodeCount"].sum()) == len(prepared.nodes) and int(packs["triangleCount"].sum()) == len(prepared.triangles)
    assert np.all(instances["pack"] >= 1)

    for instance in instances:
        forward, inverse = np.eye(4), np.eye(4)
        forward[:3], inverse[:3] = instance["forward"].reshape(3, 4), instance["inverse"].reshape(3, 4)
        assert np.allclose(forward @ inverse, np.eye(4), atol=1e-5)
        assert instance["forwardScale"] * instance["inverseScale"] == pytest.approx(1.0, rel=1e-6)

    # instance tokens of every pack stay inside its own instance range
    for pack in packs:
        tokens = prepared.nodes[pack["nodeOffset"]:pack["nodeOffset"] + pack["nodeCount"]]["token4"].reshape(-1)
        tokens = tokens[tokens != EMPTY]
        kinds = structs.token_type(tokens)
        assert np.all(structs.token_index(tokens[kinds == structs.TOKEN_TYPE_INSTANCE]) < pack["instanceCount"])
        assert np.count_nonzero(kinds == structs.TOKEN_TYPE_INSTANCE) == pack["instanceCount"]


def test_tree_matches_brute_force(instanced):
    prepared, oracle = instanced
    rays = scenes.random_rays(prepared.bounds, 1 << 14, seed=3)
    hits, layers = oracle.trace_hierarchy(rays, threads=4)
    linear_hits, linear_layers = oracle.trace_hierarchy(rays, linear=True, threads=4)

    hit = hits["token"] != EMPTY
    assert 0.1 < hit.mean() < 0.9
    assert set(np.unique(layers["instanceCount"][hit])) == {0, 1, 2}
    assert np.array_equal(hits["token"], linear_hits["token"]) and np.array_equal(layers, linear_layers)
    # entering and leaving a placement rescales TraceQuery.distance (PreparedInstance.cs:51,60): the value depends on the visit order in the last bits
    assert np.allclose(hits["distance"][hit], linear_hits["distance"][hit], rtol=1e-5)

    occluded = oracle.occlude_hierarchy(rays, threads=4)
    assert np.array_equal(occluded, oracle.occlude_hierarchy(rays, linear=True, threads=4))
    assert np.array_equal(occluded != 0, hit)  # infinite travel: occluded iff something is hit


def test_matches_baked_scene(instanced):
    prepared, oracle = instanced
    baked = oracle_lib.OracleScene(bake(prepared))
    rays = scenes.random_rays(prepared.bounds, 1 << 14, seed=5)
    hits, _ = oracle.trace_hierarchy(rays, threads=4)
    flat = baked.trace(rays, threads=4)

    hit, flat_hit = hits["token"] != EMPTY, flat["token"] != EMPTY
    assert np.mean(hit != flat_hit) < 2e-3  # silhouette rays may flip: the baked vertices are rounded differently
    both = hit & flat_hit
    relative = np.abs(hits["distance"][both] - flat["distance"][both]) / np.maximum(flat["distance"][both], 1e-3)
    assert np.quantile(relative, 0.99) < 1e-4


def test_ignore_needs_the_whole_hierarchy(instanced):
    prepared, oracle = instanced
    rays = scenes.random_rays(prepared.bounds, 1 << 14, seed=7)
    hits, layers = oracle.trace_hierarchy(rays, threads=4)
    spawned, ignore = spawn_from_hits(rays, hits, layers)
    inside = ignore["instanceCount"] > 0
    assert inside.sum() > 100

    again, again_layers = oracle.trace_hierarchy(spawned, ignore, threads=4)
    triangle = structs.token_type(spawned["ignore"]) == structs.TOKEN_TYPE_TRIANGLE
    same = (again["token"] == spawned["ignore"]) & (again_layers == ignore)
    assert not np.any(same & triangle)  # query.ignore == query.current skips the triangle, GeometryCollection.cs:93

    # without the layers the ignore hierarchy is a different one: triangles inside placements are tested again and the
    # ray re-hits its own surface at distance ~0 for a good share of the queries
    naive, naive_layers = oracle.trace_hierarchy(spawned, None, threads=4)
    self_hits = (naive["token"] == spawned["ignore"]) & (naive_layers == ignore) & triangle & inside
    assert self_hits.sum() > 10
    outside = ~inside
    assert np.array_equal(naive[outside], again[outside])


def test_light_hierarchy(instanced):
    """Lights inside placements are leaves of type Instance in their parent's light tree (LightCollection.cs:123-135); Pick
    descends through them (PreparedScene.cs:132-147) and ProbabilityMass multiplies the masses layer by layer (:166-176)."""
    prepared, _ = instanced
    packs = prepared.packs
    assert packs["lightNodeCount"].tolist() == [75, 1, 7]  # 1 emissive quad + 36 lit placements (+1 sphere per blob; 3 blobs + 1 point light per cluster)
    root_tokens = prepared.emitter_tokens[:packs[0]["emitterCount"]]
    kinds = structs.token_type(root_tokens)
    assert np.count_nonzero(kinds == structs.TOKEN_TYPE_INSTANCE) == 36 and np.count_nonzero(kinds == structs.TOKEN_TYPE_TRIANGLE) == 2
    cluster_tokens = prepared.emitter_tokens[packs[2]["emitterOffset"]:packs[2]["emitterOffset"] + packs[2]["emitterCount"]]
    assert sorted(structs.token_type(cluster_tokens).tolist()) == [structs.TOKEN_TYPE_INSTANCE] * 3 + [structs.TOKEN_TYPE_LIGHT]


def test_render_matches_baked_scene():
    """The integrator through placements (Interact with FindLayer, lights picked through the hierarchy, occlusion with full
    ignore hierarchies) converges to the image of the same scene with every placement baked into world space. One instance
    layer only: with nested placements the reference composes the inverse transforms in the order of PreparedScene.cs:273,
    which does not commute with rotations, and the restatement keeps that."""
    prepared = host.prepare(scenes.instanced_scene(grid=4, rings=12, segments=14, nested=False))
    baked = bake(prepared)
    width, height = 96, 54
    tiles = scenes.tile_grid(width, height, 16)
    params = structs.render_params(width, height, 16, extend=96, seed=3, bounce_limit=6)
    image, stats = oracle_lib.OracleScene(prepared).render_tiles(params, tiles, threads=4)
    flat, flat_stats = oracle_lib.OracleScene(baked).render_tiles(params, tiles, threads=4)
    assert not np.isnan(image).any() and int(stats["lightOcclusionPassed"][0]) > 0

    a = scenes.assemble_tiles(image, tiles, width, height, 16)[..., :3]
    b = scenes.assemble_tiles(flat, tiles, width, height, 16)[..., :3]
    assert np.all(np.abs(a.mean(axis=(0, 1)) - b.mean(axis=(0, 1))) <= 0.03 * b.mean(axis=(0, 1)))
    # block means: the two images show the same picture, not just the same average
    blocks_a = a[:48, :96].reshape(6, 8, 12, 8, 3).mean(axis=(1, 3))
    blocks_b = b[:48, :96].reshape(6, 8, 12, 8, 3).mean(axis=(1, 3))
    assert np.mean(np.abs(blocks_a - blocks_b)) <= 0.06 * blocks_b.mean()
