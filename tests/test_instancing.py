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
    """Every placement flattened into world-space triangles and spheres (float64 composition of the instance matrices)."""
    triangles, spheres = [], []

    def visit(pack_index, matrix, scale):
        pack = prepared.packs[pack_index]
        t = prepared.triangles[pack["triangleOffset"]:pack["triangleOffset"] + pack["triangleCount"]]
        if len(t):
            v0 = t["vertex0"].astype(np.float64)
            v1, v2 = v0 + t["edge1"], v0 + t["edge2"]
            world = [v @ matrix[:3, :3].T + matrix[:3, 3] for v in (v0, v1, v2)]
            triangles.append(scenes.make_triangles(*world, 0))
        s = prepared.spheres[pack["sphereOffset"]:pack["sphereOffset"] + pack["sphereCount"]].copy()
        if len(s):
            s["position"] = s["position"].astype(np.float64) @ matrix[:3, :3].T + matrix[:3, 3]
            s["radius"] = s["radius"] * scale
            spheres.append(s)
        for k in range(pack["instanceCount"]):
            instance = prepared.instances[pack["instanceOffset"] + k]
            local = np.eye(4)
            local[:3] = instance["inverse"].reshape(3, 4)
            visit(int(instance["pack"]), matrix @ local, scale * float(instance["inverseScale"]))

    visit(0, np.eye(4), 1.0)
    description = host.SceneDescription(triangles=np.concatenate(triangles), spheres=np.concatenate(spheres),
                                        materials=scenes.material(structs.MATERIAL_DIFFUSE), camera=prepared.description.camera)
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
