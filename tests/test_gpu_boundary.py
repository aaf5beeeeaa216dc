"""The drop-in boundary exercised the way a foreign host would (round 2):

* a plain-C client (tests/c_client/echo_client.c, gcc, no Python, no C++) dlopen()s libecho_b200.so, uploads the Cornell box
  from raw arrays and calls trace_batch / occlude_batch / render_tiles — what a .NET host does through [DllImport];
* counted passes (ECHO_EVALUATOR_COUNT_VISITS) return the oracle's visit counters and leave the results untouched;
* a scene without materials serves trace batches but is refused by the render entry points (api.cu:446);
* scenes on two devices driven from two threads of one process (needs two GPUs; skipped on a one-GPU box)."""
import os
import subprocess
import threading

import numpy as np
import pytest

from echorenderer_b200 import PreparedScene, _native, host, scenes, structs
from tests import oracle_lib

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_client_matches_the_python_binding_and_the_oracle(cornell, tmp_path):
    binary = tmp_path / "echo_client"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_client", "echo_client.c"), "-o", str(binary), "-ldl"], check=True)

    d = cornell.description
    rays = scenes.random_rays(cornell.bounds, 50000, seed=5)
    shadow = scenes.random_rays(cornell.bounds, 50000, seed=6, occlusion=True)
    tiles = scenes.tile_grid(64, 48, 16)
    params = structs.render_params(64, 48, 16, extend=4, min_epoch=2, max_epoch=2, bounce_limit=32, seed=9)
    scalars = np.zeros(4, dtype=np.uint32)
    scalars[0] = cornell.max_depth
    scalars[1:].view(np.float32)[:] = [cornell.infinite_threshold, cornell.infinite_pdf, cornell.bound_radius]

    for name, array in [("nodes", cornell.nodes), ("triangles", cornell.triangles), ("spheres", cornell.spheres), ("materials", cornell.materials),
                        ("light_nodes", cornell.light_nodes), ("emitter_tokens", cornell.emitter_tokens), ("emitter_paths", cornell.emitter_bitpaths),
                        ("point_lights", cornell.point_lights), ("infinite", d.infinite_lights), ("camera", d.camera), ("scalars", scalars),
                        ("rays", rays), ("shadow", shadow), ("tiles", tiles.astype(np.int32)), ("params", params)]:
        np.ascontiguousarray(array).tofile(tmp_path / f"{name}.bin")

    library = os.environ.get("ECHO_B200_LIBRARY") or os.path.join(ROOT, "echorenderer_b200", "libecho_b200.so")
    result = subprocess.run([str(binary), library, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert result.returncode == 0, result.stderr
    assert "echo_client ok" in result.stdout

    hits = np.fromfile(tmp_path / "hits.bin", dtype=structs.HIT)
    occluded = np.fromfile(tmp_path / "occluded.bin", dtype=np.uint8)
    image = np.fromfile(tmp_path / "image.bin", dtype=np.float32).reshape(len(tiles), 16, 16, 4)
    stats = np.fromfile(tmp_path / "stats.bin", dtype=structs.STATS)

    oracle = oracle_lib.OracleScene(cornell)
    expected_hits = oracle.trace(rays)
    assert np.array_equal(hits["token"], expected_hits["token"]) and np.array_equal(hits.view(np.uint32), expected_hits.view(np.uint32))
    assert np.array_equal(occluded, oracle.occlude(shadow))

    with PreparedScene(cornell) as scene:
        through_ctypes, ctypes_stats = scene.render_tiles(params, tiles)
    assert np.array_equal(image.view(np.uint32), through_ctypes.view(np.uint32))
    expected_image, expected_stats = oracle.render_tiles(params, tiles)
    assert np.sqrt(np.mean((image - expected_image) ** 2)) <= 1e-4 * np.sqrt(np.mean(expected_image ** 2))
    for name in structs.STATS_FIELDS[:12]:
        assert int(stats[name][0]) == int(ctypes_stats[name][0]) == int(expected_stats[name][0]), name


def test_cpp_host_mirror_renders_what_the_python_mirror_renders(cornell, tmp_path):
    """tests/c_client/echo_host_client.cpp, a host written against include/echo_b200.hpp: three worker threads enter
    EvaluationOperation::Execute concurrently and claim procedures atomically (Operation.cs:164-177); the frame they assemble through
    RenderTexture.Apply, the statistics rows and TotalSamples equal the Python mirror's EvaluationOperation bit for bit, with the tree
    given or built by the library; an abort requested through IWorker.CheckSchedule stops the workers between tiles."""
    from echorenderer_b200 import EvaluationOperation, EvaluationProfile, PathTracedEvaluator, RenderTexture
    package = os.path.join(ROOT, "echorenderer_b200")
    binary = tmp_path / "echo_host_client"
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_client", "echo_host_client.cpp"), "-o", str(binary),
                    "-L", package, "-lecho_b200", f"-Wl,-rpath,{package}", "-pthread"], check=True)

    d = cornell.description
    rays = scenes.random_rays(cornell.bounds, 20000, seed=5)
    shadow = scenes.random_rays(cornell.bounds, 20000, seed=6, occlusion=True)
    width, height, tile = 200, 136, 8  # 25 x 17 = 425 tiles = 2 procedures, ragged on neither edge ... and 136 = 17 * 8
    scalars = np.zeros(4, dtype=np.uint32)
    scalars[0] = cornell.max_depth
    scalars[1:].view(np.float32)[:] = [cornell.infinite_threshold, cornell.infinite_pdf, cornell.bound_radius]
    profile = EvaluationProfile(PathTracedEvaluator(bounce_limit=16), extend=4, min_epoch=1, max_epoch=3, noise_threshold=0.3, seed=7)

    def run(directory, workers, hilbert, build_on_device, abort_after=-1, size=(width, height, tile)):
        directory.mkdir()
        words = np.zeros(9, dtype=np.uint32)
        words[[0, 1, 2, 3, 5, 7]] = [profile.evaluator.code, profile.extend, profile.min_epoch, profile.max_epoch, profile.evaluator.bounce_limit, profile.seed]
        words[[4, 6]] = np.array([profile.noise_threshold, profile.evaluator.survivability], dtype=np.float32).view(np.uint32)
        words[8:].view(np.int32)[:] = abort_after
        for name, array in [("nodes", cornell.nodes), ("triangles", cornell.triangles), ("spheres", cornell.spheres), ("materials", cornell.materials),
                            ("light_nodes", cornell.light_nodes), ("emitter_tokens", cornell.emitter_tokens), ("emitter_paths", cornell.emitter_bitpaths),
                            ("point_lights", cornell.point_lights), ("infinite", d.infinite_lights), ("camera", d.camera), ("scalars", scalars),
                            ("rays", rays), ("shadow", shadow), ("profile", words), ("size", np.array(size, dtype=np.int32))]:
            np.ascontiguousarray(array).tofile(directory / f"{name}.bin")
        result = subprocess.run([str(binary), str(directory), str(workers), str(int(hilbert)), str(int(build_on_device))], capture_output=True, text=True, timeout=300)
        assert result.returncode == 0 and "echo_host_client ok" in result.stdout, result.stdout + result.stderr
        summary = dict(line.split(" ", 1) for line in (directory / "summary.txt").read_text().splitlines())
        frame = np.fromfile(directory / "frame.bin", dtype=np.float32).reshape(size[1], size[0], 4)
        return frame, np.fromfile(directory / "stats.bin", dtype=structs.STATS), np.fromfile(directory / "sequence.bin", dtype=np.int32).reshape(-1, 2), summary, directory

    with PreparedScene(cornell) as scene:
        operation = EvaluationOperation(scene, profile, RenderTexture(width, height, tile))
        expected = operation.execute().pixels.copy()
        ragged = EvaluationOperation(scene, profile, RenderTexture(100, 70, 16))  # 7 x 5 tiles, the last column 4 wide, the top row 6 high
        expected_ragged = ragged.execute().pixels.copy()

    frame, stats, sequence, summary, directory = run(tmp_path / "given", 3, True, False)
    assert np.array_equal(sequence, operation.tile_positions) and len(sequence) == 425
    assert summary["procedures"] == "2 of 2" and summary["aborted_workers"] == "0" and summary["gpu_count"] == "1"
    assert np.array_equal(frame.view(np.uint32), expected.view(np.uint32))
    assert int(summary["total_samples"]) == operation.total_samples == int(stats["sampleEvaluated"][0])
    for name in structs.STATS_FIELDS[:12]:
        assert int(stats[name][0]) == int(operation.statistics[name][0]), name
    oracle = oracle_lib.OracleScene(cornell)
    assert np.array_equal(np.fromfile(directory / "hits.bin", dtype=structs.HIT).view(np.uint32), oracle.trace(rays).view(np.uint32))
    assert np.array_equal(np.fromfile(directory / "occluded.bin", dtype=np.uint8), oracle.occlude(shadow))

    built, _, _, _, _ = run(tmp_path / "built", 2, True, True)  # the library builds the SweepBuilder's tree itself: the same frame
    assert np.array_equal(built.view(np.uint32), expected.view(np.uint32))

    edge, _, _, _, _ = run(tmp_path / "ragged", 1, True, False, size=(100, 70, 16))
    assert np.array_equal(edge.view(np.uint32), expected_ragged.view(np.uint32))

    partial, _, _, summary, _ = run(tmp_path / "aborted", 2, False, False, abort_after=100)  # CheckSchedule throws after 100 tiles
    assert int(summary["aborted_workers"]) >= 1 and summary["procedures"] != "2 of 2"
    lit = np.any(partial != 0, axis=-1)  # a black pixel of an applied tile is indistinguishable from a pixel never written: count tiles, compare lit pixels
    applied = lit.reshape(height // tile, tile, width // tile, tile).any(axis=(1, 3)).sum()
    assert 100 <= applied < (width // tile) * (height // tile)
    assert np.array_equal(partial[lit].view(np.uint32), expected[lit].view(np.uint32))


@pytest.mark.parametrize("fixture,bounce_limit", [("cornell", 32), ("lights_small", 16), ("mixed_small", 8)])
def test_counted_pass_reports_the_oracles_visit_counters(fixture, bounce_limit, request):
    """ECHO_EVALUATOR_COUNT_VISITS: same tiles bit for bit, and node / triangle / sphere / light-node visit totals equal to the
    counters instrumented into the oracle's traversal and light tree (the inputs of bench.py's algorithmic bytes per sample)."""
    prepared = request.getfixturevalue(fixture)
    width, height = 64, 48
    tiles = scenes.tile_grid(width, height, 16)
    plain = structs.render_params(width, height, 16, extend=4, bounce_limit=bounce_limit, seed=11)
    counted = structs.render_params(width, height, 16, extend=4, bounce_limit=bounce_limit, seed=11, evaluator=structs.EVALUATOR_PATH_TRACED | structs.EVALUATOR_COUNT_VISITS)

    with PreparedScene(prepared) as scene:
        image, stats = scene.render_tiles(plain, tiles)
        counted_image, counted_stats = scene.render_tiles(counted, tiles)

    assert np.array_equal(image.view(np.uint32), counted_image.view(np.uint32))
    assert all(int(stats[name][0]) == 0 for name in structs.STATS_VISIT_FIELDS)

    _, expected = oracle_lib.OracleScene(prepared).render_tiles(plain, tiles)
    for name in structs.STATS_FIELDS[:12]:
        assert int(counted_stats[name][0]) == int(stats[name][0]), name
    for name in structs.STATS_VISIT_FIELDS:
        assert int(counted_stats[name][0]) == int(expected[name][0]), name
    assert int(counted_stats["nodeVisits"][0]) > 0 and int(counted_stats["triangleVisits"][0]) > 0
    if fixture == "lights_small":
        assert int(counted_stats["lightNodeVisits"][0]) > int(counted_stats["lightSampled"][0])  # a 300-light tree is several levels deep


def test_a_scene_without_materials_traces_but_does_not_render(terrain_small):
    lib = _native.library()
    import ctypes
    handle = ctypes.c_void_p()
    _native.check(lib.echo_b200_scene_create(ctypes.byref(handle), 0))
    try:
        ptr = _native.pointer
        _native.check(lib.echo_b200_scene_set_qbvh(handle, ptr(terrain_small.nodes), len(terrain_small.nodes), terrain_small.max_depth))
        _native.check(lib.echo_b200_scene_set_triangles(handle, ptr(terrain_small.triangles), len(terrain_small.triangles)))
        _native.check(lib.echo_b200_scene_set_spheres(handle, ptr(terrain_small.spheres), len(terrain_small.spheres)))
        _native.check(lib.echo_b200_scene_commit(handle))

        rays = scenes.random_rays(terrain_small.bounds, 4096, seed=3)
        hits = np.empty(len(rays), dtype=structs.HIT)
        _native.check(lib.echo_b200_trace_batch(handle, ptr(rays), len(rays), ptr(hits)))
        assert np.array_equal(hits.view(np.uint32), oracle_lib.OracleScene(terrain_small).trace(rays).view(np.uint32))

        params = structs.render_params(32, 32, 16, extend=1)
        tiles = scenes.tile_grid(32, 32, 16).astype(np.int32)
        out = np.zeros((len(tiles), 16, 16, 4), dtype=np.float32)
        stats = np.zeros(1, dtype=structs.STATS)
        assert lib.echo_b200_render_tiles(handle, ptr(params), ptr(tiles), len(tiles), ptr(out), ptr(stats)) == _native.ERR_INVALID
        assert b"no materials" in lib.echo_b200_last_error()
        pixel = np.zeros((1, 2), dtype=np.int32)
        index = np.zeros(1, dtype=np.uint32)
        rgb = np.zeros(3, dtype=np.float32)
        assert lib.echo_b200_debug_evaluate_samples(handle, ptr(params), ptr(pixel), ptr(index), 1, ptr(rgb)) == _native.ERR_INVALID
    finally:
        lib.echo_b200_scene_destroy(handle)


def test_many_outstanding_launches_on_many_streams(terrain_small):
    """ADVICE r1 (trace.cu:172): the work counters of persistent launches are per (device, stream) and reset by the kernels
    themselves — 600 asynchronous launches over three caller streams (more than the old 256-slot ring) all answer like one."""
    import torch
    rays = scenes.random_rays(terrain_small.bounds, 1 << 16, seed=21)
    expected = oracle_lib.OracleScene(terrain_small).trace(rays)
    device = torch.device("cuda", 0)

    with PreparedScene(terrain_small) as scene:
        d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(device)
        streams = [torch.cuda.Stream(device=device) for _ in range(3)]
        outputs = [[torch.zeros(len(rays) * 16, dtype=torch.uint8, device=device) for _ in range(4)] for _ in streams]
        torch.cuda.synchronize()
        for round_index in range(200):
            for stream, buffers in zip(streams, outputs):
                scene.trace_device(d_rays.data_ptr(), len(rays), buffers[round_index % 4].data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        for buffers in outputs:
            for buffer in buffers:
                assert np.array_equal(buffer.cpu().numpy().view(np.uint32), expected.view(np.uint32).reshape(-1))


def device_total():
    try:
        return _native.device_count()
    except _native.EchoNativeError:
        return 0


@pytest.mark.skipif(device_total() < 2, reason="needs two GPUs")
def test_two_devices_from_one_process(cornell, terrain_small):
    """One process, two scenes on two devices, two host threads (the reference is one process with N workers,
    Common/Compute/Device.cs:20-24): per-device launch state (resident grids, work counters) must not leak between devices."""
    oracle = oracle_lib.OracleScene(terrain_small)
    rays = scenes.random_rays(terrain_small.bounds, 1 << 18, seed=31)
    expected = oracle.trace(rays)
    params = structs.render_params(64, 64, 16, extend=4, bounce_limit=16, seed=2)
    tiles = scenes.tile_grid(64, 64, 16)
    expected_image, _ = oracle_lib.OracleScene(cornell).render_tiles(params, tiles)
    results = {}

    def work(device):
        with PreparedScene(terrain_small, device=device) as scene, PreparedScene(cornell, device=device) as box:
            for _ in range(8):
                hits = scene.trace(rays)
                image, _ = box.render_tiles(params, tiles)
            results[device] = (hits, image)

    threads = [threading.Thread(target=work, args=(device,)) for device in (0, 1)]
    for thread in threads:
        thread.start()
    for thread in threads:
        thread.join()

    for device in (0, 1):
        hits, image = results[device]
        assert np.array_equal(hits.view(np.uint32), expected.view(np.uint32))
        assert np.sqrt(np.mean((image - expected_image) ** 2)) <= 1e-4 * np.sqrt(np.mean(expected_image ** 2))


def test_multi_device_scene_with_one_device_is_a_single_device_scene(cornell):
    """echo_b200_scene_create_multi with a one-bit mask behaves like echo_b200_scene_create; bad masks are refused."""
    import ctypes
    lib = _native.library()
    handle = ctypes.c_void_p()
    assert lib.echo_b200_scene_create_multi(ctypes.byref(handle), 0) == _native.ERR_INVALID
    assert lib.echo_b200_scene_create_multi(ctypes.byref(handle), 1 << 40) == _native.ERR_INVALID and not handle.value

    params = structs.render_params(48, 48, 16, extend=4, bounce_limit=16, seed=4)
    tiles = scenes.tile_grid(48, 48, 16)
    with PreparedScene(cornell, devices=[0]) as scene, PreparedScene(cornell) as single:
        assert scene.gpu_count == 1 and single.gpu_count == 1
        image, _ = scene.render_tiles(params, tiles)
        expected, _ = single.render_tiles(params, tiles)
    assert np.array_equal(image.view(np.uint32), expected.view(np.uint32))


@pytest.mark.skipif(device_total() < 2, reason="needs two GPUs")
def test_multi_device_scene_fans_out_and_matches_one_device(cornell, terrain_small):
    """One handle, every device of the box (SURVEY.md 8b's device_mask): batches are cut into one range per device, tiles are
    dealt in blocks of 64 — results equal the single-device ones bit for bit, the statistics add up, and the entry points that
    take device memory refuse a multi-device scene."""
    import torch
    devices = list(range(min(device_total(), 8)))
    rays = scenes.random_rays(terrain_small.bounds, (1 << 20) + 12345, seed=41)
    shadow = scenes.random_rays(terrain_small.bounds, (1 << 20) + 12345, seed=42, occlusion=True)

    with PreparedScene(terrain_small) as single, PreparedScene(terrain_small, devices=devices) as multi:
        assert multi.gpu_count == len(devices)
        assert np.array_equal(multi.trace(rays).view(np.uint32), single.trace(rays).view(np.uint32))
        assert np.array_equal(multi.occlude(shadow), single.occlude(shadow))
        d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to("cuda:0")
        d_hits = torch.empty(len(rays) * 16, dtype=torch.uint8, device="cuda:0")
        with pytest.raises(_native.EchoNativeError) as error:
            multi.trace_device(d_rays.data_ptr(), len(rays), d_hits.data_ptr())
        assert error.value.status == _native.ERR_INVALID and "single-device" in str(error.value)

    # 40 x 23 tiles = 920 tiles = 15 blocks of 64 (the last one short), a ragged frame, two epochs
    width, height = 636, 360
    params = structs.render_params(width, height, 16, extend=2, min_epoch=2, max_epoch=2, bounce_limit=16, seed=8)
    from echorenderer_b200 import hilbert_curve_pattern
    tiles = hilbert_curve_pattern(((width + 15) // 16, (height + 15) // 16))
    with PreparedScene(cornell) as single, PreparedScene(cornell, devices=devices) as multi:
        expected, expected_stats = single.render_tiles(params, tiles)
        image, stats = multi.render_tiles(params, tiles)
    assert np.array_equal(image.view(np.uint32), expected.view(np.uint32))
    for name in structs.STATS_FIELDS[:12]:
        assert int(stats[name][0]) == int(expected_stats[name][0]), name


def test_hand_derived_known_answers():
    """The pencil-and-paper known answers of tests/test_independent_kats.py (derived from the C# alone, not from the oracle) through
    the C ABI on the device: triangle / sphere hits with exact distances and barycentrics, the ignore rule, the strict comparisons,
    and BoxBound4.Intersect's asymmetric treatment of rays grazing a box face (SSE min / max return the second operand on NaN)."""
    from tests import test_independent_kats as kats
    prepared = kats.kat_scene()
    with PreparedScene(prepared) as scene:
        kats.check_trace(scene.trace(kats.trace_rays(kats.TRACE_CASES)))
        kats.check_occlude(scene.occlude(kats.trace_rays(kats.OCCLUDE_CASES)))
        # the same queries in a batch large enough for the persistent kernels' work replacement (every lane a different case)
        rays = np.tile(kats.trace_rays(kats.TRACE_CASES), 3000)
        hits = scene.trace(rays)
        kats.check_trace(hits[len(kats.TRACE_CASES) * 1500:len(kats.TRACE_CASES) * 1501])
        shadow = np.tile(kats.trace_rays(kats.OCCLUDE_CASES), 3000)
        flags = scene.occlude(shadow).reshape(3000, len(kats.OCCLUDE_CASES))
        assert np.all(flags == flags[0]) and kats.check_occlude(flags[0]) is None
        tiled = hits.reshape(3000, len(kats.TRACE_CASES))
        assert np.all(tiled.view(np.uint32) == tiled[0].view(np.uint32))


def test_host_memory_entry_points(terrain_small):
    """echo_b200_host_alloc / _register: page-locked host memory for the host-buffer entry points (what a P/Invoke caller needs, since
    a `fixed`-pinned managed array is pageable for CUDA). Results do not depend on where the buffers live."""
    import ctypes
    lib = _native.library()
    rays = scenes.random_rays(terrain_small.bounds, 1 << 16, seed=51)
    expected = oracle_lib.OracleScene(terrain_small).trace(rays)

    with PreparedScene(terrain_small) as scene:
        locked_rays, locked_hits = _native.HostBuffer(len(rays), structs.RAY), _native.HostBuffer(len(rays), structs.HIT)
        locked_rays.array[:] = rays
        scene.trace_pointers(locked_rays.address, len(rays), locked_hits.address)
        assert np.array_equal(locked_hits.array.view(np.uint32), expected.view(np.uint32))
        locked_rays.free()
        locked_hits.free()
        locked_hits.free()  # idempotent

        own_rays, own_hits = rays.copy(), np.empty(len(rays), dtype=structs.HIT)
        _native.host_register(own_rays)
        _native.host_register(own_hits)
        try:
            scene.trace_pointers(own_rays.ctypes.data, len(rays), own_hits.ctypes.data)
        finally:
            _native.host_unregister(own_rays)
            _native.host_unregister(own_hits)
        assert np.array_equal(own_hits.view(np.uint32), expected.view(np.uint32))

    pointer = ctypes.c_void_p()
    assert lib.echo_b200_host_alloc(None, 16) == _native.ERR_INVALID
    assert lib.echo_b200_host_register(None, 16) == _native.ERR_INVALID
    assert lib.echo_b200_host_free(None) == _native.OK and lib.echo_b200_host_unregister(None) == _native.OK
    assert lib.echo_b200_host_alloc(ctypes.byref(pointer), 0) == _native.OK and pointer.value  # an empty buffer is still a buffer
    assert lib.echo_b200_host_free(pointer) == _native.OK


def test_runtime_options(cornell):
    """echo_b200_debug_set_option: unknown names are refused; the wavefront's switches change how a render is scheduled, never its result."""
    lib = _native.library()
    assert lib.echo_b200_debug_set_option(b"NO_SUCH_SWITCH", 1) == _native.ERR_INVALID
    assert lib.echo_b200_debug_set_option(None, 1) == _native.ERR_INVALID
    params = structs.render_params(64, 64, 16, extend=4, bounce_limit=24, seed=6)
    tiles = scenes.tile_grid(64, 64, 16)
    defaults = {"RUN_AHEAD": -1, "BLOCKING_SYNC": -1, "RENDER_WORKERS": 8, "BATCH_PATHS": 1 << 24, "TAIL_LIMIT": 8192, "NARROW_LIMIT": 1 << 20}

    with PreparedScene(cornell) as scene:
        reference, reference_stats = scene.render_tiles(params, tiles)
        try:
            for switches in ({"RUN_AHEAD": 3}, {"BLOCKING_SYNC": 1, "RUN_AHEAD": 2}, {"BLOCKING_SYNC": 2}, {"BLOCKING_SYNC": 2, "RUN_AHEAD": 1}, {"RENDER_WORKERS": 1},
                             {"TAIL_LIMIT": 0}, {"TAIL_LIMIT": 1 << 30}, {"NARROW_LIMIT": 0}, {"NARROW_LIMIT": 1 << 30}, {"BATCH_PATHS": 1 << 16}):
                for name, value in {**defaults, **switches}.items():
                    _native.set_option(name, value)
                image, stats = scene.render_tiles(params, tiles)
                assert np.array_equal(image.view(np.uint32), reference.view(np.uint32)), switches
                for name in structs.STATS_FIELDS[:12]:
                    assert int(stats[name][0]) == int(reference_stats[name][0]), (switches, name)
        finally:
            for name, value in defaults.items():
                _native.set_option(name, value)


def test_concurrent_calls_on_one_handle(cornell, terrain_small):
    """Echo's workers all enter Operation.Execute at once (Operation.cs:164-177): concurrent render / batch calls on ONE scene handle take turns
    on its scratch and every caller gets the single-threaded answer."""
    params = structs.render_params(64, 64, 16, extend=4, bounce_limit=16, seed=12)
    tiles = scenes.tile_grid(64, 64, 16)
    rays = scenes.random_rays(cornell.bounds, 1 << 16, seed=61)

    with PreparedScene(cornell) as scene:
        expected_image, _ = scene.render_tiles(params, tiles)
        expected_hits = scene.trace(rays)
        results, errors = {}, []

        def work(index):
            try:
                for _ in range(3):
                    image, _ = scene.render_tiles(params, tiles)
                    hits = scene.trace(rays)
                results[index] = (image, hits)
            except Exception as error:  # noqa: BLE001
                errors.append(error)

        threads = [threading.Thread(target=work, args=(index,)) for index in range(6)]
        for thread in threads:
            thread.start()
        for thread in threads:
            thread.join()

    assert not errors, errors
    for image, hits in results.values():
        assert np.array_equal(image.view(np.uint32), expected_image.view(np.uint32))
        assert np.array_equal(hits.view(np.uint32), expected_hits.view(np.uint32))
