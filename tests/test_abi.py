"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/*.h declares, keeps the POD layouts of
the reference's structs, and refuses compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from echorenderer_b200 import _native, structs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(echo_(?:b200|host)_\w+)\s*\(", text))


def have_gpu():
    try:
        return _native.device_count() > 0
    except _native.EchoNativeError:
        return False


def test_library_exports_every_declared_symbol():
    lib = _native.library()
    declared = declared_functions("echo_b200.h") | declared_functions("echo_b200_debug.h")
    assert declared == set(_native.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.echo_b200_version()


def test_host_library_exports_every_declared_symbol():
    from echorenderer_b200 import host
    lib = host._library()
    for name in declared_functions("echo_host.h"):
        assert hasattr(lib, name), name


# every POD of include/echo_b200.h and its numpy mirror in echorenderer_b200/structs.py
POD_MIRRORS = {
    "EchoQbvhNode": "QBVH_NODE", "EchoTriangle": "TRIANGLE", "EchoSphere": "SPHERE", "EchoRay": "RAY", "EchoHit": "HIT", "EchoMaterial": "MATERIAL",
    "EchoTexture": "TEXTURE", "EchoMaterialTextures": "MATERIAL_TEXTURES", "EchoLightNode": "LIGHT_NODE", "EchoPointLight": "POINT_LIGHT",
    "EchoInfiniteLight": "INFINITE_LIGHT", "EchoPack": "PACK", "EchoInstance": "INSTANCE", "EchoTokenHierarchy": "TOKEN_HIERARCHY", "EchoCamera": "CAMERA",
    "EchoRenderParams": "RENDER_PARAMS", "EchoStats": "STATS",
}


def header_structs():
    text = open(os.path.join(ROOT, "include", "echo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"typedef struct (\w+)\s*\{", text))


def test_pod_layouts_match_the_header():
    """sizeof and the offsetof of EVERY member of EVERY struct of the header, recomputed by compiling it (gcc), equal the numpy
    mirrors; no struct of the header is left without a mirror; every struct carries its size assertion in the header."""
    import subprocess
    import tempfile
    assert header_structs() == set(POD_MIRRORS), "a struct of echo_b200.h has no mirror in this test"
    header = open(os.path.join(ROOT, "include", "echo_b200.h")).read()

    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "echo_b200.h"', "int main(void) {"]
    for c_name, mirror in POD_MIRRORS.items():
        assert f"ECHO_B200_ASSERT_SIZE({c_name}, {getattr(structs, mirror).itemsize});" in header, f"{c_name} has no size assertion in the header"
        lines.append(f'printf("{c_name} %zu\\n", sizeof({c_name}));')
        for name in getattr(structs, mirror).names:
            lines.append(f'printf("{c_name}.{name} %zu %zu\\n", offsetof({c_name}, {name}), sizeof((({c_name}*)0)->{name}));')
    lines.append("return 0; }")

    with tempfile.TemporaryDirectory() as directory:
        source, binary = os.path.join(directory, "layout.c"), os.path.join(directory, "layout")
        open(source, "w").write("\n".join(lines))
        subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), source, "-o", binary], check=True)
        output = subprocess.run([binary], check=True, capture_output=True, text=True).stdout

    checked = 0
    for line in output.splitlines():
        key, *values = line.split()
        c_name, _, member = key.partition(".")
        dtype = getattr(structs, POD_MIRRORS[c_name])
        if member:
            assert int(values[0]) == dtype.fields[member][1] and int(values[1]) == dtype.fields[member][0].itemsize, line
        else:
            assert int(values[0]) == dtype.itemsize, line
        checked += 1
    assert checked == len(POD_MIRRORS) + sum(len(getattr(structs, m).names) for m in POD_MIRRORS.values())

    # the members of every C struct are all mirrored (the mirrors tile their structs without gaps)
    for mirror in POD_MIRRORS.values():
        dtype = getattr(structs, mirror)
        assert sum(dtype.fields[name][0].itemsize for name in dtype.names) == dtype.itemsize, mirror

    # the sizes the reference's structs have (a1, a9, a10)
    assert structs.QBVH_NODE.itemsize == 128 and structs.TRIANGLE.itemsize == 100 and structs.SPHERE.itemsize == 20


def test_integration_document_states_the_header_sizes():
    """INTEGRATION.md's C# binding must declare every struct at the size the header asserts (a stale size there corrupts memory
    in a host that follows the document)."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for c_name, mirror in POD_MIRRORS.items():
        size = getattr(structs, mirror).itemsize
        assert re.search(rf"Size\s*=\s*{size}\)\]\s*(?:public\s+)?(?:unsafe\s+)?struct\s+{c_name[4:]}\b", text), f"{c_name}: INTEGRATION.md does not declare it with Size = {size}"


def test_plain_c_client_compiles_without_warnings(tmp_path):
    """tests/c_client/echo_client.c (the non-Python, non-C++ host of the GPU suite) builds as C11 without a warning against the header alone (dlsym results become function pointers, which -pedantic alone objects to)."""
    import subprocess
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_client", "echo_client.c"),
                    "-o", str(tmp_path / "echo_client"), "-ldl"], check=True)


def test_integration_structs_are_the_generated_block():
    """The struct block of INTEGRATION.md is exactly what tools/gen_csharp_structs.py derives from the checked layouts."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_csharp_structs", os.path.join(ROOT, "tools", "gen_csharp_structs.py"))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    assert {pod[0] for pod in module.PODS} == set(POD_MIRRORS)
    assert module.BEGIN + "\n" + module.block() + "\n" + module.END in open(os.path.join(ROOT, "INTEGRATION.md")).read()


@pytest.mark.skipif(have_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_a_device(cornell):
    """Every compute entry point fails loudly (ECHO_B200_ERR_NO_DEVICE) instead of falling back to the CPU."""
    from echorenderer_b200 import PreparedScene
    lib = _native.library()
    count = ctypes.c_int32(-1)
    assert lib.echo_b200_device_count(ctypes.byref(count)) == _native.ERR_NO_DEVICE and count.value == 0
    assert b"CUDA" in lib.echo_b200_last_error()

    handle = ctypes.c_void_p()
    assert lib.echo_b200_scene_create(ctypes.byref(handle), 0) == _native.ERR_NO_DEVICE and not handle.value

    with pytest.raises(_native.EchoNativeError) as error:
        PreparedScene(cornell)
    assert error.value.status == _native.ERR_NO_DEVICE

    out = np.zeros(8, dtype=np.float32)
    zeros = np.zeros(8, dtype=np.float32)
    assert lib.echo_b200_debug_math(0, 0, _native.pointer(zeros), _native.pointer(zeros), _native.pointer(zeros), 8, _native.pointer(out)) == _native.ERR_NO_DEVICE

    # the device-side builds have no host twin inside this library either (the host mirror is another library, libecho_host.so)
    from echorenderer_b200 import build_light_tree_device, build_qbvh_device
    for build in (lambda: build_qbvh_device(cornell.triangles, cornell.spheres), lambda: build_light_tree_device(cornell.description)):
        with pytest.raises(_native.EchoNativeError) as error:
            build()
        assert error.value.status == _native.ERR_NO_DEVICE


def test_null_scene_is_rejected():
    lib = _native.library()
    assert lib.echo_b200_scene_commit(None) == _native.ERR_INVALID
    assert lib.echo_b200_trace_batch(None, None, 0, None) == _native.ERR_INVALID
    assert lib.echo_b200_render_tiles(None, None, None, 0, None, None) == _native.ERR_INVALID
    assert lib.echo_b200_scene_destroy(None) == _native.OK


def _run_bench(arguments, environment=None):
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    result = subprocess.run([sys.executable, os.path.join(root, "bench.py")] + arguments, capture_output=True, text=True, timeout=600,
                            env={**os.environ, **(environment or {})})
    assert result.returncode == 0, result.stderr[-2000:]
    lines = [line for line in result.stdout.splitlines() if line.strip()]
    return [json.loads(line) for line in lines]


@pytest.mark.parametrize("workload", ["trace", "render"])
def test_bench_reference_arm_contract(workload):
    """`bench.py --impl reference` (runs on the host cores, no GPU): ONE JSON line on stdout with the keys of the measurement
    contract, `cpu_baseline` and `e2e` describing this very run; under torchrun every rank but 0 exits 0 without a line."""
    arguments = ["--impl", "reference", "--steps", "1", "--warmup", "1"]
    arguments += ["--quads", "64", "32", "--rays", "65536", "--cpu-sample", "65536"] if workload == "trace" else \
        ["--workload", "render", "--scene", "cornell", "--width", "512", "--height", "512", "--spp", "2"]
    lines = _run_bench(arguments)
    assert len(lines) == 1
    line = lines[0]
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "impl", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["warmup"] >= 3 and line["vs_baseline"] is None and line["gpu_launches"] == 0
    assert line["unit"] == ("Mrays/s" if workload == "trace" else "samples/s") and "workload" in line["config"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert _run_bench(arguments, {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []


def test_cpp_host_mirror_compiles_without_warnings_and_its_host_logic_holds(tmp_path):
    """include/echo_b200.hpp — the compiled-language host side above the C ABI (PreparedScene, EvaluationProfile, RenderTexture, tile
    patterns, IWorker, EvaluationOperation + Factory): builds as C++17 with -Wall -Wextra -Werror, its tile patterns equal the Python
    mirror's (which restates the reference's TilePatternTests), and the device-free logic checks of tests/c_client/echo_host_logic.cpp pass."""
    import subprocess
    import numpy as np
    from echorenderer_b200 import hilbert_curve_pattern, ordered_pattern
    package = os.path.join(ROOT, "echorenderer_b200")
    flags = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include")]
    link = ["-L", package, "-lecho_b200", f"-Wl,-rpath,{package}", "-pthread"]
    logic, client = tmp_path / "echo_host_logic", tmp_path / "echo_host_client"
    subprocess.run(flags + [os.path.join(ROOT, "tests", "c_client", "echo_host_logic.cpp"), "-o", str(logic)] + link, check=True)
    subprocess.run(flags + [os.path.join(ROOT, "tests", "c_client", "echo_host_client.cpp"), "-o", str(client)] + link, check=True)

    result = subprocess.run([str(logic), "checks"], capture_output=True, text=True, timeout=120)
    assert result.returncode == 0 and "echo_host_logic ok" in result.stdout, result.stdout + result.stderr

    for size in [(1, 1), (2, 1), (1, 9), (5, 3), (8, 8), (13, 7), (120, 68)]:
        for hilbert, expected in ((1, hilbert_curve_pattern(size)), (0, ordered_pattern(size))):
            printed = subprocess.run([str(logic), "pattern", str(hilbert), str(size[0]), str(size[1])], capture_output=True, text=True, check=True).stdout
            assert np.array_equal(np.array(printed.split(), dtype=np.int32).reshape(-1, 2), expected), (size, hilbert)
