"""The measurement contract of bench.py on a GPU, at a size that takes seconds: ONE JSON line with the roofline, cpu_baseline, e2e,
clocks and gpu_launches objects the driver reads (the CPU suite checks the reference arm, tests/test_abi.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(arguments):
    result = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + arguments, capture_output=True, text=True, timeout=900)
    assert result.returncode == 0, result.stderr[-2000:]
    lines = [line for line in result.stdout.splitlines() if line.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def check_common(line):
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["warmup"] >= 3 and line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["gpu_launches"] > 0 and "workload" in line["config"] and "l2" in line["config"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in line["roofline"], key
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in line["e2e"], key
    for key in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert key in line["clocks"], key


def test_trace_bench_line():
    rays = 1 << 18
    line = run_bench(["--steps", "2", "--quads", "128", "64", "--rays", str(rays), "--cpu-sample", "65536", "--no-secondary"])
    check_common(line)
    assert line["unit"] == "Mrays/s" and line["scaling"] == "weak" and line["dtype"] == "f32"
    roofline = line["roofline"]
    assert roofline["bound"] == "hbm" and roofline["unit"] == "GB/s" and roofline["achieved"] > 0
    assert roofline["frac"] == pytest.approx(roofline["achieved"] / roofline["peak"], rel=1e-6)
    assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 2 * rays * 32 and line["e2e"]["d2h_bytes_per_step"] == rays * 16 + rays
    assert line["gpu_launches"] == 2 * line["steps"]
    baseline = line["cpu_baseline"]
    assert baseline["kind"] == "port" and baseline["cores"] >= 1 and baseline["value"] > 0 and baseline["unit"] == "Mrays/s" and baseline["sample"]


def test_render_bench_line():
    line = run_bench(["--workload", "render", "--scene", "cornell", "--width", "256", "--height", "256", "--spp", "4", "--steps", "1", "--bounce-limit", "8",
                      "--no-cpu-baseline"])
    check_common(line)
    assert line["unit"] == "samples/s" and line["scaling"] == "strong"
    assert line["e2e"]["value"] > 0 and line["e2e"]["d2h_bytes_per_step"] == 256 * 256 * 16
    assert line["stats_last_step"]["Sample/Evaluated"] == 256 * 256 * 4
