#!/usr/bin/env python
"""Summaries of the ncu evidence for profiles/ (read on the CPU box, no GPU needed).

  python tools/ncu_summary.py launches <ncu --csv log> [--steps N]   per-kernel launch count, device time and share of the list
  python tools/ncu_summary.py traffic  <ncu --csv log>               + DRAM bytes per kernel (dram__bytes_read / _write .sum)
  python tools/ncu_summary.py report   <file.ncu-rep>                key `--set full` metrics of every captured launch

The launch lists come from `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --clock-control none --csv --log-file ...`
(B200_PROFILING.md): per-launch times are cold-cache and serialised, so SHARES are what compares with the CUDA-event timings.
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys


def read_csv(path):
    lines = [line for line in open(path, errors="replace") if not line.startswith("==")]
    return list(csv.DictReader(io.StringIO("".join(lines))))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("echo::", "")
    return name[:90]


def launches(path, with_traffic):
    rows = read_csv(path)
    per_launch = collections.defaultdict(dict)
    for row in rows:
        per_launch[(row["ID"], row["Kernel Name"])][row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    kernels = collections.OrderedDict()
    for (_, name), metrics in per_launch.items():
        entry = kernels.setdefault(short(name), {"launches": 0, "ns": 0.0, "read": 0.0, "write": 0.0})
        entry["launches"] += 1
        entry["ns"] += metrics.get("gpu__time_duration.sum", 0.0)
        entry["read"] += metrics.get("dram__bytes_read.sum", 0.0)
        entry["write"] += metrics.get("dram__bytes_write.sum", 0.0)
    total = sum(entry["ns"] for entry in kernels.values()) or 1.0
    out = {"source": path, "launches": len(per_launch), "total_ms": total / 1e6, "kernels": {}}
    for name, entry in sorted(kernels.items(), key=lambda item: -item[1]["ns"]):
        record = {"launches": entry["launches"], "ms": entry["ns"] / 1e6, "share": entry["ns"] / total}
        if with_traffic:
            record["dram_read_mb"] = entry["read"] / 1e6
            record["dram_write_mb"] = entry["write"] / 1e6
        out["kernels"][name] = record
    if with_traffic:
        out["dram_bytes_total"] = sum(entry["read"] + entry["write"] for entry in kernels.values())
    return out


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.avg.per_cycle_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
        "launch__grid_size", "launch__block_size"]


def report(path):
    text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    header, units, data = rows[0], rows[1], rows[2:]
    index = {name: i for i, name in enumerate(header)}
    out = []
    for row in data:
        record = {"kernel": short(row[index["Kernel Name"]])}
        for key in KEYS:
            if key in index:
                value = row[index[key]].replace(",", "")
                try:
                    record[key] = float(value)
                except ValueError:
                    record[key] = value
                record.setdefault("units", {})[key] = units[index[key]]
        out.append(record)
    return out


def main():
    mode, path = sys.argv[1], sys.argv[2]
    if mode == "launches":
        print(json.dumps(launches(path, False), indent=1))
    elif mode == "traffic":
        print(json.dumps(launches(path, True), indent=1))
    else:
        print(json.dumps(report(path), indent=1))


if __name__ == "__main__":
    main()
