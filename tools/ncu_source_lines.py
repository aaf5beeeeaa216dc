#!/usr/bin/env python
"""Per-source-line instruction share, active lanes and stall-sample share of one kernel from an ncu report captured with --import-source on
(`ncu -i <rep> --page source --csv --print-source cuda,sass`, read on the CPU box).

  python tools/ncu_source_lines.py <file.ncu-rep> "<substring of the kernel's function name>" [top N]
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    path, needle = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    text = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True, check=True).stdout
    sections, current = [], None
    for row in csv.reader(io.StringIO(text)):
        if not row:
            continue
        if row[0] == "File Path":
            current = {"file": row[1], "rows": [], "function": None}
            sections.append(current)
        elif row[0] == "Function Name" and current is not None:
            current["function"] = row[1]
        elif row[0] == "Line No" and current is not None:
            current["header"] = row
        elif row[0] != "Kernel Name" and current is not None and "header" in current:
            current["rows"].append(row)

    lines = collections.OrderedDict()
    for section in sections:
        if section["function"] is None or needle not in section["function"]:
            continue
        index = {}
        for i, name in enumerate(section["header"]):
            index.setdefault(name, i)
        for row in section["rows"]:
            if row[0] == "" or len(row) <= index["Thread Instructions Executed"]:
                continue
            try:
                values = [float(row[index[name]] or 0) for name in ("Instructions Executed", "Thread Instructions Executed", "# Samples")]
            except ValueError:
                continue
            entry = lines.setdefault((section["file"].split("/")[-1], int(row[0]), row[1].strip()[:100]), [0.0, 0.0, 0.0])
            for i in range(3):
                entry[i] += values[i]

    total = [sum(entry[i] for entry in lines.values()) for i in range(3)]
    print(f"{needle}: {total[0]:.4g} warp instructions, {total[1] / total[0]:.2f} active lanes on average, {int(total[2])} stall samples")
    files = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    for (name, _, _), entry in lines.items():
        for i in range(3):
            files[name][i] += entry[i]
    for name, entry in files.items():
        print(f"  {name:28s} instructions {100 * entry[0] / total[0]:5.1f} %  lanes {entry[1] / max(entry[0], 1):4.1f}  samples {100 * entry[2] / total[2]:5.1f} %")
    print()
    for (name, line, source), entry in sorted(lines.items(), key=lambda item: -item[1][0])[:top]:
        print(f"{name:22s} L{line:<4d} instructions {100 * entry[0] / total[0]:5.2f} %  lanes {entry[1] / max(entry[0], 1):4.1f}  samples {100 * entry[2] / total[2]:5.2f} % | {source[:95]}")


if __name__ == "__main__":
    main()
