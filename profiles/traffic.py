#!/usr/bin/env python
"""Collect per-launch DRAM traffic and instruction counts of the codec kernels from ncu --set full captures into the JSON
that bench.py reads for roofline.traffic (so the bench line can quote them with their source; they are NOT measured in
the bench run itself).

    python profiles/traffic.py profiles/r2_traffic.json tokens chunk data coder  name=report.ncu-rep [name=report ...]

name is the key bench.py looks up: absmax_kernel, encode_kernel, decode_kernel, compact_kernel.  Capture the reports
with `bench.py --wave 32` so that one launch covers the whole block (a launch then is a step)."""
import csv
import io
import json
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return {h: v for h, v in zip(rows[0], rows[2])}, {h: u for h, u in zip(rows[0], rows[1])}


def to_bytes(val, unit):
    v = float(val)
    return int(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1))


def main():
    dst, tokens, chunk, data, coder = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
    out = {"workload": {"tokens": tokens, "chunk": chunk, "data": data, "coder": coder},
           "source": "ncu --set full --clock-control none, one launch per kernel covering the whole block (bench.py --wave 32)"}
    for arg in sys.argv[6:]:
        name, rep = arg.split("=", 1)
        d, u = raw(rep)
        out[name] = {
            "report": rep, "kernel": d.get("Kernel Name"),
            "duration_ms_under_ncu": round(float(d["gpu__time_duration.sum"]) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}[u["gpu__time_duration.sum"]], 4),
            "dram_read_bytes": to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]),
            "dram_write_bytes": to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"]),
            "warp_inst_executed": int(float(d["smsp__inst_executed.sum"])),
            "ncu_alu_pipe_pct": round(float(d.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "nan")), 1),
            "ncu_issue_active_pct": round(float(d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "nan")), 1),
        }
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
