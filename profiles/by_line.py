#!/usr/bin/env python
"""Warp-instructions per source line from an .ncu-rep captured with --import-source on (kernels built with -lineinfo).

    python profiles/by_line.py gpurun_out/prof_enc.ncu-rep [symbols_per_launch] [top_n]

Prints, for the first kernel in the report, the source lines ordered by executed warp-instructions and -- when
symbols_per_launch is given -- the same figure per warp-symbol step (instructions each lane spends per symbol)."""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    nsym = float(sys.argv[2]) if len(sys.argv) > 2 else None
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    per = defaultdict(lambda: [0, "", 0])     # (file, line) -> [warp instr, text, shared wavefronts]
    cur_file = ""
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = {n: i for i, n in enumerate(r)}
            continue
        if hdr is None or r[0] in ("Function Name", "Kernel Name") or not r[0].isdigit():
            continue
        try:
            n = int(r[hdr["Instructions Executed"]])
        except (ValueError, KeyError):
            continue
        k = (cur_file, int(r[0]))
        per[k][0] += n
        per[k][1] = r[1].strip()
        try:
            per[k][2] += int(r[hdr["L1 Wavefronts Shared"]])
        except (ValueError, KeyError):
            pass
    tot = sum(v[0] for v in per.values())
    print(f"total warp-instructions attributed to source lines: {tot}")
    step = nsym / 32.0 if nsym else None
    for (f, ln), (n, text, wf) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        extra = f"  {n / step:6.2f}/sym  smem-wavefronts {wf / step:5.2f}/sym" if step else ""
        print(f"{n:12d} {100.0 * n / tot:5.1f}%{extra}  {f}:{ln}  {text[:90]}")


if __name__ == "__main__":
    main()
