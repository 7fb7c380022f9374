#!/usr/bin/env python
"""Summarise an .ncu-rep (captured on the B200 box under gpurun) into a small text file for profiles/.

    python profiles/summarize.py gpurun_out/prof_encode_r1c.ncu-rep profiles/r1c_encode_kernel.txt [symbols_per_launch]

symbols_per_launch = KV elements one launch codes (2^31 for the 8192-token block): the summary then states
warp-instructions per warp-symbol step = instructions each lane spends per symbol.

Reads the report with `ncu -i ... --page raw --csv` / `--page source --csv` (no GPU needed)."""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    # the coder kernels are bounded by the ALU pipe (one warp instruction per 2 cycles per SMSP) and by issue slots
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "not_selected", "barrier", "math_pipe_throttle",
          "branch_resolving", "mio_throttle", "lg_throttle", "no_instruction", "dispatch_stall", "sleeping"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    nsym = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = ncu_csv(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    lines = [f"source report: {rep}", f"kernel: {d.get('Kernel Name', ('?', ''))[0]}", ""]
    for k in KEYS:
        if k in d:
            lines.append(f"{k:75s} {d[k][0]} {d[k][1]}")
    lines.append("")
    lines.append("warp stall reasons (warps stalled per issue-active cycle):")
    for s_ in STALLS:
        k = f"smsp__average_warps_issue_stalled_{s_}_per_issue_active.ratio"
        if k in d:
            lines.append(f"  {s_:22s} {float(d[k][0]):.3f}")
    src = ncu_csv(rep, "source")
    if len(src) > 2:
        h = src[1]
        ix = {n: i for i, n in enumerate(h)}
        rows = src[2:]
        tot = sum(int(r[ix["Instructions Executed"]] or 0) for r in rows)
        ops = Counter()
        for r in rows:
            t = r[ix["Source"]].split()
            if not t:
                continue
            op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            ops[op.split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
        lines += ["", f"SASS: {len(rows)} instructions, {tot} warp-instructions executed"]
        if nsym:
            lines.append(f"warp-instructions per warp-symbol step (= per 32 symbols, one per lane): {tot / (nsym / 32.0):.1f}")
        lines.append("opcode mix: " + ", ".join(f"{o} {100 * n / tot:.1f}%" for o, n in ops.most_common(14)))
        mn = {o for o in ops}
        lines.append("tensor / TMA mnemonics present: " + (", ".join(sorted(m for m in mn if m.startswith(("UTC", "UTMA", "UBLKCP", "HMMA", "LDTM")))) or "none (integer/byte path, by design)"))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
