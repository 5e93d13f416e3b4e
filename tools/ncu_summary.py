#!/usr/bin/env python3
"""Turn an ncu report (--set full) into the markdown summary committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_summary.md "title"
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads / warp instruction (warp execution efficiency x32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % of peak"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % of peak"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe % of peak"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__cycles_elapsed.avg", "SM cycles"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["no_instruction", "wait", "not_selected", "short_scoreboard", "long_scoreboard", "math_pipe_throttle",
               "branch_resolving", "dispatch_stall", "barrier", "lg_throttle", "mio_throttle", "imc_miss"]


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    names = []
    for r in data:
        n = r[ix["Kernel Name"]]
        n = n[n.find("::") + 2:] if "::" in n else n
        names.append(n.replace("<unnamed>::", "").replace("(bool)", "").split("(")[0])
    lines = [f"# {title}", "", f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`); one column per "
             "captured launch, in launch order.  Durations under ncu are serialised / cold-cache: compare shares, not absolutes.", ""]
    lines.append("| metric | " + " | ".join(f"{i}: `{n}`" for i, n in enumerate(names)) + " |")
    lines.append("|---|" + "---|" * len(names))
    for key, label in METRICS:
        if key not in ix:
            continue
        vals = []
        for r in data:
            v = r[ix[key]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:,.0f}" if abs(f) >= 1000 else f"{f:.2f}"
            except ValueError:
                pass
            vals.append(v)
        lines.append(f"| {label} [{units[ix[key]]}] | " + " | ".join(vals) + " |")
    lines += ["", "Warp stall reasons (average warps stalled per issue-active cycle; > 0.2 shown):", "",
              "| stall | " + " | ".join(str(i) for i in range(len(names))) + " |", "|---|" + "---|" * len(names)]
    for s in STALL_NAMES:
        key = STALLS % s
        if key not in ix:
            continue
        vals = [float(r[ix[key]]) for r in data]
        if max(vals) > 0.2:
            lines.append(f"| {s} | " + " | ".join(f"{v:.2f}" for v in vals) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
