#!/usr/bin/env python3
"""Per-source-line instruction counts of one kernel from an ncu report captured with
--import-source on (kernels compiled -lineinfo).

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [LAUNCH_INDEX] [TOP_N]

Prints the hottest CUDA source lines: warp instructions executed, their share of the kernel,
average active threads per instruction and stall samples.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern,
           "--launch-skip", str(skip), "--launch-count", "1"]
    raw = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    lines = []
    cur_file, hdr = None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            print("#", r[1])
        elif r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
        elif r[0] not in ("",) and hdr and r[0].isdigit():
            try:
                inst = int(r[hdr["Instructions Executed"]])
                tinst = int(r[hdr["Thread Instructions Executed"]])
                samp = int(r[hdr["# Samples"]])
            except (ValueError, KeyError):
                continue
            lines.append((inst, tinst, samp, cur_file, int(r[0]), r[1].strip()))
    total = sum(l[0] for l in lines) or 1
    tsamp = sum(l[2] for l in lines) or 1
    print(f"# total warp instructions {total:,}; samples {tsamp:,}")
    print(f"{'inst%':>6} {'samp%':>6} {'thr/inst':>8}  location")
    for inst, tinst, samp, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{100 * inst / total:6.2f} {100 * samp / tsamp:6.2f} {tinst / max(inst, 1):8.1f}  {f}:{ln}  {src[:110]}")


if __name__ == "__main__":
    main()
