#!/usr/bin/env python3
"""Warp instructions / stall samples / active lanes of one kernel, aggregated by code region, from an
ncu report captured with --import-source on (kernels compiled -lineinfo).

    python tools/ncu_regions.py REPORT.ncu-rep KERNEL_REGEX [LAUNCH_INDEX]

Regions: every function of csrc/device.cuh (found by its BT_DEV / template header), the functions of
csrc/kernels.cu and the phases of the pooled kernel in csrc/render_pool.cuh (found by marker lines).  Inlined code is attributed to the
line it was written on, i.e. to the helper, not to its caller.
"""
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CSRC = os.path.join(ROOT, "bendy_tracer_b200", "csrc")

GROUPS = [  # (region label, function-name regex) for device.cuh
    ("vector / math helpers", r"^(v3|operator|dot|xdot|xat|cross|m_\w+|normalize_\w+|mat_vec|any_orthonormal_pair|lerpf|reflect|refract|fresnel)$"),
    ("rng (xoshiro256++, seeding, Uniform / Bernoulli)", r"^(rotl64|splitmix_mix|next_u64|next_u32|seed_from_u64|path_seed|uniform_f32|standard_f32|gen_bool|uniform_index|Rng)$"),
    ("sincos", r"^bt_sincos$"),
    ("sphere test", r"^(sphere_roots\w*|sqrt_approx|sphere_free_bound|free_update|FreeInfo)$"),
    ("rect test (general)", r"^(rect_test|rect_test_q|sdot|sat|sdiv)$"),
    ("rect test (axis-aligned)", r"^(rect_test_aa|comp|sat1)$"),
    ("box slab test", r"^(box_test|rcp_approx)$"),
    ("scan loop", r"^(scan_prims\w*|Hit)$"),
    ("bvh", r"^(slab|bvh_\w+|BvhTrav)$"),
    ("resolve_hit", r"^(resolve_hit|Surface)$"),
    ("light_pdf / light_point", r"^(light_pdf|light_point)$"),
    ("density (trilinear)", r"^density_\w+$"),
    ("camera ray", r"^(with_frustum_dir|camera_ray)$"),
    ("stepper arithmetic (lens_accel, rk4)", r"^(lens_accel|rk4_from_k1|axpy|step_size|normalize_fma|rsqrt_approx|Lens\w+|D0Cache)$"),
]
KERNEL_MARKS = [  # (first line matching -> region label) for kernels.cu, in file order
    (r"struct SceneView", "scene staging"),
    (r"struct Traced", "trace_straight / flight state"),
    (r"struct GridFetch", "free-distance grid lookup"),
    (r"BT_DEV int geodesic_step", "geodesic_step bookkeeping (capture / far / chord length / skip test)"),
    (r"BT_DEV int geodesic_scan", "geodesic_scan bookkeeping"),
    (r"BT_DEV Traced flight_result", "flight_result / trace_ray"),
    (r"struct PathQ", "path state"),
    (r"BT_DEV bool shade_event", "shade_event: classify / material"),
    (r"---- 2\. the shared direction sampler", "shade_event: direction sampler"),
    (r"---- 3\. the scattered ray", "shade_event: scatter / pdf / throughput"),
    (r"    if \(finish\) \{", "shade_event: finish"),
    (r"BT_DEV void render_body", "render_body (lane kernel) control"),
    (r"__global__ void", "other kernels"),
]
POOL_MARKS = [  # render_pool.cuh
    (r"^enum \{", "pool: layout / collect"),
    (r"BT_DEV void render_pool_body", "pool: prologue + phase selection"),
    (r"=+ STEP =+", "pool: STEP loop control + refill rounds"),
    (r"=+ NODE =+", "pool: NODE pass (BVH scenes: state load / store, queues)"),
    (r"=+ LEAF =+", "pool: LEAF pass (BVH scenes: state load / store, queues)"),
    (r"=+ SCAN =+", "pool: SCAN pass (state load / store, queues)"),
    (r"=+ SHADE =+", "pool: SHADE pass (state load / store, queues)"),
    (r"=+ REGEN =+", "pool: REGEN pass (retire, pixel stream, issue)"),
]


def device_regions():
    lines = open(os.path.join(CSRC, "device.cuh")).read().splitlines()
    starts = []
    for i, l in enumerate(lines, 1):
        m = re.match(r"\s*(?:BT_DEV|struct|template)\b.*?(\w+)\s*(?:\(|\{|$)", l)
        if l.startswith("template"):
            continue
        m = re.match(r"(?:BT_DEV\s+[\w:<> \*&]+?\s+|struct\s+)(operator\S*|\w+)\s*[\(\{<]", l)
        if m:
            name = "operator" if m.group(1).startswith("operator") else m.group(1)
            starts.append((i, name))
    regions = []
    for (ln, name), nxt in zip(starts, starts[1:] + [(len(lines) + 1, None)]):
        label = next((g for g, rx in GROUPS if re.match(rx, name)), "device.cuh: " + name)
        regions.append((ln, nxt[0] - 1, label))
    return regions


def kernel_regions(fname="kernels.cu", table=None):
    table = table or KERNEL_MARKS
    lines = open(os.path.join(CSRC, fname)).read().splitlines()
    marks, k = [], 0
    for i, l in enumerate(lines, 1):
        if k < len(table) and re.search(table[k][0], l):
            marks.append((i, table[k][1]))
            k += 1
    return [(ln, nxt[0] - 1, label) for (ln, label), nxt in zip(marks, marks[1:] + [(len(lines) + 1, None)])]


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern,
           "--launch-skip", skip, "--launch-count", "1"]
    rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True, check=True).stdout)))
    regions = {"device.cuh": device_regions(), "kernels.cu": kernel_regions(), "render_pool.cuh": kernel_regions("render_pool.cuh", POOL_MARKS)}
    agg, cur, hdr, name = {}, None, None, ""
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            name = r[1]
        elif r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
        elif hdr and r[0].isdigit():
            try:
                inst, tinst, samp = int(r[hdr["Instructions Executed"]]), int(r[hdr["Thread Instructions Executed"]]), int(r[hdr["# Samples"]])
            except (ValueError, KeyError):
                continue
            ln, label = int(r[0]), str(cur) + " (other)"
            for lo, hi, lab in regions.get(cur, []):
                if lo <= ln <= hi:
                    label = lab
            a = agg.setdefault(label, [0, 0, 0])
            a[0] += inst
            a[1] += tinst
            a[2] += samp
    tot, ts = sum(a[0] for a in agg.values()) or 1, sum(a[2] for a in agg.values()) or 1
    print(f"`{name}`: {tot:,} warp instructions\n")
    print("| region | warp instructions | stall samples | active lanes / instruction |\n|---|---|---|---|")
    for label, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if a[0] * 1000 >= tot:
            print(f"| {label} | {100 * a[0] / tot:.1f} % | {100 * a[2] / ts:.1f} % | {a[1] / max(a[0], 1):.1f} |")


if __name__ == "__main__":
    main()
