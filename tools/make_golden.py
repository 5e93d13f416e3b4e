#!/usr/bin/env python3
"""Generates tests/golden/oracle_vectors.npz: small outputs of the CPU oracle (oracle/) for fixed
seeds -- images of the five shipped scenes (Full + the three AOVs on cornell), the lensed C3 / C4
scenes, camera rays, first-hit probe segments and stepper endpoints.

The reference ships no golden vectors and cannot be built here ("parity unpinned"), so
these vectors pin the ORACLE: tests/test_golden.py checks that the oracle still reproduces them bit
for bit (CPU) and that the CUDA engine matches them within the north-star tolerances (GPU).

    python tools/make_golden.py        # rewrites the fixture; commit the result
"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import oracle_ffi as O
from golden_cases import CASES, W, H, load_oracle_case, probe_rays, stepper_case

out = {}
for key, case in CASES.items():
    osc, cam = load_oracle_case(case)
    cfg = O.make_config(samples=case["samples"], subsample=case["subsample"], output=case["output"])
    img, n, _ = osc.render(cam, cfg, W, H, seed=case["seed"])
    assert n == case["samples"] * max(case["subsample"], 1) ** 2
    out["image/" + key] = img.astype(np.float32)
    if case["output"] == 0:
        xs, ys, pidx = probe_rays()
        rays = osc.camera_rays(cam, O.make_config(samples=1), W, H, xs, ys, pidx, seed=case["seed"])
        out["rays/" + key] = rays.astype(np.float32)
        seg = osc.probe(O.make_config(samples=1), rays[:, :3], rays[:, 3:], use_f64=False)
        for f in ("face", "steps", "object_ref", "t", "position", "normal", "direction"):
            out[f"segments/{key}/{f}"] = seg[f]
for m in (1, 4):
    lenses, xv = stepper_case(m)
    out[f"stepper/M{m}/f32"] = O.integrate(lenses, xv, 64, use_f64=False).astype(np.float32)
    out[f"stepper/M{m}/f64"] = O.integrate(lenses, xv, 64, use_f64=True).astype(np.float64)
path = os.path.join(ROOT, "tests", "golden", "oracle_vectors.npz")
np.savez_compressed(path, **out)
print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")
