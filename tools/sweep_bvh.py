"""Sweep of the lane kernel's BVH phase-compaction knobs on the synthetic 32k-primitive scene."""
import itertools
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import bendy_tracer_b200 as bt
from common import synthetic_scene

w, h = 1920, 1080
doc = json.dumps(synthetic_scene(20000, 6000, 1000, seed=1, extent=14.0))


eng = bt.Engine.default(0)
scene = bt.Scene.from_json(doc)
cam = scene.find_by_tag("camera")
scene.set_camera_aspect(cam, w / h)
rc = bt.RenderConfig.with_samples_subsample(4, bt.Subsample(2))


def run(**knobs):
    eng.set_tuning(**knobs)
    tracer = bt.Tracer(bt.Config(), seed=0, engine=eng)
    best = 1e9
    for rep in range(3):
        buf = bt.Buffer(w, h, device="cuda:0")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tracer.render(scene, cam, rc, buf, sync=False)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    eng.set_tuning(**{k: None for k in knobs})
    return w * h * 16 / best / 1e3


grid = sys.argv[1:] or ["default"]
print("default", f"{run():.1f} Msamples/s", flush=True)
if "sweep" in grid:
    for cl, cp, rl, rp in itertools.product((8, 16, 24), (32, 96, 256), (8, 16), (16, 64)):
        print(f"compact {cl}/{cp} regen {rl}/{rp}: {run(compact_lanes=cl, compact_patience=cp, regen_lanes=rl, regen_patience=rp):.1f}", flush=True)
if "quick" in grid:
    for cl, cp, rl, rp in ((16, 32, 8, 16), (16, 32, 4, 16), (16, 32, 6, 8), (12, 32, 4, 8), (16, 48, 2, 4)):
        print(f"compact {cl}/{cp} regen {rl}/{rp}: {run(compact_lanes=cl, compact_patience=cp, regen_lanes=rl, regen_patience=rp):.1f}", flush=True)
if "pool" in grid:
    for pw, turn in itertools.product((1, 2, 3), (1, 2, 3, 4)):
        try:
            print(f"pool_w {pw} steps_per_turn {turn}: {run(pool_w=pw, steps_per_turn=turn):.1f}", flush=True)
        except Exception as e:
            print(f"pool_w {pw}: {e}", flush=True)
