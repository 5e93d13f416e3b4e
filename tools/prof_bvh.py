"""Profiling target: one render_kernel launch over the synthetic 32k-primitive scene (BVH path)."""
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import bendy_tracer_b200 as bt
from common import synthetic_scene

scene = bt.Scene.from_json(json.dumps(synthetic_scene(20000, 6000, 1000, seed=1, extent=14.0)))
cam = scene.find_by_tag("camera")
w, h = 1920, 1080
scene.set_camera_aspect(cam, w / h)
tracer = bt.Tracer(bt.Config(), seed=0)
rc = bt.RenderConfig.with_samples_subsample(4, bt.Subsample(2))
print(scene.info())
print(tracer.render_stats(scene, cam, rc, w, h))
for rep in range(2 if "--once" not in sys.argv else 1):
    buf = bt.Buffer(w, h, device="cuda:0")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tracer.render(scene, cam, rc, buf, sync=False)
    e1.record()
    torch.cuda.synchronize()
    print(f"bvh render {w}x{h} @ 16 spp: {e0.elapsed_time(e1):.2f} ms  {w * h * 16 / e0.elapsed_time(e1) / 1e3:.1f} Msamples/s")
