"""ncu target: the lane kernel and the pooled kernel on the same frames (1920x1080 slices of C3 / C2 / C4-cloud).
One warm-up render per (workload, kernel) first, then the profiled ones in the order printed:
    ncu --set full --import-source on -k regex:render --launch-skip 6 --launch-count 6 python tools/prof_pool.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bendy_tracer_b200 as bt  # noqa: E402
from bench import SCENE_DIR, WORKLOADS  # noqa: E402

eng = bt.Engine.default(0)
jobs = []
for name, passes in (("C3", 4), ("C2", 8), ("C4-cloud", 8)):
    scene_name, w, h, _, sub, lens = WORKLOADS[name]
    w, h = 1920, 1080
    scene = bt.Scene.load(os.path.join(SCENE_DIR, scene_name + ".json.gz"))
    cam = scene.find_by_tag("camera")
    scene.set_camera_aspect(cam, float(np.float32(w) / np.float32(h)))
    if lens:
        scene.set_lenses(np.array([lens], np.float32))
    jobs.append((name, scene, cam, bt.RenderConfig.with_samples_subsample(passes, bt.Subsample(sub)), bt.Buffer(w, h, device="cuda:0"), passes * sub * sub))
tr = bt.Tracer(bt.Config(), engine=eng, seed=0)
pool_w = int(os.environ.get("PROF_POOL_W", "3"))
for timed in (False, True):
    for name, scene, cam, rc, buf, spp in jobs:
        for w_ in (0, pool_w):
            eng.set_tuning(pool_w=w_)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            tr.render(scene, cam, rc, buf, sample_base=7, sync=False)
            e1.record()
            torch.cuda.synchronize()
            if timed:
                ms = e0.elapsed_time(e1)
                print(f"{name} pool_w={w_}: {ms:.2f} ms  {1920 * 1080 * spp / ms / 1e3:.1f} Msamples/s", flush=True)
