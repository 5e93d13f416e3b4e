"""ncu target: one lensed render (scene.json.gz + the C3 mass, 1920x1080 at 16 spp) after a warm-up.
Usage: ncu --launch-skip 1 --launch-count 1 ... python tools/prof_lens.py [scene|cloud] [passes]"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bendy_tracer_b200 as bt

name = sys.argv[1] if len(sys.argv) > 1 else "scene"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
lens = {"scene": (1.362, 1.577, 6.114, 0.2), "cloud": (2.4 - 6 * 0.1705, 2.7 - 6 * 0.1908, 12.0 - 6 * 0.9667, 0.2)}[name]
w, h = 1920, 1080
scene = bt.Scene.load(os.path.join(ROOT, "tests", "golden", "scenes", name + ".json.gz"))
cam = scene.find_by_tag("camera")
scene.set_camera_aspect(cam, w / h)
if "--flat" not in sys.argv:
    scene.set_lenses(np.array([lens], np.float32))
buf = bt.Buffer(w, h, device="cuda:0")
tracer = bt.Tracer(bt.Config(), seed=0)
rc = bt.RenderConfig.with_samples_subsample(passes, bt.Subsample(2))
for i in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tracer.render(scene, cam, rc, buf, sample_base=passes * i, sync=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name}: {w}x{h} @ {passes * 4} spp  {ms:.2f} ms  {w * h * passes * 4 / ms / 1e3:.1f} Msamples/s")
