"""ncu target: the default (pooled) render of the C3 scene, 1920x1080 at PROF_SPP (default 16) spp, twice
(the first launch is the warm-up: ncu --launch-skip 1 --launch-count 1 -k regex:render_pool)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bendy_tracer_b200 as bt  # noqa: E402
from bench import SCENE_DIR, WORKLOADS  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
scene_name, w, h, _, sub, lens = WORKLOADS[name]
w, h = 1920, 1080
passes = int(os.environ.get("PROF_SPP", "16")) // (sub * sub)
scene = bt.Scene.load(os.path.join(SCENE_DIR, scene_name + ".json.gz"))
cam = scene.find_by_tag("camera")
scene.set_camera_aspect(cam, float(np.float32(w) / np.float32(h)))
if lens:
    scene.set_lenses(np.array([lens], np.float32))
tr = bt.Tracer(bt.Config(), seed=0)
rc = bt.RenderConfig.with_samples_subsample(passes, bt.Subsample(sub))
buf = bt.Buffer(w, h, device="cuda:0")
for i in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tr.render(scene, cam, rc, buf, sample_base=7, sync=False)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1):.2f} ms  {w * h * passes * sub * sub / e0.elapsed_time(e1) / 1e3:.1f} Msamples/s", flush=True)
