"""ncu target: warm-up render + one measured render of a shipped scene at 1920x1080.
    python tools/prof_one.py SCENE PASSES [--lens] [--precision fast|exact|auto]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np, torch
import bendy_tracer_b200 as bt
name, passes = sys.argv[1], int(sys.argv[2])
LENS = {"scene": (1.362, 1.577, 6.114, 0.2), "cloud": (2.4 - 6 * 0.1705, 2.7 - 6 * 0.1908, 12.0 - 6 * 0.9667, 0.2),
        "volume": (2.4 - 6 * 0.1705, 2.7 - 6 * 0.1908, 12.0 - 6 * 0.9667, 0.2), "cornell2": (0.3, 2.2, 2.0, 0.1)}
w, h = 1920, 1080
scene = bt.Scene.load(os.path.join(ROOT, "tests", "golden", "scenes", name + ".json.gz"))
cam = scene.find_by_tag("camera")
scene.set_camera_aspect(cam, w / h)
if "--lens" in sys.argv:
    scene.set_lenses(np.array([LENS[name]], np.float32))
if "--precision" in sys.argv:
    scene.set_precision(sys.argv[sys.argv.index("--precision") + 1])
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
