"""Sweep of the regeneration-phase thresholds (env BT_REGEN_LANES / BT_REGEN_PATIENCE) on the shipped scenes."""
import os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

def one():
    import numpy as np, torch
    import bendy_tracer_b200 as bt
    out = []
    for name, lens, passes in (("cornell2", None, 16), ("scene", None, 16), ("cloud", None, 16), ("scene", (1.362, 1.577, 6.114, 0.2), 4)):
        w, h = 1920, 1080
        scene = bt.Scene.load(os.path.join(ROOT, "tests", "golden", "scenes", name + ".json.gz"))
        cam = scene.find_by_tag("camera")
        scene.set_camera_aspect(cam, w / h)
        if lens: scene.set_lenses(np.array([lens], np.float32))
        buf = bt.Buffer(w, h, device="cuda:0")
        tracer = bt.Tracer(bt.Config(), seed=0)
        rc = bt.RenderConfig.with_samples_subsample(passes, bt.Subsample(2))
        tracer.render(scene, cam, rc, buf)
        best = 1e30
        for i in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); tracer.render(scene, cam, rc, buf, sample_base=passes * (i + 1), sync=False); e1.record()
            torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        out.append(f"{name}{'+lens' if lens else ''} {w * h * passes * 4 / best / 1e3:7.1f}")
    print(" | ".join(out), flush=True)

if "--one" in sys.argv:
    one()
else:
    for lanes, pat in ((1, 1), (4, 4), (8, 8), (8, 16), (12, 16), (16, 16), (16, 32), (24, 32)):
        env = dict(os.environ, BT_REGEN_LANES=str(lanes), BT_REGEN_PATIENCE=str(pat))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, capture_output=True, text=True)
        print(f"regen lanes {lanes:2d} patience {pat:2d}: {r.stdout.strip()} {r.stderr.strip()[-200:] if r.returncode else ''}", flush=True)
