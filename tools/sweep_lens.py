"""Sweep of the lensed render kernel's scheduling knobs (env: BT_SCAN_LANES, BT_SCAN_PATIENCE,
BT_COMPACT_LANES, BT_COMPACT_PATIENCE, BT_LENS_NO_SKIP) on the C3 / lensed-cloud frames
at reduced spp.  Each configuration runs in its own process (some knobs are read once).

    python tools/sweep_lens.py            # the sweep
    python tools/sweep_lens.py --one      # one measurement with the current environment
"""
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def one():
    import numpy as np
    import torch

    import bendy_tracer_b200 as bt
    scenes = os.path.join(ROOT, "tests", "golden", "scenes")
    out = []
    for name, lens, passes in (("scene", (1.362, 1.577, 6.114, 0.2), 4), ("cloud", (2.4 - 6 * 0.1705, 2.7 - 6 * 0.1908, 12.0 - 6 * 0.9667, 0.2), 4)):
        w, h = 1920, 1080
        scene = bt.Scene.load(os.path.join(scenes, name + ".json.gz"))
        cam = scene.find_by_tag("camera")
        scene.set_camera_aspect(cam, w / h)
        scene.set_lenses(np.array([lens], np.float32))
        buf = bt.Buffer(w, h, device="cuda:0")
        tracer = bt.Tracer(bt.Config(), seed=0)
        rc = bt.RenderConfig.with_samples_subsample(passes, bt.Subsample(2))
        tracer.render(scene, cam, rc, buf)
        best = 1e30
        for i in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            tracer.render(scene, cam, rc, buf, sample_base=passes * (i + 1), sync=False)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        st = tracer.render_stats(scene, cam, rc, w, h, sample_base=passes)
        out.append(f"{name}+lens {w * h * passes * 4 / best / 1e3:7.1f} Ms/s ({best:6.1f} ms; steps/path {st['rk4_steps'] / st['paths']:.1f}, "
                   f"scans/path {st['scans'] / st['paths']:.1f})")
    print(" | ".join(out), flush=True)


def main():
    if "--one" in sys.argv:
        return one()
    configs = [
        {},
        {"BT_SCAN_LANES": "4", "BT_SCAN_PATIENCE": "1"},
        {"BT_SCAN_LANES": "6", "BT_SCAN_PATIENCE": "2"},
        {"BT_SCAN_LANES": "8", "BT_SCAN_PATIENCE": "3"},
        {"BT_SCAN_LANES": "8", "BT_SCAN_PATIENCE": "4"},
        {"BT_SCAN_LANES": "12", "BT_SCAN_PATIENCE": "4"},
        {"BT_SCAN_LANES": "16", "BT_SCAN_PATIENCE": "6"},
        {"BT_SCAN_LANES": "8", "BT_SCAN_PATIENCE": "3", "BT_STEPS_PER_TURN": "3"},
        {"BT_SCAN_LANES": "8", "BT_SCAN_PATIENCE": "2", "BT_STEPS_PER_TURN": "4"},
    ]
    if "--old" in sys.argv:
        configs = [
            {"BT_LENS_NO_SKIP": "1", "BT_SCAN_LANES": "1", "BT_SCAN_PATIENCE": "1"},   # scan with every step
            {"BT_LENS_NO_SKIP": "1"},
            {"BT_SCAN_LANES": "24", "BT_SCAN_PATIENCE": "8"},
            {"BT_COMPACT_LANES": "16", "BT_COMPACT_PATIENCE": "32"},
            {"BT_COMPACT_LANES": "4", "BT_COMPACT_PATIENCE": "8"},
        ]
    for c in configs:
        env = dict(os.environ)
        env.update(c)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, capture_output=True, text=True)
        print(f"{c}: {r.stdout.strip()} {r.stderr.strip()[-300:] if r.returncode else ''}", flush=True)


if __name__ == "__main__":
    main()
