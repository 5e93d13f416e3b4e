"""Short profiling target for ncu: one launch each of the render kernel (flat C2 frame, lensed C3
scene, C4 cloud volume) and the geodesic stepper kernel, at reduced spp so that ncu's ~40 replays
stay short.  Usage: python tools/prof_target.py"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bendy_tracer_b200 as bt

SCENES = os.path.join(ROOT, "tests", "golden", "scenes")


def render(name, w, h, passes, lens=None):
    scene = bt.Scene.load(os.path.join(SCENES, name + ".json.gz"))
    cam = scene.find_by_tag("camera")
    scene.set_camera_aspect(cam, w / h)
    if lens is not None:
        scene.set_lenses(np.array([lens], np.float32))
    buf = bt.Buffer(w, h, device="cuda:0")
    tracer = bt.Tracer(bt.Config(), seed=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tracer.render(scene, cam, bt.RenderConfig.with_samples_subsample(passes, bt.Subsample(2)), buf, sync=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name}{' +lens' if lens else ''}: {w}x{h} @ {passes * 4} spp  {ms:.2f} ms  {w * h * passes * 4 / ms / 1e3:.1f} Msamples/s")


def stepper(m, n=1 << 20, steps=512):
    rng = np.random.default_rng(1234)
    b, phi = rng.uniform(2.6, 40.0, n), rng.uniform(0, 2 * np.pi, n)
    xv = np.zeros((n, 6), np.float32)
    xv[:, 0], xv[:, 1], xv[:, 2], xv[:, 5] = b * np.cos(phi), b * np.sin(phi), 20.0, -1.0
    lenses = np.zeros((m, 4), np.float32)
    lenses[:, 3] = 1.0 / m
    lenses[1:, :3] = rng.uniform(-3, 3, (m - 1, 3))
    d = torch.from_numpy(xv).cuda()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    bt.Engine.default().geodesic_integrate(lenses, d, steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"stepper M={m}: {n} rays x {steps} steps  {ms:.2f} ms  {n * steps * (136 * m + 78) / ms / 1e9:.2f} TFLOP/s")


if __name__ == "__main__":
    for rep in range(2):   # first round warms up (module load, clocks); profile with --launch-skip 6
        print("warm-up round" if rep == 0 else "timed round")
        render("cornell2", 1920, 1080, 16)
        render("scene", 1920, 1080, 4, lens=(1.362, 1.577, 6.114, 0.2))
        render("cloud", 1920, 1080, 16)
        stepper(1)
        stepper(4)
        stepper(16)
