"""Throughput of the pooled render kernel against its scheduling knobs (in-process: bt_engine_set_tuning).

    python tools/sweep_pool.py [workload ...]      # default: C3 C2 C4-cloud
"""
import itertools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bendy_tracer_b200 as bt  # noqa: E402
from bench import SCENE_DIR, WORKLOADS  # noqa: E402

eng = bt.Engine.default(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")


def measure(name, reps=2, **knobs):
    scene_name, w, h, passes, sub, lens = WORKLOADS[name]
    scene = bt.Scene.load(os.path.join(SCENE_DIR, scene_name + ".json.gz"))
    cam = scene.find_by_tag("camera")
    scene.set_camera_aspect(cam, float(np.float32(w) / np.float32(h)))
    if lens:
        scene.set_lenses(np.array([lens], np.float32))
    eng.set_tuning(**knobs)
    tr = bt.Tracer(bt.Config(), engine=eng, seed=0)
    rc = bt.RenderConfig.with_samples_subsample(passes, bt.Subsample(sub))
    buf = bt.Buffer(w, h, device="cuda:0")
    tr.render(scene, cam, rc, buf)
    best = 1e30
    for i in range(reps):
        buf.clear()
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tr.render(scene, cam, rc, buf, sample_base=(i + 1) * passes, sync=False)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    eng.set_tuning(**{k: None for k in knobs})
    return w * h * passes * max(sub, 1) ** 2 / best / 1e3


if __name__ == "__main__":
    names = sys.argv[1:] or ["C3", "C2", "C4-cloud"]
    for name in names:
        print(f"{name}: lane kernel (pool_w=0) {measure(name, pool_w=0):9.1f} Msamples/s", flush=True)
        for w, threads in itertools.product((2, 3, 4, 5), (64, 128)):
            try:
                print(f"{name}: pool_w={w} threads={threads:3d} {measure(name, pool_w=w, pool_threads=threads):9.1f}", flush=True)
            except Exception as e:   # the pool does not fit shared memory
                print(f"{name}: pool_w={w} threads={threads}: {e}", flush=True)
        if WORKLOADS[name][5]:
            for w, refill, smin in itertools.product((3, 4), (3, 6, 9), (24, 28, 32)):
                print(f"{name}: pool_w={w} refill={refill:2d} step_min={smin} {measure(name, pool_w=w, pool_refill=refill, pool_step_min=smin):9.1f}", flush=True)
