"""Host-buffer render (bt_render, BT_MEM_HOST) against the number of pipeline bands (tuning knob host_bands)."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch
import bendy_tracer_b200 as bt

for name, w, h, passes in (("cornell2", 1920, 1080, 64), ("cornell", 512, 512, 4), ("scene", 3840, 2160, 16)):
    scene = bt.Scene.load(os.path.join(ROOT, "tests", "golden", "scenes", name + ".json.gz"))
    cam = scene.find_by_tag("camera")
    scene.set_camera_aspect(cam, w / h)
    tracer = bt.Tracer(bt.Config(), seed=0)
    rc = bt.RenderConfig.with_samples_subsample(passes, bt.Subsample(2))
    host = torch.zeros((h, w, 4), dtype=torch.float32).pin_memory()
    hb = bt.Buffer(w, h)
    hb.data = host.numpy()
    dev = bt.Buffer(w, h, device="cuda:0")
    tracer.render(scene, cam, rc, dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tracer.render(scene, cam, rc, dev, sync=False); e1.record(); torch.cuda.synchronize()
    line = [f"{name} {w}x{h}x{passes * 4}: device {e0.elapsed_time(e1):7.2f} ms"]
    for bands in ("1", "2", "4", "8", "16", ""):
        bt.Engine.default(0).set_tuning(host_bands=int(bands) if bands else None)   # (knobs are read per engine, not per call)
        tracer.render(scene, cam, rc, hb)
        best = 1e9
        for i in range(3):
            t0 = time.perf_counter(); tracer.render(scene, cam, rc, hb); best = min(best, time.perf_counter() - t0)
        line.append(f"bands {bands or 'auto'} {best * 1e3:7.2f}")
    print(" | ".join(line), flush=True)
