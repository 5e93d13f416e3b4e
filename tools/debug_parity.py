import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np
import oracle_ffi as O
from common import *
import bendy_tracer_b200 as bt
out = {}
for name, lens in [("scene", LENS_SCENE), ("cloud", LENS_VOLUME)]:
    w, h = 128, 72
    osc, esc, cam = load_pair(name, w, h, lenses=lens)
    ref, n, _ = oracle_render(osc, cam, w, h, 2, 2, 0, seed=2)
    got, n2, _ = engine_render(esc, cam, w, h, 2, 2, 0, seed=2)
    d = np.abs(got - ref)[..., :3].sum(-1) / n
    print(name, "mae", mae_per_channel(got, ref, n), "npix diff>1e-3:", (d > 1e-3).sum(), "max", d.max())
    ys, xs = np.nonzero(d > 1e-3)
    print(" diff pixels (x,y,d):", [(int(x), int(y), float(d[y, x])) for x, y in zip(xs[:40], ys[:40])])
    out[name + "_got"] = got; out[name + "_ref"] = ref
    # per-path first segments
    yy, xx = np.mgrid[0:h, 0:w]; xx, yy = xx.ravel(), yy.ravel()
    cfg = O.make_config(samples=1)
    rays = osc.camera_rays(cam, cfg, w, h, xx, yy, np.zeros(len(xx), np.uint64))
    r64 = osc.probe(cfg, rays[:, :3], rays[:, 3:], use_f64=True)
    r32 = osc.probe(cfg, rays[:, :3], rays[:, 3:], use_f64=False)
    g = bt.Tracer(bt.Config()).trace_segments(esc, rays[:, :3], rays[:, 3:])
    for nm, r in (("f64", r64), ("f32", r32)):
        same = (g["face"] == r["face"]) & (g["object_ref"] == r["object_ref"])
        print(" vs", nm, "same", same.mean(), "steps gpu/ref mean", g["steps"].mean(), r["steps"].mean(), "faces", np.unique(g["face"], return_counts=True))
        hit = same & (r["face"] >= 0)
        err = np.linalg.norm(g["position"][hit] - r["position"][hit], axis=1) / (np.linalg.norm(r["position"][hit], axis=1) + 1)
        print("   pos err quantiles", np.quantile(err, [0.5, 0.9, 0.99, 0.999, 1.0]))
        e = same & (r["face"] == -1)
        derr = np.linalg.norm(g["direction"][e] - r["direction"][e], axis=1)
        if e.any(): print("   dir err quantiles", np.quantile(derr, [0.5, 0.9, 0.99, 0.999, 1.0]))
        bad = np.nonzero(~same)[0][:10]
        print("   mismatches:", [(int(xx[i]), int(yy[i]), int(g["face"][i]), int(r["face"][i]), int(g["steps"][i]), int(r["steps"][i])) for i in bad])
np.savez_compressed(os.path.join(os.path.dirname(__file__), '..', 'gpurun_out', 'debug_parity.npz'), **out)
