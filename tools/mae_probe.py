"""MAE of engine renders against the oracle on identical sample sets, per scene (diagnostic)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bendy_tracer_b200 as bt
import oracle_ffi as O
from common import load_pair, oracle_render, engine_render, mae_per_channel, LENS_SCENE, LENS_VOLUME
w, h = 128, 72
for name, lens in (("cornell2", None), ("scene", None), ("cloud", None), ("volume", None), ("scene", LENS_SCENE), ("cloud", LENS_VOLUME)):
    osc, esc, cam = load_pair(name, w, h)
    if lens is not None:
        osc.set_lenses(lens); esc.set_lenses(lens, bt.LensConfig(exact_rsqrt=True))
    ref, n, _ = oracle_render(osc, cam, w, h, 2, 2, 0, seed=2)
    got, _, _ = engine_render(esc, cam, w, h, 2, 2, 0, seed=2)
    d = np.abs(got[..., :3] - ref[..., :3]).sum(-1) / n
    print(f"{name:9s} lens={lens is not None}  MAE {mae_per_channel(got, ref, n)}  pixels differing > 1e-3: {(d > 1e-3).mean():.5f}  > 1e-6: {(d > 1e-6).mean():.5f}")
