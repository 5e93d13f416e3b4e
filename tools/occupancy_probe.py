"""How many resident warps per SM does the geodesic stepper need?  integrate_kernel (256-thread CTAs) is
run with extra dynamic shared memory per CTA so that only 1..8 CTAs fit an SM (8..64 warps).
Run in separate processes (the pad is read from the environment at launch time)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
import bendy_tracer_b200 as bt
e = bt.Engine.default(0)
n = 1 << 22
rng = np.random.default_rng(1234)
b = rng.uniform(2.6, 40.0, n); phi = rng.uniform(0, 2 * np.pi, n)
xv = np.zeros((n, 6), np.float32)
xv[:, 0], xv[:, 1], xv[:, 2], xv[:, 5] = b * np.cos(phi), b * np.sin(phi), 20.0, -1.0
lenses = np.array([[0, 0, 0, 1.0]], np.float32)
d = torch.from_numpy(xv).cuda()
e.geodesic_integrate(lenses, d, 16); torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    d.copy_(torch.from_numpy(xv))
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); e.geodesic_integrate(lenses, d, 256); z.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(z))
print("%%.2f ms  %%.1f TFLOP/s" %% (best, n * 256 * 214 / best / 1e9))
''' % ROOT
for ctas in (8, 6, 5, 4, 3, 2, 1):
    pad = 0 if ctas == 8 else (227 * 1024) // ctas - 1024
    env = dict(os.environ, BT_INTEGRATE_SMEM_PAD=str(pad))
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(f"CTAs/SM <= {ctas} ({ctas * 8} warps), pad {pad} B: {out.stdout.strip()} {out.stderr.strip()[-200:]}", flush=True)
