import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
from sweep_pool import measure
for name in ("C3", "C4-cloud-lens"):
    print(f"{name} default: {measure(name, reps=3):7.1f}  lane: {measure(name, reps=2, pool_w=0):7.1f}", flush=True)
