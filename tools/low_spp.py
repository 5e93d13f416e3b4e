"""Throughput against samples per pixel per call: what one rank of a strong-scaled frame sees (the C5 frame,
7680x4320, at 64 / 32 / 16 / 8 / 4 spp per GPU).  Lane kernel (pool_w=0) against the pooled kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
from sweep_pool import measure  # noqa: E402

for passes in (16, 8, 4, 2, 1):
    bench.WORKLOADS["tmp"] = ("scene", 7680, 4320, passes, 2, (1.362, 1.577, 6.114, 0.2))
    row = [f"{passes * 4:3d} spp"]
    for w in (0, 3):
        row.append(f"pool_w={w}: {measure('tmp', reps=2, pool_w=w):7.1f}")
    print("  ".join(row), flush=True)
