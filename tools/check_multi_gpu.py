"""torchrun --nproc-per-node N tools/check_multi_gpu.py : spp-sharded render + one NCCL reduce must equal the
1-GPU render of the same global pass range (same sample set, f32 summation order aside)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist

import bendy_tracer_b200 as bt

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "scenes")
ok = True
for name, lens in (("cornell", None), ("cloud", None), ("scene", (1.362, 1.577, 6.114, 0.2))):
    w, h, passes = 256, 144, 8
    scene = bt.Scene.load(os.path.join(root, name + ".json.gz"))
    cam = scene.find_by_tag("camera")
    scene.set_camera_aspect(cam, w / h)
    if lens:
        scene.set_lenses(np.array([lens], np.float32))
    tracer = bt.Tracer(bt.Config(), engine=bt.Engine.default(local), seed=3)
    rc = bt.RenderConfig.with_samples_subsample(passes, bt.Subsample(2))
    sharded = bt.Buffer(w, h, device=f"cuda:{local}")
    bt.render_sharded(tracer, scene, cam, rc, sharded, all_ranks=True)
    single = bt.Buffer(w, h, device=f"cuda:{local}")
    tracer.render(scene, cam, rc, single, sample_base=0)
    diff = (sharded.data - single.data).abs().max().item()
    scale = single.data[..., :3].abs().max().item()
    good = diff <= 2e-5 * scale and sharded.samples() == single.samples() == passes * 4
    ok &= good
    if rank == 0:
        print(f"{name}: world {world} max|sharded - single| = {diff:.3e} (scale {scale:.3e}) samples {sharded.samples()} -> {'ok' if good else 'FAIL'}")
if rank == 0:
    print("ALL OK" if ok else "FAILED")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
