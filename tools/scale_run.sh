#!/bin/bash
# Strong scaling of the C5 frame on one 8 x B200 box: C5-64 at N = 1, 2, 4, 8 (torchrun, NCCL), N = 8 in one process through
# bt_engine_create_multi, and the full BASELINE configs[4] frame (1024 spp) once at N = 8.
mkdir -p gpurun_out
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --steps 5 --warmup 3 --no-extras > gpurun_out/final_scale_n$n.json 2> gpurun_out/final_scale_n$n.err
done
python bench.py --workload C5-64 --steps 3 --warmup 3 --no-extras > gpurun_out/final_scale_n1.json 2> gpurun_out/final_scale_n1.err
python bench.py --gpus 8 --steps 5 --warmup 3 --no-extras > gpurun_out/final_scale_n8_inproc.json 2> gpurun_out/final_scale_n8_inproc.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --workload C5 --steps 2 --warmup 1 --no-extras > gpurun_out/final_scale_C5_full_n8.json 2> gpurun_out/final_scale_C5_full_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/final_scale_ref_n8.json 2> gpurun_out/final_scale_ref_n8.err
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -2
python - <<PY
import json
for f in ("final_scale_n1","final_scale_n2","final_scale_n4","final_scale_n8","final_scale_n8_inproc","final_scale_C5_full_n8","final_scale_ref_n8"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, {k:d.get(k) for k in ("value","ms_per_step","n_gpus","reduce_ms")}, "e2e", d["e2e"]["value"], d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
