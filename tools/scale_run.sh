#!/bin/bash
# 8-GPU box: multi-GPU correctness check, weak scaling of the C2 step, strong scaling of the C5 frame.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
$TR --master-port 29511 --nproc-per-node 8 tools/check_multi_gpu.py > gpurun_out/mgpu_check8.log 2>&1
$TR --master-port 29512 --nproc-per-node 8 bench.py --gpus 8 --steps 3 --warmup 3 --no-extras > gpurun_out/scale_s3_C2_8.json 2> gpurun_out/scale_s3_C2_8.err
python bench.py --gpus 1 --workload C5-128 --steps 1 --warmup 1 --no-extras > gpurun_out/scale_s3_C5q_1.json 2> gpurun_out/scale_s3_C5q_1.err
port=29520
for n in 2 4 8; do
  port=$((port+1))
  $TR --master-port $port --nproc-per-node $n bench.py --gpus $n --workload C5-128 --steps 1 --warmup 1 --no-extras > gpurun_out/scale_s3_C5q_$n.json 2> gpurun_out/scale_s3_C5q_$n.err
done
$TR --master-port 29530 --nproc-per-node 8 bench.py --gpus 8 --workload C5 --steps 1 --warmup 1 --no-extras > gpurun_out/scale_s3_C5_8.json 2> gpurun_out/scale_s3_C5_8.err
cat gpurun_out/mgpu_check8.log | tail -4
for f in gpurun_out/scale_s3_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(d['n_gpus'], d['scaling'], round(d['value'],1), 'Ms/s', round(d['ms_per_step'],1), 'ms', d['config']['workload'][:60])
except Exception as e: print('ERR', e)
"; done
