#!/bin/bash
# Round-end evidence on one GPU: tests, smoke, bench lines, the ncu launch list of the bench command, one --set full capture of
# the dominant kernel and its DRAM traffic at full frame size (each ncu run only after the plain command has exited 0).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
python bench.py > gpurun_out/final_bench_C3.json 2> gpurun_out/final_bench_C3.err || exit 1
python bench.py --workload C2 --steps 5 --warmup 3 --no-extras > gpurun_out/final_bench_C2.json 2> gpurun_out/final_bench_C2.err
python bench.py --workload C4-cloud --steps 5 --warmup 3 --no-extras > gpurun_out/final_bench_C4.json 2> gpurun_out/final_bench_C4.err
python bench.py --workload C4-cloud-lens --steps 3 --warmup 3 --no-extras > gpurun_out/final_bench_C4_lens.json 2> gpurun_out/final_bench_C4_lens.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
python bench.py --steps 2 --warmup 3 --no-extras > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_C3.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/final_ncu_launch.log 2>&1
python tools/prof_c3.py > gpurun_out/final_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:render_pool --launch-skip 1 --launch-count 1 -o gpurun_out/final_prof_c3 -f python tools/prof_c3.py > gpurun_out/final_ncu_full.log 2>&1
python bench.py --steps 1 --warmup 3 --no-extras > /dev/null 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:render_pool --launch-skip 3 --launch-count 1 --csv --log-file gpurun_out/final_traffic_C3.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/final_ncu_traffic.log 2>&1
cat gpurun_out/final_pytest.log gpurun_out/final_smoke.log
tail -3 gpurun_out/final_ncu_full.log
tail -5 gpurun_out/final_traffic_C3.csv
