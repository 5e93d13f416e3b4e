#!/bin/bash
# Round-end evidence: tests, bench lines, the ncu launch list of the bench command and one --set full
# capture per dominant kernel (each only after the plain command has exited 0).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_pytest.log; python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
python bench.py > gpurun_out/final_bench_C2.json 2> gpurun_out/final_bench_C2.err || exit 1
python bench.py --workload C3 --steps 2 --warmup 3 > gpurun_out/final_bench_C3.json 2> gpurun_out/final_bench_C3.err
python bench.py --workload C4-cloud --steps 3 --warmup 3 --no-extras > gpurun_out/final_bench_C4.json 2> gpurun_out/final_bench_C4.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
python bench.py --steps 2 --warmup 3 --no-extras > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_C2.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/final_ncu_launch.log 2>&1
python tools/prof_target.py > gpurun_out/final_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:render_kernel|integrate_kernel' --launch-skip 6 --launch-count 6 -o gpurun_out/final_prof -f python tools/prof_target.py > gpurun_out/final_ncu_full.log 2>&1
tail -3 gpurun_out/final_ncu_full.log
cat gpurun_out/final_pytest.log
python tools/prof_one.py cloud 16 > gpurun_out/final_plain_cloud.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:render_kernel' --launch-skip 1 --launch-count 1 -o gpurun_out/final_prof_cloud -f python tools/prof_one.py cloud 16 > gpurun_out/final_ncu_cloud.log 2>&1
