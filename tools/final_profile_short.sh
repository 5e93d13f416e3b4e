#!/bin/bash
# The parts of final_profile.sh that depend on the stepper / test results: tests, smoke, the default bench line, the --set full capture.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_pytest.log; python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
python bench.py > gpurun_out/final_bench_C2.json 2> gpurun_out/final_bench_C2.err || exit 1
python tools/prof_target.py > gpurun_out/final_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:render_kernel|integrate_kernel' --launch-skip 6 --launch-count 6 -o gpurun_out/final_prof -f python tools/prof_target.py > gpurun_out/final_ncu_full.log 2>&1
cat gpurun_out/final_pytest.log
