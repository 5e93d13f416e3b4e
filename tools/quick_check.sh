#!/bin/bash
# A short GPU check: the parity tests, smoke(), the default bench line and the C3 / C4 lines.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/quick_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/quick_smoke.log 2>&1
python bench.py > gpurun_out/quick_bench_C2.json 2> gpurun_out/quick_bench_C2.err
python bench.py --workload C3 --steps 2 --warmup 3 --no-extras > gpurun_out/quick_bench_C3.json 2> gpurun_out/quick_bench_C3.err
python bench.py --workload C4-cloud --steps 3 --warmup 3 --no-extras > gpurun_out/quick_bench_C4.json 2> gpurun_out/quick_bench_C4.err
python tools/prof_target.py > gpurun_out/quick_plain.log 2>&1
cat gpurun_out/quick_pytest.log gpurun_out/quick_smoke.log
python - <<'PY'
import json
for w in ("C2", "C3", "C4"):
    try:
        d = json.loads(open(f"gpurun_out/quick_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, round(d["value"], 1), "Ms/s", round(d["ms_per_step"], 2), "ms", "e2e", d.get("e2e", {}).get("value"))
    except Exception as e:
        print(w, "failed", e)
PY
cat gpurun_out/quick_plain.log | tail -7
