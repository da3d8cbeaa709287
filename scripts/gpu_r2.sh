#!/bin/bash
# Round-2 GPU check: build, smoke, GPU parity tests, B=1 / B=8 bench (no CPU leg).  Usage: bash scripts/gpu_r2.sh [tag]
set -u
TAG=${1:-r2}
O=gpurun_out/$TAG
mkdir -p $O
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" | tee $O/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
tail -n 3 $O/smoke.log
timeout 900 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -n 15 $O/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --steps 100 > $O/bench_b1.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python - <<PY
import json
for f in ("bench_b1.json",):
    try:
        d = json.loads(open("$O/" + f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches/step", d["gpu_launches"] / d["steps"])
        print(d["kernel_families_us"])
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 300 python bench.py --chunks-per-gpu 8 --no-cpu-baseline --steps 50 > $O/bench_b8.json 2>> $O/bench.err
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_b8.json").read().strip().splitlines()[-1])
    print("b8 ms/step", round(d["ms_per_step"], 4), "value", round(d["value"], 1))
    print(d["kernel_families_us"])
except Exception as e:
    print("b8 unreadable", e)
PY
tail -n 5 $O/bench.err
