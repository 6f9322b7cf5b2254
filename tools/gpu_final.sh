#!/bin/bash
# Round-end check: full GPU suite, smoke(), default bench (with the CPU baseline leg), launch list of one step.
# usage: bash tools/gpu_final.sh <tag>
TAG=${1:-fin}; O=gpurun_out/$TAG; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -2 $O/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_bench.csv \
   python bench.py --steps 1 --warmup 1 --cpu-sample 0 > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("compress", d["ms_per_step"], d["value"], "frac", d["roofline"]["frac"], "decompress", d["decompress"]["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["decompress"]["e2e"]["ms_per_step"])
print("resident", d["e2e_resident_reference"]["compress"]["ms_per_step"], d["e2e_resident_reference"]["decompress"]["ms_per_step"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], d["clocks"])
PY
