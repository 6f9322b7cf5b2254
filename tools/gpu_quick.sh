#!/bin/bash
# Quick GPU visit: parity tests + one bench line.  usage: bash tools/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}; O=gpurun_out/$TAG; mkdir -p $O
if [ -n "${2:-}" ]; then K=(-k "$2"); else K=(); fi
timeout 1500 python -m pytest tests -m gpu -x -q "${K[@]}" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 900 python bench.py --cpu-sample 0 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("compress ms", d["ms_per_step"], "match ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"])
print("decompress ms", d["decompress"]["ms_per_step"], "gather ms", d["decompress"]["roofline"]["kernel_ms"], "frac", d["decompress"]["roofline"]["frac"])
print("e2e compress ms", d["e2e"]["ms_per_step"], "e2e decompress ms", d["decompress"]["e2e"]["ms_per_step"], "launches", d["gpu_launches"]); print("resident", d["e2e_resident_reference"]["compress"]["ms_per_step"], d["e2e_resident_reference"]["decompress"]["ms_per_step"])
PY
timeout 600 python tools/time_global.py > $O/global.json 2> $O/global.err; echo "global rc=$?"
python - <<PY
import json
for l in open("$O/global.json"):
    d=json.loads(l); r=d["runs"][-1]
    print(d["config"], "kernels_ms", r["kernels_ms"], "match_ms", r["match_ms"], "launches", r["launches"], "rounds", r["spec_rounds"])
PY
