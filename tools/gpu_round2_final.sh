#!/bin/bash
# Round-2 closing visit: full GPU suite, smoke(), default bench line, launch lists (chr1-sized local pair, both global pairs).
# usage: bash tools/gpu_round2_final.sh <tag>
TAG=${1:-fin}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -2 $O/bench.err
for shape in gap divergent; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_global_$shape.csv python tools/one_global.py $shape > $O/ncu_global_$shape.log 2>&1; echo "global $shape rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_chr1.csv python tools/one_chr1.py 2 > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
for k in seg_match_defer_k seg_match_queue_k; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$k" --launch-skip 1 -c 1 -o $O/full_$k python tools/one_chr1.py 2 > $O/ncu_full_$k.log 2>&1; echo "ncu $k rc=$?"
done
python - <<PY
import json
d=json.loads([l for l in open("$O/bench.json") if l.startswith("{")][-1])      # (NCCL may print its version banner first)
print("compress", d["ms_per_step"], d["value"], "frac", d["roofline"]["frac"], "decompress", d["decompress"]["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["decompress"]["e2e"]["ms_per_step"])
for k in ("global_gap_chr19", "global_divergent_chr21"):
    g=d[k]; print(k, g["ms_per_step"], g["index_ms"], g["parse_ms"], g["e2e"]["ms_per_step"], g.get("verified_against_oracle"), g.get("parity_vs_reference"), g.get("index_stride"))
print("verified", d.get("verified_against_oracle"), d.get("parity_vs_reference"), "launches", d["gpu_launches"], d["clocks"])
PY
