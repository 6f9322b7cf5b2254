#!/bin/bash
# global-mode round: GPU tests of the global path, timing of both global configs (verified against the oracle), launch list.
# usage: bash tools/gpu_global_round.sh <tag> [ncu kernel ...]
TAG=$1; shift; O=gpurun_out/$TAG; mkdir -p $O
python -m pytest tests -x -q -m gpu -k "global or abort or strip" 2>&1 | tail -2
python tools/time_global.py 1.0 --verify > $O/time_global.json 2> $O/time_global.err; python - <<PY
import json
for l in open("$O/time_global.json"):
    d = json.loads(l); print(d["config"], "ms", [round(r["wall_ms"], 2) for r in d["runs"]], "ok", d.get("matches_oracle"))
PY
for S in gap divergent; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$S.csv python tools/one_global.py $S > $O/l_$S.log 2>&1
  python tools/launch_shares.py $O/launches_$S.csv 12
done
for K in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 1 -c 1 -o $O/$K python tools/one_global.py gap > $O/$K.log 2>&1; echo "$K rc=$?"
done
