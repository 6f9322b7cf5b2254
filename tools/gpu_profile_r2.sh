#!/bin/bash
# Round-2 profile set for profiles/: bench line, launch lists (chr1-sized local pair, both global pairs), ncu --set full of the two dominant kernels.
# usage: bash tools/gpu_profile_r2.sh <tag>
TAG=${1:-prof}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python tools/one_chr1.py 3 > $O/one_chr1.log 2>&1; echo "plain rc=$?"; cat $O/one_chr1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_chr1.csv python tools/one_chr1.py 2 > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
for k in seg_match_k dec_gather_k; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" --launch-skip 1 -c 1 -o $O/full_$k python tools/one_chr1.py 2 > $O/ncu_full_$k.log 2>&1; echo "ncu $k rc=$?"
done
for shape in gap divergent; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_global_$shape.csv python tools/one_global.py $shape > $O/ncu_global_$shape.log 2>&1; echo "global $shape rc=$?"
done
ls -la $O
