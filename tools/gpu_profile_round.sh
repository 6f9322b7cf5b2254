#!/bin/bash
# Profiles to commit under profiles/: launch list of one bench step and `ncu --set full` of the dominant kernels.
# usage: bash tools/gpu_profile_round.sh <tag>
TAG=${1:-prof}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_bench.csv \
   python bench.py --steps 1 --warmup 1 --cpu-sample 0 > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
for k in seg_match_k dec_gather_k; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$k" -c 1 -o $O/full_$k \
     python bench.py --steps 1 --warmup 1 --cpu-sample 0 > $O/ncu_full_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 900 ncu --set full --clock-control none -k regex:'rle_count_k|rle_write_k|seg_write_k|seg_bytes_k|dec_count_k|dec_emit_k|scan_onepass_k|dec_tile_win_k' -c 12 -o $O/full_small \
     python bench.py --steps 1 --warmup 1 --cpu-sample 0 > $O/ncu_full_small.log 2>&1; echo "ncu small rc=$?"
for shape in gap divergent; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_global_$shape.csv python tools/one_global.py $shape > $O/ncu_global_$shape.log 2>&1; echo "global $shape rc=$?"
done
ls -la $O
