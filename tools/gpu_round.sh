#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list, ncu --set full of the top kernels.
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [skip-tests]
set -u
TAG=${1:-rX}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
if [ "${2:-}" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
  tail -5 $O/pytest_gpu.log
fi
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
cat $O/bench.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
cat $O/bench_ref.json
timeout 600 python tools/time_global.py > $O/global.json 2> $O/global.err; echo "global rc=$?"
cat $O/global.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv \
   python bench.py --steps 1 --warmup 1 --cpu-sample 0 > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'seg_match_k|dec_gather_k|seg_write_k|rle_count_k|rle_write_k' -c 10 \
   -o $O/full_top python bench.py --steps 1 --warmup 1 --cpu-sample 0 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O
