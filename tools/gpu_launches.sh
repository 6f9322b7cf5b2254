#!/bin/bash
# ncu launch list of one bench step.  usage: bash tools/gpu_launches.sh <tag>
TAG=${1:-l}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_bench.csv \
   python bench.py --steps 1 --warmup 1 --cpu-sample 0 > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_shares.py $O/launches_bench.csv 40
