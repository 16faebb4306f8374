#!/bin/bash
# ncu evidence for one round (run under gpurun, 1 GPU).  $1 = tag (e.g. r1a)
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --quick"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
cat gpurun_out/plain_$TAG.log | tail -1
# every launch with its device time (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
# the stage-A kernels and the contraction, full set, a few launches each
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kp_weighted|gemm_tc" -s 60 -c 12 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
