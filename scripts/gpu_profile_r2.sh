#!/bin/bash
# Round-2 ncu evidence (run under gpurun, 1 GPU):  scripts/gpu_profile_r2.sh TAG
#   launches_TAG.csv   every launch of `bench.py --steps 1 --warmup 3 --quick` with its device time (cold-cache,
#                      serialised: compare SHARES); the training step runs as CUDA-graph replays (per-node profiling)
#   traffic_TAG.csv    DRAM / L2 bytes + duration of the stage-A and contraction launches of an eager step
TAG=${1:-r2}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --quick"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_$TAG.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"kp_fwd_fast|kp_bwd_fast|kp_fwd_tiny|gemm_tc" -s 600 -c 200 --csv --log-file gpurun_out/traffic_$TAG.csv $CMD --no-graphs > gpurun_out/ncu_traffic_$TAG.log 2>&1
echo "traffic rc=$?"
ls -la gpurun_out/launches_$TAG.csv gpurun_out/traffic_$TAG.csv
