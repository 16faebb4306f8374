#!/bin/bash
# ncu evidence for one round (run under gpurun, 1 GPU).
#   $1 = tag (e.g. r1c)   $2 = what: "list" (launch list of the bench command), "full" (full-set captures), "all"
# gpurun copies back at most 64 MiB: the .ncu-rep files are exported to text on the box and only
# kept when small enough.
TAG=${1:-r1}
WHAT=${2:-all}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --quick"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_$TAG.log
if [ "$WHAT" = "list" ] || [ "$WHAT" = "all" ]; then
  # every launch with its device time (cold-cache, serialised: compare SHARES)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
  echo "launch list rc=$?"
fi
if [ "$WHAT" = "full" ] || [ "$WHAT" = "all" ]; then
  PAT=${PAT:-'regex:gemm_tc|col_stats|act_bwd|scale_shift|k_query|split_bf16|pool_|kp_fwd|kp_bwd'}
  for part in fwd bwd; do
    if [ $part = fwd ]; then SKIP=${SKIP_FWD:-1560}; else SKIP=${SKIP_BWD:-1900}; fi
    REP=gpurun_out/prof_${TAG}_$part
    timeout 900 ncu --set full --clock-control none --import-source on -k "$PAT" -s $SKIP -c ${CNT:-24} -o $REP -f $CMD > gpurun_out/ncu_full_${TAG}_$part.log 2>&1
    echo "full capture ($part) rc=$?"
    ncu -i $REP.ncu-rep --page details > gpurun_out/details_${TAG}_$part.txt 2>/dev/null
    ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}_$part.csv 2>/dev/null
    SZ=$(stat -c %s $REP.ncu-rep)
    if [ "$SZ" -gt 28000000 ]; then rm -f $REP.ncu-rep; echo "dropped $REP.ncu-rep ($SZ bytes)"; fi
  done
fi
du -sh gpurun_out; ls -la gpurun_out | tail -12
if [ "$WHAT" = "traffic" ]; then
  # DRAM bytes of every contraction launch of one step (feeds profiles/traffic.json)
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tc -s ${SKIP_GEMM:-470} -c 152 --csv --log-file gpurun_out/gemm_traffic_$TAG.csv $CMD > gpurun_out/ncu_traffic_$TAG.log 2>&1
  echo "traffic rc=$?"
fi
