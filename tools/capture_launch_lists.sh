#!/bin/bash
# ncu launch lists (device time + DRAM bytes of every launch) of bench.py's headline step for the workloads whose
# roofline.traffic bench.py reads from profiles/ (TRAFFIC_PROFILES).  Run under gpurun; copies go to gpurun_out/.
set -u
mkdir -p gpurun_out
for W in "$@"; do
  python bench.py --only-headline --workload $W --steps 2 --warmup 1 > gpurun_out/plain_$W.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/launches_$W.csv python bench.py --only-headline --workload $W --steps 2 --warmup 1 > gpurun_out/ncu_$W.log 2>&1
  echo "$W: rc=$? $(grep -c k_fwd_cluster gpurun_out/launches_$W.csv) cluster-kernel rows"
done
