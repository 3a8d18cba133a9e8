#!/bin/bash
# DRAM traffic of one gradient step under ncu (cold caches, serialised kernels): three consecutive steps of bench.py's timed
# region per workload.  Run under gpurun: bash scripts/traffic_gpu.sh <tag>
set -u
TAG=${1:-r1b}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --no-cpu --ring 65536 --e2e-steps 2"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 150 -c 78 --csv \
    --log-file $OUT/${TAG}_traffic_LL_ffma.csv $B --workload LL --precision 0 --steps 20 --warmup 5 > $OUT/ncu_tr_ll.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 200 -c 87 --csv \
    --log-file $OUT/${TAG}_traffic_VS_tc.csv $B --workload VS --precision 1 --steps 20 --warmup 5 > $OUT/ncu_tr_vs.log 2>&1
ls -la $OUT/${TAG}_traffic_*
