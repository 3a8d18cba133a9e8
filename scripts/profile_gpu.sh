#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, one GPU):  bash scripts/profile_gpu.sh <tag>
# 1. launch lists of bench.py's timed region (three steps each; cold-cache, serialised: compare SHARES)
# 2. one --set full capture of the dominant GEMM kernel per back-end (+ the thin backward kernel)
set -u
TAG=${1:-r1b}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --no-cpu --ring 65536 --e2e-steps 2"
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 78 --csv --log-file $OUT/${TAG}_launches_LL_ffma.csv \
    $B --workload LL --precision 0 --steps 20 --warmup 5 > $OUT/ncu_ll.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 87 --csv --log-file $OUT/${TAG}_launches_VS_tc.csv \
    $B --workload VS --precision 1 --steps 20 --warmup 5 > $OUT/ncu_vs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_simt -s 40 -c 4 -f -o $OUT/${TAG}_prof_ffma_LL \
    $B --workload LL --precision 0 --steps 10 --warmup 3 > $OUT/ncu_ll_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 60 -c 4 -f -o $OUT/${TAG}_prof_tc_VS \
    $B --workload VS --precision 1 --steps 10 --warmup 3 > $OUT/ncu_vs_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_thin -s 9 -c 3 -f -o $OUT/${TAG}_prof_thin_LL \
    $B --workload LL --precision 0 --steps 10 --warmup 3 > $OUT/ncu_thin_full.log 2>&1
for f in ${TAG}_prof_ffma_LL ${TAG}_prof_tc_VS ${TAG}_prof_thin_LL; do
  ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/$f.raw.csv 2>/dev/null
done
ls -la $OUT | tail -15
tail -2 $OUT/ncu_ll.log $OUT/ncu_vs.log | cut -c1-300
