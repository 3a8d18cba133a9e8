#!/bin/bash
# Round-2 ncu evidence for profiles/ (run under gpurun, one GPU):  bash scripts/profile_r2.sh <tag>
# 1. launch lists of bench.py's timed region (cold-cache, serialised: compare SHARES, not absolutes)
# 2. --set full captures: the layer-chained kernels of the LL step (chain_kernel, chain2_kernel, wgrad_kernel) and the
#    tcgen05 GEMM at the VS / MS shapes
# 3. per-step DRAM traffic with --cache-control none (the caches keep what the running step keeps: params + Adam state in L2)
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --no-cpu --ring 65536 --e2e-steps 2 --configs none"
ncu --metrics gpu__time_duration.sum --clock-control none -s 220 -c 55 --csv --log-file $OUT/${TAG}_launches_LL_chain.csv \
    $B --workload LL --precision 0 --steps 20 --warmup 5 > $OUT/ncu_ll.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 87 --csv --log-file $OUT/${TAG}_launches_VS_tc.csv \
    $B --workload VS --precision 1 --steps 20 --warmup 5 > $OUT/ncu_vs.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -s 220 -c 55 --csv \
    --log-file $OUT/${TAG}_traffic_LL_chain.csv $B --workload LL --precision 0 --steps 20 --warmup 5 > $OUT/ncu_tr_ll.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -s 200 -c 87 --csv \
    --log-file $OUT/${TAG}_traffic_VS_tc.csv $B --workload VS --precision 1 --steps 20 --warmup 5 > $OUT/ncu_tr_vs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 12 -c 3 -f -o $OUT/${TAG}_prof_chain_LL \
    $B --workload LL --precision 0 --steps 10 --warmup 3 > $OUT/ncu_chain_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain2_kernel -s 8 -c 2 -f -o $OUT/${TAG}_prof_chain2_LL \
    $B --workload LL --precision 0 --steps 10 --warmup 3 > $OUT/ncu_chain2_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 8 -c 2 -f -o $OUT/${TAG}_prof_wgrad_LL \
    $B --workload LL --precision 0 --steps 10 --warmup 3 > $OUT/ncu_wgrad_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 60 -c 6 -f -o $OUT/${TAG}_prof_tc_VS \
    $B --workload VS --precision 1 --steps 10 --warmup 3 > $OUT/ncu_vs_full.log 2>&1
for f in ${TAG}_prof_chain_LL ${TAG}_prof_chain2_LL ${TAG}_prof_wgrad_LL ${TAG}_prof_tc_VS; do
  ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/$f.raw.csv 2>/dev/null
done
ls -la $OUT | tail -15
tail -2 $OUT/ncu_ll.log $OUT/ncu_vs.log $OUT/ncu_chain_full.log | cut -c1-300
