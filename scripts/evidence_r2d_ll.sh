set -u
mkdir -p gpurun_out
TAG=r2d
B="python bench.py --no-cpu --ring 65536 --e2e-steps 2 --configs none"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 220 -c 55 --csv --log-file gpurun_out/${TAG}_launches_LL_chain.csv $B --workload LL --precision 0 --steps 20 --warmup 5 > gpurun_out/ncu_ll.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -s 220 -c 55 --csv --log-file gpurun_out/${TAG}_traffic_LL_chain.csv $B --workload LL --precision 0 --steps 20 --warmup 5 > gpurun_out/ncu_tr_ll.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 12 -c 2 -f -o gpurun_out/${TAG}_prof_chain_LL $B --workload LL --precision 0 --steps 10 --warmup 3 > gpurun_out/ncu_chain_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain2_kernel -s 8 -c 2 -f -o gpurun_out/${TAG}_prof_chain2_LL $B --workload LL --precision 0 --steps 10 --warmup 3 > gpurun_out/ncu_chain2_full.log 2>&1
for f in ${TAG}_prof_chain_LL ${TAG}_prof_chain2_LL; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
timeout 200 python scripts/chain_timeline.py LL > gpurun_out/${TAG}_chain_timeline.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_driver_args.json 2> gpurun_out/${TAG}_bench_driver_args.err
timeout 600 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
python - <<'E'
import json
for f in ('gpurun_out/r2d_bench_driver_args.json','gpurun_out/r2d_bench_default.json'):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['pipelined_update_many']['value'], d['e2e']['with_publication'], {k:(round(v['value']), round(v.get('e2e',{}).get('value',0))) for k,v in d['configs'].items()}, d['by_precision'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['traffic'])
            print(d['roofline']['per_launch_us_in_graph'], d['replicas_per_gpu_sweep'])
E
ls gpurun_out | grep r2d_
