set -u
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 1200 bash scripts/profile_r2.sh $TAG > gpurun_out/${TAG}_profile.log 2>&1
timeout 200 python scripts/chain_timeline.py LL > gpurun_out/${TAG}_chain_timeline.log 2>&1
rm -f gpurun_out/${TAG}_tc_timeline.log
for v in "64 1" "128 1" "160 1"; do set -- $v; echo "== BN $1" >> gpurun_out/${TAG}_tc_timeline.log; B200SAC_TC_BN=$1 timeout 100 python scripts/tc_timeline.py 0,4096,400,400 1,2048,400,400 2,400,400,1024 >> gpurun_out/${TAG}_tc_timeline.log 2>&1; done
for w in VS MS C10 C10O; do timeout 300 python bench.py --workload $w --precision 1 --configs none --steps 200 --warmup 20 --cpu-seconds 4 > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; done
timeout 600 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/${TAG}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/${TAG}_sanitizer_racecheck.log
python - <<E
import json
for w in ('VS','MS','C10','C10O'):
    for line in open('gpurun_out/${TAG}_bench_%s.json' % w):
        if line.startswith('{'):
            d=json.loads(line); print(w, round(d['value']), round(d['e2e']['value']), d['cpu_baseline']['value'] if d.get('cpu_baseline') else None)
for line in open('gpurun_out/${TAG}_bench_default.json'):
    if line.startswith('{'):
        d=json.loads(line); print('LL', d['value'], d['e2e']['value'], d['e2e']['with_publication']['overlapped_run_loop'], {k:(round(v['value']), round(v.get('e2e',{}).get('value',0))) for k,v in d['configs'].items()}, d['bench_wall_s'])
E
