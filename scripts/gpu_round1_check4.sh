set -x
mkdir -p gpurun_out
for i in 1 2 3 4; do timeout 400 python bench.py --no-cpu --steps 10 > gpurun_out/r01chk5_bench_fp64_$i.json 2> /dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r01chk5_bench_fp64_$i.json').read().strip().splitlines()[-1]); e=d['e2e']; print(round(d['value']/1e6,1), 'e2e', round(e['value']/1e6,1), [round(x,1) for x in e['ms_per_call_min_median_max']], e['slowest_call'])"; done
