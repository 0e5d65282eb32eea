set -x
mkdir -p gpurun_out
for i in 1 2 3; do timeout 400 python bench.py --no-cpu > gpurun_out/r01chk3_bench_fp64_$i.json 2> /dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r01chk3_bench_fp64_$i.json').read().strip().splitlines()[-1]); e=d['e2e']; print(round(d['value']/1e6,1), round(d['ms_per_step'],2), 'e2e', round(e['value']/1e6,1), e['ms_per_call_min_median_max'], e['device_ms_per_call_min_median_max'])"; done
timeout 400 python bench.py --no-cpu --epsilon 1e-6 --steps 5 > gpurun_out/r01chk3_bench_fp64_eps.json 2> /dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r01chk3_bench_fp64_eps.json').read().strip().splitlines()[-1]); e=d['e2e']; print(d['value']/1e6, d['ms_per_step'], e['value']/1e6, e['ms_per_call_min_median_max'], e['device_ms_per_call_min_median_max'])"
timeout 400 python bench.py --no-cpu --precision fp32 > gpurun_out/r01chk3_bench_fp32.json 2> /dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r01chk3_bench_fp32.json').read().strip().splitlines()[-1]); e=d['e2e']; print(d['value']/1e6, d['ms_per_step'], e['value']/1e6, e['ms_per_call_min_median_max'], e['device_ms_per_call_min_median_max'])"
