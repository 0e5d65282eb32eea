set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r01last_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r01last_pytest_gpu.log
timeout 500 python bench.py --no-cpu > gpurun_out/r01last_bench_fp64.json 2> /dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r01last_bench_fp64.json').read().strip().splitlines()[-1]); e=d['e2e']; print(round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['roofline']['frac'],3), 'e2e', round(e['value']/1e6,1), [round(x,2) for x in e['ms_per_call_min_median_max']], [round(x,2) for x in e['device_ms_per_call_min_median_max']])"
timeout 500 python bench.py --no-cpu --precision fp32 > gpurun_out/r01last_bench_fp32.json 2> /dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r01last_bench_fp32.json').read().strip().splitlines()[-1]); e=d['e2e']; print(round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['roofline']['frac'],3), 'e2e', round(e['value']/1e6,1), [round(x,2) for x in e['ms_per_call_min_median_max']], [round(x,2) for x in e['device_ms_per_call_min_median_max']])"
BNBP_TRACE=1 timeout 300 python bench.py --no-cpu --steps 2 > /dev/null 2> gpurun_out/r01last_trace_fp64.err; grep "wave\|chunk [0-9]*:" gpurun_out/r01last_trace_fp64.err | tail -n 6 | tee gpurun_out/r01last_e2e_device_timeline_fp64.txt
