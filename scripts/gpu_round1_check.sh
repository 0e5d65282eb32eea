set -x
mkdir -p gpurun_out
timeout 300 python scripts/e2e_probe.py fp64 40 gc reuse 2>&1 | tail -n 1 | cut -c1-420
timeout 300 python scripts/e2e_probe.py fp64 40 gc alloc 2>&1 | tail -n 1 | cut -c1-420
timeout 300 python scripts/e2e_probe.py fp64 40 gc reuse 2>&1 | tail -n 1 | cut -c1-420
timeout 300 python scripts/e2e_probe.py fp64 40 gc alloc 2>&1 | tail -n 1 | cut -c1-420
timeout 600 python -m pytest tests/test_gpu_spec.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -n 2
timeout 400 python bench.py --no-cpu --epsilon 1e-6 --steps 5 > gpurun_out/r01chk_bench_fp64_eps.json 2> /dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r01chk_bench_fp64_eps.json').read().strip().splitlines()[-1]); e=d['e2e']; print(d['value']/1e6, d['ms_per_step'], e['value']/1e6, e['ms_per_call_min_median_max'], e['device_ms_per_call_min_median_max'])"
