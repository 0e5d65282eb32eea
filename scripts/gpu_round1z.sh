# round 1z: polling wait vs sleeping wait in the host-buffer call
set -x
mkdir -p gpurun_out
timeout 300 python scripts/e2e_probe.py fp64 40 2>&1 | tail -n 1 | cut -c1-420
BNBP_BLOCKING_SYNC=1 timeout 300 python scripts/e2e_probe.py fp64 40 2>&1 | tail -n 1 | cut -c1-420
timeout 300 python scripts/e2e_probe.py fp64 40 2>&1 | tail -n 1 | cut -c1-420
BNBP_BLOCKING_SYNC=1 timeout 300 python scripts/e2e_probe.py fp64 40 2>&1 | tail -n 1 | cut -c1-420
timeout 500 python bench.py --no-cpu > gpurun_out/r01z_bench_fp64.json 2> gpurun_out/r01z_bench_fp64.err; python -c "
import json; d=json.loads(open('gpurun_out/r01z_bench_fp64.json').read().strip().splitlines()[-1]); e=d['e2e']; print(d['value']/1e6, d['ms_per_step'], e['value']/1e6, e['ms_per_step'], e['ms_per_call_min_median_max'], e['device_ms_per_call_min_median_max'])"
timeout 500 python bench.py --no-cpu --epsilon 1e-6 --steps 5 > gpurun_out/r01z_bench_fp64_eps.json 2> /dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r01z_bench_fp64_eps.json').read().strip().splitlines()[-1]); e=d['e2e']; print(d['value']/1e6, d['ms_per_step'], e['value']/1e6, e['ms_per_step'], e['ms_per_call_min_median_max'])"
