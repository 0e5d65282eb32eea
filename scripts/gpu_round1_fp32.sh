set -x
mkdir -p gpurun_out
for mb in 4 5 6; do BNBP_SPEC_MINB=$mb timeout 300 python bench.py --no-cpu --no-e2e --precision fp32 > gpurun_out/r01fp32_mb$mb.json 2> /dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r01fp32_mb$mb.json').read().strip().splitlines()[-1]); print('minb $mb', round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['roofline']['frac'],3), round(d['roofline']['ms_per_launch'],4))"; done
