set -x
mkdir -p gpurun_out
for mb in 2 3; do BNBP_SPEC_MINB_CHECK=$mb BNBP_TRACE=1 timeout 300 python bench.py --no-cpu --no-e2e --epsilon 1e-6 --steps 3 --warmup 3 > gpurun_out/r01u_eps_mb$mb.json 2> gpurun_out/r01u_eps_mb$mb.err; grep census gpurun_out/r01u_eps_mb$mb.err | tail -n 9 | head -n 3; python -c "
import json; d=json.loads(open('gpurun_out/r01u_eps_mb$mb.json').read().strip().splitlines()[-1]); print('fp64 minb $mb', d['value']/1e6, d['ms_per_step'], d['roofline']['frac'])"; done
for mb in 3 4; do BNBP_SPEC_MINB_CHECK=$mb BNBP_TRACE=1 timeout 300 python bench.py --no-cpu --no-e2e --epsilon 1e-4 --precision fp32 --steps 3 --warmup 3 > gpurun_out/r01u_eps32_mb$mb.json 2> gpurun_out/r01u_eps32_mb$mb.err; grep census gpurun_out/r01u_eps32_mb$mb.err | tail -n 9 | head -n 2; python -c "
import json; d=json.loads(open('gpurun_out/r01u_eps32_mb$mb.json').read().strip().splitlines()[-1]); print('fp32 minb $mb', d['value']/1e6, d['ms_per_step'], d['roofline']['frac'])"; done
