# round 1m: K0 fused into the first sweep, K4 into the last (spec variants 5, 6/7); chunk plan by copy/compute ratio
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r01m_pytest_gpu.log 2>&1; tail -n 15 gpurun_out/r01m_pytest_gpu.log
timeout 400 python bench.py --no-cpu > gpurun_out/r01m_bench_fp64.json 2> gpurun_out/r01m_bench_fp64.err; cat gpurun_out/r01m_bench_fp64.json; tail -n 3 gpurun_out/r01m_bench_fp64.err
BNBP_NO_FUSE=1 timeout 400 python bench.py --no-cpu --no-e2e > gpurun_out/r01m_bench_fp64_nofuse.json 2> gpurun_out/r01m_bench_fp64_nofuse.err; cut -c1-200 gpurun_out/r01m_bench_fp64_nofuse.json
BNBP_TRACE=1 timeout 300 python bench.py --no-cpu --steps 2 > gpurun_out/r01m_trace_fp64.json 2> gpurun_out/r01m_trace_fp64.err; grep "wave\|chunk [0-9]*:\|done" gpurun_out/r01m_trace_fp64.err | tail -n 12
timeout 300 python bench.py --no-cpu --precision fp32 > gpurun_out/r01m_bench_fp32.json 2> gpurun_out/r01m_bench_fp32.err; cut -c1-1500 gpurun_out/r01m_bench_fp32.json
BNBP_TRACE=1 timeout 300 python bench.py --no-cpu --steps 2 --precision fp32 > gpurun_out/r01m_trace_fp32.json 2> gpurun_out/r01m_trace_fp32.err; grep "wave\|chunk [0-9]*:\|done" gpurun_out/r01m_trace_fp32.err | tail -n 12
for k in 3 5 6; do BNBP_CHUNKS=$k timeout 300 python bench.py --no-cpu --steps 5 > gpurun_out/r01m_bench_fp64_k$k.json 2>/dev/null; python -c "
import json,sys; d=json.loads(open('gpurun_out/r01m_bench_fp64_k$k.json').read().strip().splitlines()[-1]); print('chunks $k', d['ms_per_step'], d['e2e']['ms_per_step'])"; done
ls -la gpurun_out
