# round 1l: three-stream host-buffer pipeline (whole-batch staging, wave-planned chunks): parity suite,
# e2e bench in both precisions, device timeline, ring-staging fallback
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01l_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r01l_pytest_gpu.log
timeout 400 python bench.py --no-cpu > gpurun_out/r01l_bench_fp64.json 2> gpurun_out/r01l_bench_fp64.err; cat gpurun_out/r01l_bench_fp64.json; tail -n 3 gpurun_out/r01l_bench_fp64.err
BNBP_TRACE=1 timeout 300 python bench.py --no-cpu --steps 2 > gpurun_out/r01l_trace_fp64.json 2> gpurun_out/r01l_trace_fp64.err; grep "wave\|chunk [0-9]*:\|done" gpurun_out/r01l_trace_fp64.err | tail -n 24
timeout 300 python bench.py --no-cpu --precision fp32 > gpurun_out/r01l_bench_fp32.json 2> gpurun_out/r01l_bench_fp32.err; cut -c1-1200 gpurun_out/r01l_bench_fp32.json
BNBP_TRACE=1 timeout 300 python bench.py --no-cpu --steps 2 --precision fp32 > gpurun_out/r01l_trace_fp32.json 2> gpurun_out/r01l_trace_fp32.err; grep "wave\|chunk [0-9]*:\|done" gpurun_out/r01l_trace_fp32.err | tail -n 24
BNBP_STAGING_RING=1 timeout 300 python bench.py --no-cpu --steps 5 > gpurun_out/r01l_bench_fp64_ring.json 2> gpurun_out/r01l_bench_fp64_ring.err; cut -c1-300 gpurun_out/r01l_bench_fp64_ring.json
ls -la gpurun_out
