# round 1n: looped middle sweeps + fused ends + explicit belief arithmetic; LW / CPT-estimation kernels; e2e probe
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec.py tests/test_lw.py tests/test_dropin_cpp.py -m gpu -x -q > gpurun_out/r01n_pytest_new.log 2>&1; tail -n 15 gpurun_out/r01n_pytest_new.log
timeout 400 python bench.py --no-cpu > gpurun_out/r01n_bench_fp64.json 2> gpurun_out/r01n_bench_fp64.err; cat gpurun_out/r01n_bench_fp64.json; tail -n 3 gpurun_out/r01n_bench_fp64.err
BNBP_NO_LOOP=1 timeout 400 python bench.py --no-cpu > gpurun_out/r01n_bench_fp64_noloop.json 2> gpurun_out/r01n_bench_fp64_noloop.err; cut -c1-200 gpurun_out/r01n_bench_fp64_noloop.json
BNBP_TRACE=1 timeout 300 python bench.py --no-cpu --steps 2 > gpurun_out/r01n_trace_fp64.json 2> gpurun_out/r01n_trace_fp64.err; grep "wave\|chunk [0-9]*:\|done" gpurun_out/r01n_trace_fp64.err | tail -n 6
timeout 300 python bench.py --no-cpu --precision fp32 > gpurun_out/r01n_bench_fp32.json 2> gpurun_out/r01n_bench_fp32.err; cut -c1-200 gpurun_out/r01n_bench_fp32.json
BNBP_TRACE=1 timeout 300 python bench.py --no-cpu --steps 2 --precision fp32 > gpurun_out/r01n_trace_fp32.json 2> gpurun_out/r01n_trace_fp32.err; grep "wave\|chunk [0-9]*:\|done" gpurun_out/r01n_trace_fp32.err | tail -n 7
timeout 200 python scripts/e2e_probe.py fp64 16 2>&1 | tail -n 1
timeout 200 python scripts/e2e_probe.py fp32 16 2>&1 | tail -n 1
for k in 5 6 8; do BNBP_CHUNKS=$k timeout 200 python scripts/e2e_probe.py fp64 8 2>&1 | tail -n 1; done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01n_pytest_gpu.log 2>&1; tail -n 5 gpurun_out/r01n_pytest_gpu.log
ls gpurun_out | head -50
