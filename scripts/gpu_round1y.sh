# round 1y: is the slow e2e call a Python garbage collection?  split-path test; default bench
set -x
mkdir -p gpurun_out
timeout 300 python scripts/e2e_probe.py fp64 40 2>&1 | tail -n 1 | cut -c1-400
timeout 300 python scripts/e2e_probe.py fp64 40 nogc 2>&1 | tail -n 1 | cut -c1-400
timeout 300 python scripts/e2e_probe.py fp64 40 2>&1 | tail -n 1 | cut -c1-400
timeout 300 python scripts/e2e_probe.py fp64 40 nogc 2>&1 | tail -n 1 | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_spec.py -m gpu -x -q 2>&1 | tail -n 3
timeout 500 python bench.py --no-cpu > gpurun_out/r01y_bench_fp64.json 2> gpurun_out/r01y_bench_fp64.err; cat gpurun_out/r01y_bench_fp64.json | cut -c1-200; tail -n 3 gpurun_out/r01y_bench_fp64.err
