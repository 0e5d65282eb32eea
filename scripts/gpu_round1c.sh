set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r01c_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01c_pytest.log 2>&1; tail -5 gpurun_out/r01c_pytest.log
timeout 300 python bench.py --no-cpu > gpurun_out/r01c_bench_fp64.json 2> gpurun_out/r01c_bench_fp64.err; cat gpurun_out/r01c_bench_fp64.json | cut -c1-600
timeout 600 python scripts/tune_spec.py run > gpurun_out/r01c_tune.txt 2>&1; cat gpurun_out/r01c_tune.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bnbp_spec_sweep -s 10 -c 2 -o gpurun_out/r01c_spec_fp64 python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 --cases 262144 > gpurun_out/r01c_ncu.log 2>&1; tail -3 gpurun_out/r01c_ncu.log
