# round 1x: split eps path (plain sweeps + delta_retire_kernel) with compaction; full suite; default bench with CPU legs
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r01x_pytest_new.log 2>&1; tail -n 15 gpurun_out/r01x_pytest_new.log
BNBP_TRACE=1 timeout 400 python bench.py --no-cpu --epsilon 1e-6 --steps 5 > gpurun_out/r01x_bench_fp64_eps.json 2> gpurun_out/r01x_bench_fp64_eps.err; cut -c1-300 gpurun_out/r01x_bench_fp64_eps.json; grep census gpurun_out/r01x_bench_fp64_eps.err | tail -n 10
BNBP_NO_SPLIT=1 timeout 400 python bench.py --no-cpu --no-e2e --epsilon 1e-6 --steps 5 > gpurun_out/r01x_bench_fp64_eps_nosplit.json 2> /dev/null; cut -c1-260 gpurun_out/r01x_bench_fp64_eps_nosplit.json
timeout 400 python bench.py --no-cpu --epsilon 1e-3 --steps 5 > gpurun_out/r01x_bench_fp64_eps1e3.json 2> /dev/null; cut -c1-260 gpurun_out/r01x_bench_fp64_eps1e3.json
timeout 400 python bench.py --no-cpu --no-e2e --epsilon 1e-4 --precision fp32 --steps 5 > gpurun_out/r01x_bench_fp32_eps.json 2> /dev/null; cut -c1-260 gpurun_out/r01x_bench_fp32_eps.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01x_pytest_gpu.log 2>&1; tail -n 5 gpurun_out/r01x_pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01x_bench_reference.json 2> gpurun_out/r01x_bench_reference.err; cut -c1-300 gpurun_out/r01x_bench_reference.json
timeout 500 python bench.py > gpurun_out/r01x_bench_fp64.json 2> gpurun_out/r01x_bench_fp64.err; cat gpurun_out/r01x_bench_fp64.json; tail -n 3 gpurun_out/r01x_bench_fp64.err
