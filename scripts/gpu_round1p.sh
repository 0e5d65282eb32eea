# round 1p: eps-mode compaction of active cases; loop policy by grid shape
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r01p_pytest_new.log 2>&1; tail -n 15 gpurun_out/r01p_pytest_new.log
timeout 400 python bench.py --no-cpu --epsilon 1e-6 --steps 5 > gpurun_out/r01p_bench_fp64_eps.json 2> gpurun_out/r01p_bench_fp64_eps.err; cat gpurun_out/r01p_bench_fp64_eps.json; tail -n 3 gpurun_out/r01p_bench_fp64_eps.err
BNBP_NO_COMPACT=1 timeout 400 python bench.py --no-cpu --epsilon 1e-6 --steps 5 > gpurun_out/r01p_bench_fp64_eps_nocompact.json 2> gpurun_out/r01p_bench_fp64_eps_nocompact.err; cut -c1-260 gpurun_out/r01p_bench_fp64_eps_nocompact.json
timeout 400 python bench.py --no-cpu --epsilon 1e-3 --steps 5 > gpurun_out/r01p_bench_fp64_eps1e3.json 2> gpurun_out/r01p_bench_fp64_eps1e3.err; cut -c1-260 gpurun_out/r01p_bench_fp64_eps1e3.json
timeout 400 python bench.py --no-cpu > gpurun_out/r01p_bench_fp64.json 2> gpurun_out/r01p_bench_fp64.err; cat gpurun_out/r01p_bench_fp64.json
timeout 400 python bench.py --no-cpu --precision fp32 > gpurun_out/r01p_bench_fp32.json 2> gpurun_out/r01p_bench_fp32.err; cut -c1-260 gpurun_out/r01p_bench_fp32.json
ls -la gpurun_out | tail -n 12
