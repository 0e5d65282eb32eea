# round 1 final 2: whole GPU suite at HEAD, the other BASELINE configs (cfg 3/4/5), eps mode on the generic family
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01fin2_pytest_gpu.log 2>&1; tail -n 4 gpurun_out/r01fin2_pytest_gpu.log
timeout 400 python bench.py --workload grid100 --cases 16384 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r01fin2_grid100_fp64.json 2> gpurun_out/r01fin2_grid100_fp64.err; cut -c1-200 gpurun_out/r01fin2_grid100_fp64.json; tail -n 2 gpurun_out/r01fin2_grid100_fp64.err
timeout 400 python bench.py --workload grid100 --cases 16384 --steps 2 --warmup 3 --no-cpu --no-e2e --precision fp32 > gpurun_out/r01fin2_grid100_fp32.json 2> /dev/null; cut -c1-200 gpurun_out/r01fin2_grid100_fp32.json
timeout 400 python bench.py --workload dag2000 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r01fin2_dag2000_fp64.json 2> /dev/null; cut -c1-200 gpurun_out/r01fin2_dag2000_fp64.json
timeout 400 python bench.py --workload dag2000 --steps 2 --warmup 3 --no-cpu --no-e2e --precision fp32 > gpurun_out/r01fin2_dag2000_fp32.json 2> /dev/null; cut -c1-200 gpurun_out/r01fin2_dag2000_fp32.json
timeout 400 python bench.py --workload card32 --steps 2 --warmup 3 --no-cpu --no-e2e --precision fp32 > gpurun_out/r01fin2_card32_fp32.json 2> /dev/null; cut -c1-200 gpurun_out/r01fin2_card32_fp32.json
timeout 400 python bench.py --specialize never --no-cpu --no-e2e --epsilon 1e-6 --steps 3 > gpurun_out/r01fin2_alarm37_generic_eps.json 2> /dev/null; cut -c1-200 gpurun_out/r01fin2_alarm37_generic_eps.json
BNBP_NO_SPLIT=1 BNBP_NO_COMPACT=1 timeout 400 python bench.py --specialize never --no-cpu --no-e2e --epsilon 1e-6 --steps 3 > gpurun_out/r01fin2_alarm37_generic_eps_old_path.json 2> /dev/null; cut -c1-200 gpurun_out/r01fin2_alarm37_generic_eps_old_path.json
timeout 400 python bench.py --workload grid100 --cases 16384 --steps 1 --warmup 3 --no-cpu --no-e2e --epsilon 1e-6 > gpurun_out/r01fin2_grid100_eps.json 2> /dev/null; cut -c1-200 gpurun_out/r01fin2_grid100_eps.json
timeout 400 python bench.py > gpurun_out/r01fin2_bench_fp64.json 2> /dev/null; cut -c1-200 gpurun_out/r01fin2_bench_fp64.json
