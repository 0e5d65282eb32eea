# end of round 1: chunk-count sweep of the host-buffer call, then both bench arms as the driver runs them
set -x
mkdir -p gpurun_out
for k in 4 5 6 8; do BNBP_CHUNKS=$k timeout 200 python scripts/e2e_probe.py fp64 14 gc reuse 2>&1 | tail -n 1 | cut -c1-260; done
for k in 3 4 5; do BNBP_CHUNKS=$k timeout 200 python scripts/e2e_probe.py fp32 14 gc reuse 2>&1 | tail -n 1 | cut -c1-260; done
timeout 300 python bench.py --impl reference > gpurun_out/r01end_bench_reference.json 2> /dev/null; cut -c1-200 gpurun_out/r01end_bench_reference.json
timeout 500 python bench.py > gpurun_out/r01end_bench_fp64.json 2> gpurun_out/r01end_bench_fp64.err; cat gpurun_out/r01end_bench_fp64.json; tail -n 2 gpurun_out/r01end_bench_fp64.err
timeout 400 python bench.py --no-cpu --precision fp32 > gpurun_out/r01end_bench_fp32.json 2> /dev/null; cut -c1-200 gpurun_out/r01end_bench_fp32.json
timeout 400 python bench.py --no-cpu --epsilon 1e-6 --steps 5 > gpurun_out/r01end_bench_fp64_eps1e-6.json 2> /dev/null; cut -c1-200 gpurun_out/r01end_bench_fp64_eps1e-6.json
