set -x
mkdir -p gpurun_out
BNBP_SPEC_MINB_CHECK=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:bnbp_spec_sweep -s 2 -c 1 -o gpurun_out/r01s_spec_check_fp64 python bench.py --no-cpu --no-e2e --epsilon 1e-6 --steps 1 --warmup 3 > gpurun_out/r01s_ncu.log 2>&1; tail -n 2 gpurun_out/r01s_ncu.log
