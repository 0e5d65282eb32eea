# round 1j: tensor-core (tcgen05, 3xTF32) dense products for fp32 handles: kernel check, parity, card32 bench
set -x
mkdir -p gpurun_out
timeout 120 tests/cuda/_build/test_dense_tc > gpurun_out/r01j_tc_kernel.log 2>&1; echo "rc=$?" >> gpurun_out/r01j_tc_kernel.log; tail -25 gpurun_out/r01j_tc_kernel.log
if grep -q "dense_tc ok" gpurun_out/r01j_tc_kernel.log; then
BNBP_MARGIN_LOG=gpurun_out/r01j_tc_margins.txt timeout 900 python -m pytest tests/test_gpu_dense_tc.py -q > gpurun_out/r01j_pytest_tc.log 2>&1; tail -15 gpurun_out/r01j_pytest_tc.log; cat gpurun_out/r01j_tc_margins.txt
timeout 600 python bench.py --workload card32 --precision fp32 --no-cpu --no-e2e --steps 2 > gpurun_out/r01j_card32_fp32_tc.json 2> gpurun_out/r01j_card32_fp32_tc.err; tail -2 gpurun_out/r01j_card32_fp32_tc.err; cat gpurun_out/r01j_card32_fp32_tc.json
timeout 600 python bench.py --workload dag2000 --precision fp32 --no-cpu --no-e2e --steps 2 > gpurun_out/r01j_dag2000_fp32_tc.json 2> gpurun_out/r01j_dag2000_fp32_tc.err; cat gpurun_out/r01j_dag2000_fp32_tc.json
timeout 600 python bench.py --workload dag2000 --precision fp32 --dense-tensor -1 --no-cpu --no-e2e --steps 2 > gpurun_out/r01j_dag2000_fp32_fma.json 2> gpurun_out/r01j_dag2000_fp32_fma.err; cat gpurun_out/r01j_dag2000_fp32_fma.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_tc -s 2 -c 1 -o gpurun_out/r01j_dense_tc python bench.py --workload card32 --precision fp32 --cases 4096 --sweeps 3 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/r01j_ncu.log 2>&1; tail -2 gpurun_out/r01j_ncu.log
fi
