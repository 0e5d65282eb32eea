# round 1o: looped middle sweeps as a real call with restrict parameters; LW / CPT estimation; full suite; profiles
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec.py tests/test_lw.py tests/test_dropin_cpp.py -m gpu -x -q > gpurun_out/r01o_pytest_new.log 2>&1; tail -n 15 gpurun_out/r01o_pytest_new.log
timeout 400 python bench.py --no-cpu > gpurun_out/r01o_bench_fp64.json 2> gpurun_out/r01o_bench_fp64.err; cat gpurun_out/r01o_bench_fp64.json; tail -n 3 gpurun_out/r01o_bench_fp64.err
BNBP_NO_LOOP=1 timeout 400 python bench.py --no-cpu > gpurun_out/r01o_bench_fp64_noloop.json 2> gpurun_out/r01o_bench_fp64_noloop.err; cut -c1-200 gpurun_out/r01o_bench_fp64_noloop.json
BNBP_TRACE=1 timeout 300 python bench.py --no-cpu --steps 2 > gpurun_out/r01o_trace_fp64.json 2> gpurun_out/r01o_trace_fp64.err; grep "wave\|chunk [0-9]*:\|done" gpurun_out/r01o_trace_fp64.err | tail -n 6
timeout 300 python bench.py --no-cpu --precision fp32 > gpurun_out/r01o_bench_fp32.json 2> gpurun_out/r01o_bench_fp32.err; cut -c1-200 gpurun_out/r01o_bench_fp32.json
BNBP_TRACE=1 timeout 300 python bench.py --no-cpu --steps 2 --precision fp32 > gpurun_out/r01o_trace_fp32.json 2> gpurun_out/r01o_trace_fp32.err; grep "wave\|chunk [0-9]*:\|done" gpurun_out/r01o_trace_fp32.err | tail -n 7
for k in 4 6 8; do BNBP_CHUNKS=$k timeout 200 python scripts/e2e_probe.py fp64 10 2>&1 | tail -n 1; done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01o_pytest_gpu.log 2>&1; tail -n 5 gpurun_out/r01o_pytest_gpu.log
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01o_launches_alarm37_fp64.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r01o_launches.log 2>&1; tail -n 2 gpurun_out/r01o_launches.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bnbp_spec_sweep -s 6 -c 3 -o gpurun_out/r01o_spec_sweep_fp64 python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/r01o_ncu_fp64.log 2>&1; tail -n 2 gpurun_out/r01o_ncu_fp64.log
ls -la gpurun_out | tail -n 20
