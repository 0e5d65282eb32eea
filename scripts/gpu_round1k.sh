# round 1k: fresh-container confirmation of the whole GPU suite, the default bench line beside the reference
# arm, the e2e device timeline in fp64, the launch list of the default command and the fp32 traffic capture
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01k_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r01k_pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01k_bench_reference.json 2> gpurun_out/r01k_bench_reference.err; cut -c1-400 gpurun_out/r01k_bench_reference.json
timeout 400 python bench.py > gpurun_out/r01k_bench_fp64.json 2> gpurun_out/r01k_bench_fp64.err; cat gpurun_out/r01k_bench_fp64.json; tail -n 3 gpurun_out/r01k_bench_fp64.err
BNBP_TRACE=1 timeout 300 python bench.py --no-cpu --steps 2 > gpurun_out/r01k_trace_fp64.json 2> gpurun_out/r01k_trace_fp64.err; grep "chunk\|done\|enqueue" gpurun_out/r01k_trace_fp64.err | tail -n 24
timeout 300 python bench.py --no-cpu --precision fp32 > gpurun_out/r01k_bench_fp32.json 2> gpurun_out/r01k_bench_fp32.err; cut -c1-300 gpurun_out/r01k_bench_fp32.json
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01k_launches_alarm37_fp64.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r01k_launches.log 2>&1; tail -n 2 gpurun_out/r01k_launches.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bnbp_spec_sweep -s 10 -c 1 -o gpurun_out/r01k_spec_sweep_fp32 python bench.py --precision fp32 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/r01k_ncu_fp32.log 2>&1; tail -n 2 gpurun_out/r01k_ncu_fp32.log
ls -la gpurun_out
