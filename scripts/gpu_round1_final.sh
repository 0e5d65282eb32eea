# round 1 final: whole GPU suite, both bench arms as the driver runs them, eps-mode lines, launch list, host-stall probe
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01fin_pytest_gpu.log 2>&1; tail -n 4 gpurun_out/r01fin_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 200 python scripts/host_stall_probe.py 2>&1 | tee gpurun_out/r01fin_host_stall_probe.txt
timeout 300 python bench.py --impl reference > gpurun_out/r01fin_bench_reference.json 2> gpurun_out/r01fin_bench_reference.err; cut -c1-300 gpurun_out/r01fin_bench_reference.json
timeout 500 python bench.py > gpurun_out/r01fin_bench_fp64.json 2> gpurun_out/r01fin_bench_fp64.err; cat gpurun_out/r01fin_bench_fp64.json; tail -n 3 gpurun_out/r01fin_bench_fp64.err
timeout 400 python bench.py --no-cpu --precision fp32 > gpurun_out/r01fin_bench_fp32.json 2> gpurun_out/r01fin_bench_fp32.err; cut -c1-200 gpurun_out/r01fin_bench_fp32.json
timeout 400 python bench.py --no-cpu --epsilon 1e-6 --steps 5 > gpurun_out/r01fin_bench_fp64_eps1e-6.json 2> /dev/null; cut -c1-200 gpurun_out/r01fin_bench_fp64_eps1e-6.json
BNBP_NO_SPLIT=1 BNBP_NO_COMPACT=1 timeout 400 python bench.py --no-cpu --no-e2e --epsilon 1e-6 --steps 5 > gpurun_out/r01fin_bench_fp64_eps1e-6_old_path.json 2> /dev/null; cut -c1-200 gpurun_out/r01fin_bench_fp64_eps1e-6_old_path.json
timeout 400 python bench.py --no-cpu --epsilon 1e-3 --steps 5 > gpurun_out/r01fin_bench_fp64_eps1e-3.json 2> /dev/null; cut -c1-200 gpurun_out/r01fin_bench_fp64_eps1e-3.json
BNBP_TRACE=1 timeout 300 python bench.py --no-cpu --steps 2 > /dev/null 2> gpurun_out/r01fin_trace_fp64.err; grep "wave\|chunk [0-9]*:" gpurun_out/r01fin_trace_fp64.err | tail -n 5 > gpurun_out/r01fin_e2e_device_timeline_fp64.txt; cat gpurun_out/r01fin_e2e_device_timeline_fp64.txt
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01fin_launches_alarm37_fp64.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r01fin_launches.log 2>&1; tail -n 1 gpurun_out/r01fin_launches.log | cut -c1-200
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01fin_launches_alarm37_fp64_eps.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --epsilon 1e-6 > gpurun_out/r01fin_launches_eps.log 2>&1; tail -n 1 gpurun_out/r01fin_launches_eps.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:delta_retire -s 1 -c 1 -o gpurun_out/r01fin_delta_retire_fp64 python bench.py --no-cpu --no-e2e --epsilon 1e-6 --steps 1 --warmup 3 > gpurun_out/r01fin_ncu_delta.log 2>&1; tail -n 1 gpurun_out/r01fin_ncu_delta.log | cut -c1-200
ls gpurun_out | grep r01fin
