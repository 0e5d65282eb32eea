set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01e_pytest.log 2>&1; tail -5 gpurun_out/r01e_pytest.log
timeout 600 python bench.py > gpurun_out/r01e_bench_fp64.json 2> gpurun_out/r01e_bench_fp64.err; cut -c1-300 gpurun_out/r01e_bench_fp64.json; tail -3 gpurun_out/r01e_bench_fp64.err
timeout 300 python bench.py --precision fp32 --no-cpu > gpurun_out/r01e_bench_fp32.json 2> gpurun_out/r01e_bench_fp32.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30 -c 30 --csv --log-file gpurun_out/r01e_launches.csv python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/r01e_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 -o gpurun_out/r01e_generic_dag2000 python bench.py --workload dag2000 --cases 4096 --sweeps 3 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/r01e_ncu_dag.log 2>&1; tail -2 gpurun_out/r01e_ncu_dag.log
timeout 900 python bench.py --workload card32 --cases 1024 --sweeps 2 --no-cpu --no-e2e --steps 1 > gpurun_out/r01e_bench_card32.json 2> gpurun_out/r01e_bench_card32.err; cut -c1-300 gpurun_out/r01e_bench_card32.json; tail -2 gpurun_out/r01e_bench_card32.err
