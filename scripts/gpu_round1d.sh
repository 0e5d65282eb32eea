set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01d_pytest.log 2>&1; tail -5 gpurun_out/r01d_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01d_smoke.log 2>&1; tail -2 gpurun_out/r01d_smoke.log
timeout 600 python bench.py > gpurun_out/r01d_bench_fp64.json 2> gpurun_out/r01d_bench_fp64.err; cut -c1-400 gpurun_out/r01d_bench_fp64.json; tail -3 gpurun_out/r01d_bench_fp64.err
timeout 300 python bench.py --precision fp32 --no-cpu > gpurun_out/r01d_bench_fp32.json 2> gpurun_out/r01d_bench_fp32.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01d_bench_reference.json 2> gpurun_out/r01d_bench_reference.err; cat gpurun_out/r01d_bench_reference.json | cut -c1-300
# launch list of one full step (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30 -c 30 --csv --log-file gpurun_out/r01d_launches.csv python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/r01d_launches.log 2>&1
# full-size capture of the specialised kernel for the traffic figure
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bnbp_spec_sweep -s 10 -c 1 -o gpurun_out/r01d_spec_fp64_1m python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/r01d_ncu.log 2>&1; tail -2 gpurun_out/r01d_ncu.log
for w in grid100 dag2000; do timeout 600 python bench.py --workload $w --no-cpu --no-e2e --cases 16384 --steps 2 > gpurun_out/r01d_bench_$w.json 2> gpurun_out/r01d_bench_$w.err; cut -c1-300 gpurun_out/r01d_bench_$w.json; tail -2 gpurun_out/r01d_bench_$w.err; done
