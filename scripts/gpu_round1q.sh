set -x
mkdir -p gpurun_out
BNBP_TRACE=1 timeout 300 python bench.py --no-cpu --no-e2e --epsilon 1e-6 --steps 2 --warmup 3 > gpurun_out/r01q_eps_trace.json 2> gpurun_out/r01q_eps_trace.err; grep census gpurun_out/r01q_eps_trace.err | tail -n 12
BNBP_TRACE=1 BNBP_NO_COMPACT=1 timeout 300 python bench.py --no-cpu --no-e2e --epsilon 1e-6 --steps 2 --warmup 3 > gpurun_out/r01q_eps_trace_nc.json 2> gpurun_out/r01q_eps_trace_nc.err; cut -c1-200 gpurun_out/r01q_eps_trace_nc.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01q_eps_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --epsilon 1e-6 > gpurun_out/r01q_eps_launches.log 2>&1; tail -n 2 gpurun_out/r01q_eps_launches.log
