set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r01f_pytest.log 2>&1; tail -3 gpurun_out/r01f_pytest.log
BNBP_TRACE=1 timeout 300 python bench.py --precision fp32 --no-cpu --steps 2 > gpurun_out/r01f_trace_fp32.json 2> gpurun_out/r01f_trace_fp32.err; tail -40 gpurun_out/r01f_trace_fp32.err
timeout 300 python bench.py --no-cpu > gpurun_out/r01f_bench_fp64.json 2> gpurun_out/r01f_bench_fp64.err
timeout 300 python bench.py --no-cpu --precision fp32 > gpurun_out/r01f_bench_fp32.json 2> gpurun_out/r01f_bench_fp32.err
timeout 300 python bench.py --no-cpu --epsilon 1e-6 > gpurun_out/r01f_bench_fp64_eps.json 2> gpurun_out/r01f_bench_fp64_eps.err
timeout 300 python bench.py --workload dag2000 --cases 16384 --no-cpu --no-e2e --steps 2 > gpurun_out/r01f_dag_vec1.json 2>&1
BNBP_VEC=2 timeout 300 python bench.py --workload dag2000 --cases 16384 --no-cpu --no-e2e --steps 2 > gpurun_out/r01f_dag_vec2.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r01f_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms/step %.2f frac %.3f e2e %s"%(j["value"], j["ms_per_step"], j["roofline"]["frac"], (j.get("e2e") or {}).get("value")), j["config"].get("eps_mode"))
    except Exception as e:
        print(f, "ERR", e)
PY
