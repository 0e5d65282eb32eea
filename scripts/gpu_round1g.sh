# round 1g: dense contraction path (csrc/bnbp_dense.cuh) -- parity first, then card32 / dag2000 numbers
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -x -q > gpurun_out/r01g_pytest_dense.log 2>&1; tail -15 gpurun_out/r01g_pytest_dense.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_dense.py -x -q -k "card8_k3_default_threshold and fp64 or wide_parents_k7 and fp64" > gpurun_out/r01g_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r01g_memcheck.log
timeout 600 python bench.py --workload card32 --no-cpu --no-e2e --steps 2 > gpurun_out/r01g_card32_fp64.json 2> gpurun_out/r01g_card32_fp64.err; tail -2 gpurun_out/r01g_card32_fp64.err
timeout 600 python bench.py --workload card32 --precision fp32 --no-cpu --no-e2e --steps 2 > gpurun_out/r01g_card32_fp32.json 2> gpurun_out/r01g_card32_fp32.err
timeout 600 python bench.py --workload dag2000 --no-cpu --no-e2e --steps 2 > gpurun_out/r01g_dag2000_fp64.json 2> gpurun_out/r01g_dag2000_fp64.err
timeout 600 python bench.py --workload dag2000 --dense-min 1024 --no-cpu --no-e2e --steps 2 > gpurun_out/r01g_dag2000_fp64_d1024.json 2> gpurun_out/r01g_dag2000_fp64_d1024.err
timeout 600 python bench.py --workload dag2000 --precision fp32 --no-cpu --no-e2e --steps 2 > gpurun_out/r01g_dag2000_fp32.json 2> gpurun_out/r01g_dag2000_fp32.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r01g_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        d=j.get("dense") or {}
        print(f, "value %.4g ms/step %.2f hbm-frac %.3f sweep-kernel ms %.3f | dense ms/sweep %s TF %s frac %s share %s"%(j["value"], j["ms_per_step"], j["roofline"]["frac"], j["roofline"]["ms_per_launch"], d.get("ms_per_sweep"), d.get("achieved"), d.get("frac"), d.get("share_of_sweep_time")))
    except Exception as e:
        print(f, "ERR", e)
PY
