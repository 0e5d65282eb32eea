# round 1i: fp64 DMMA variant of the dense products vs DFMA, default threshold 256, whole GPU suite
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -x -q > gpurun_out/r01i_pytest_dense.log 2>&1; tail -3 gpurun_out/r01i_pytest_dense.log
for m in 1 0; do
BNBP_DENSE_MMA=$m timeout 600 python bench.py --workload card32 --no-cpu --no-e2e --steps 2 > gpurun_out/r01i_card32_fp64_mma$m.json 2> gpurun_out/r01i_card32_fp64_mma$m.err; tail -2 gpurun_out/r01i_card32_fp64_mma$m.err
BNBP_DENSE_MMA=$m timeout 600 python bench.py --workload dag2000 --no-cpu --no-e2e --steps 2 > gpurun_out/r01i_dag2000_fp64_mma$m.json 2> gpurun_out/r01i_dag2000_fp64_mma$m.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r01i_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        d=j.get("dense") or {}
        print(f, "value %.4g ms/step %.2f hbm-frac %.3f sweep-kernel ms %.3f | dense ms/sweep %s TF %s frac %s share %s nodes %s"%(j["value"], j["ms_per_step"], j["roofline"]["frac"], j["roofline"]["ms_per_launch"], d.get("ms_per_sweep"), d.get("achieved"), d.get("frac"), d.get("share_of_sweep_time"), d.get("nodes")))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01i_pytest_gpu.log 2>&1; tail -3 gpurun_out/r01i_pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_gemm -s 2 -c 1 -o gpurun_out/r01i_dense_fp64_dmma python bench.py --workload card32 --cases 4096 --sweeps 3 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/r01i_ncu_fp64.log 2>&1; tail -2 gpurun_out/r01i_ncu_fp64.log
