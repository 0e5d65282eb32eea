# round 1h: dense kernel v2 (conflict-free operand mapping, 128/64/32-column tile classes): parity, thresholds, ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -x -q > gpurun_out/r01h_pytest_dense.log 2>&1; tail -3 gpurun_out/r01h_pytest_dense.log
timeout 600 python bench.py --workload card32 --no-cpu --no-e2e --steps 2 > gpurun_out/r01h_card32_fp64.json 2> gpurun_out/r01h_card32_fp64.err; tail -2 gpurun_out/r01h_card32_fp64.err
timeout 600 python bench.py --workload card32 --precision fp32 --no-cpu --no-e2e --steps 2 > gpurun_out/r01h_card32_fp32.json 2> gpurun_out/r01h_card32_fp32.err
for d in 128 256 512 1024; do
timeout 600 python bench.py --workload dag2000 --dense-min $d --no-cpu --no-e2e --steps 2 > gpurun_out/r01h_dag2000_fp64_d$d.json 2> gpurun_out/r01h_dag2000_fp64_d$d.err
done
timeout 600 python bench.py --workload dag2000 --dense-min 256 --precision fp32 --no-cpu --no-e2e --steps 2 > gpurun_out/r01h_dag2000_fp32_d256.json 2> gpurun_out/r01h_dag2000_fp32_d256.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r01h_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        d=j.get("dense") or {}
        print(f, "value %.4g ms/step %.2f hbm-frac %.3f sweep-kernel ms %.3f | dense ms/sweep %s TF %s frac %s share %s nodes %s"%(j["value"], j["ms_per_step"], j["roofline"]["frac"], j["roofline"]["ms_per_launch"], d.get("ms_per_sweep"), d.get("achieved"), d.get("frac"), d.get("share_of_sweep_time"), d.get("nodes")))
    except Exception as e:
        print(f, "ERR", e)
PY
# one full capture of the dense kernel (card32 fp64, 4096 cases: 32 M-tiles x 976 column tiles)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_gemm -s 2 -c 1 -o gpurun_out/r01h_dense_fp64 python bench.py --workload card32 --cases 4096 --sweeps 3 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/r01h_ncu_fp64.log 2>&1; tail -2 gpurun_out/r01h_ncu_fp64.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_gemm -s 2 -c 1 -o gpurun_out/r01h_dense_fp32 python bench.py --workload card32 --precision fp32 --cases 4096 --sweeps 3 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/r01h_ncu_fp32.log 2>&1; tail -2 gpurun_out/r01h_ncu_fp32.log
ls -la gpurun_out/*.ncu-rep
