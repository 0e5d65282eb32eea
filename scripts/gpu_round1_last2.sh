set -x
mkdir -p gpurun_out
timeout 600 python scripts/next_rows_bench.py > gpurun_out/r01last_next_rows.jsonl 2> gpurun_out/r01last_next_rows.err; cat gpurun_out/r01last_next_rows.jsonl; tail -n 3 gpurun_out/r01last_next_rows.err
timeout 600 python -m pytest tests/test_dropin_cpp.py tests/test_lw.py -m gpu -q 2>&1 | tail -n 2
