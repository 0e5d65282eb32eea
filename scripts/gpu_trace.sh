set -x
mkdir -p gpurun_out
BNBP_TRACE=1 timeout 300 python bench.py --no-cpu --steps 2 > gpurun_out/r01g_trace_fp64.json 2> gpurun_out/r01g_trace_fp64.err; grep "chunk\|done" gpurun_out/r01g_trace_fp64.err | tail -24
BNBP_TRACE=1 timeout 300 python bench.py --no-cpu --steps 2 --precision fp32 > gpurun_out/r01g_trace_fp32.json 2> gpurun_out/r01g_trace_fp32.err; grep "chunk\|done" gpurun_out/r01g_trace_fp32.err | tail -12
