# the N=8 path on one box: weak scaling bench as the driver launches it
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r01x4_bench_4gpu.json 2> gpurun_out/r01x4_bench_4gpu.err; cut -c1-300 gpurun_out/r01x4_bench_4gpu.json; tail -n 3 gpurun_out/r01x4_bench_4gpu.err
