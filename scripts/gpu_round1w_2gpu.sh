# round 1w: the N>1 path on real NCCL (2 GPUs of one box): weak scaling bench, gather variant, reference arm under torchrun
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r01w_devices.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r01w_bench_2gpu.json 2> gpurun_out/r01w_bench_2gpu.err; cut -c1-400 gpurun_out/r01w_bench_2gpu.json; tail -n 3 gpurun_out/r01w_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-e2e --gather > gpurun_out/r01w_bench_2gpu_gather.json 2> gpurun_out/r01w_bench_2gpu_gather.err; cut -c1-300 gpurun_out/r01w_bench_2gpu_gather.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-e2e --epsilon 1e-6 > gpurun_out/r01w_bench_2gpu_eps.json 2> gpurun_out/r01w_bench_2gpu_eps.err; cut -c1-300 gpurun_out/r01w_bench_2gpu_eps.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r01w_bench_reference_n2.json 2> gpurun_out/r01w_bench_reference_n2.err; cut -c1-300 gpurun_out/r01w_bench_reference_n2.json
timeout 300 python bench.py --no-cpu --steps 10 > gpurun_out/r01w_bench_1gpu.json 2>/dev/null; cut -c1-200 gpurun_out/r01w_bench_1gpu.json
