#!/bin/bash
# One parameterised GPU-box script (replaces the per-call scripts of round 1).  Run through gpurun:
#   gpurun --timeout 1500 -- 'bash scripts/gpu_run.sh r02a tests bench "ncu:sweep_kernel:--workload dag2000 --cases 4096"'
# First argument = tag (file prefix under gpurun_out/), then any number of steps:
#   tests[:pytest args]         pytest -m gpu (default: the whole tests/ directory)
#   bench[:name:bench args]     python bench.py <args>      -> <tag>_bench_<name>.json / .err
#   ref[:bench args]            python bench.py --impl reference <args>
#   torchrun:N:name:bench args  N ranks on one box through torch.distributed.run
#   launches:name:bench args    ncu launch list (gpu__time_duration + dram bytes) of a short bench run
#   ncu:regex:name:skip:count:bench args   ncu --set full capture of `count` launches of the kernels matching regex, after `skip`
#   sanitizer:tool:pytest args  compute-sanitizer --tool <memcheck|racecheck> over a pytest selection
#   cmd:name:shell command      anything else, logged to <tag>_<name>.log
set -u
tag=$1; shift
mkdir -p gpurun_out
out=gpurun_out/$tag
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > ${out}_devices.txt 2>&1
for step in "$@"; do
  kind=${step%%:*}; rest=${step#*:}; [ "$rest" == "$step" ] && rest=""
  echo "=== $step"
  case $kind in
    tests)
      timeout 1500 python -m pytest ${rest:-tests} -m gpu -x -q --durations=15 > ${out}_pytest_gpu.log 2>&1; tail -n 24 ${out}_pytest_gpu.log ;;
    bench)
      name=${rest%%:*}; args=${rest#*:}; [ "$args" == "$rest" ] && args=""; name=${name:-default}
      timeout 900 python bench.py $args > ${out}_bench_${name}.json 2> ${out}_bench_${name}.err; cut -c1-400 ${out}_bench_${name}.json; tail -n 3 ${out}_bench_${name}.err ;;
    ref)
      timeout 900 python bench.py --impl reference $rest > ${out}_bench_reference.json 2> ${out}_bench_reference.err; cut -c1-300 ${out}_bench_reference.json ;;
    torchrun)
      n=${rest%%:*}; rest=${rest#*:}; name=${rest%%:*}; args=${rest#*:}
      timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n $args > ${out}_bench_${name}.json 2> ${out}_bench_${name}.err; cut -c1-400 ${out}_bench_${name}.json; tail -n 3 ${out}_bench_${name}.err ;;
    launches)
      name=${rest%%:*}; args=${rest#*:}
      timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file ${out}_launches_${name}.csv python bench.py $args > ${out}_launches_${name}.log 2>&1; tail -n 2 ${out}_launches_${name}.log ;;
    ncu)
      regex=${rest%%:*}; rest=${rest#*:}; name=${rest%%:*}; rest=${rest#*:}; skip=${rest%%:*}; rest=${rest#*:}; count=${rest%%:*}; args=${rest#*:}
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count -f -o ${out}_ncu_${name} python bench.py $args > ${out}_ncu_${name}.log 2>&1; tail -n 2 ${out}_ncu_${name}.log ;;
    sanitizer)
      tool=${rest%%:*}; args=${rest#*:}
      timeout 1200 compute-sanitizer --tool $tool python -m pytest $args -m gpu -x -q > ${out}_${tool}.log 2>&1; tail -n 8 ${out}_${tool}.log ;;
    cmd)
      name=${rest%%:*}; c=${rest#*:}
      timeout 1200 bash -c "$c" > ${out}_${name}.log 2>&1; tail -n 12 ${out}_${name}.log ;;
    *) echo "unknown step $step" ;;
  esac
done
ls -la gpurun_out | tail -n 30
