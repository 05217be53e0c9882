#!/bin/bash
# Multi-GPU bench lines (one box, N ranks over NCCL): tools/run_scale.sh N config [steps] -> gpurun_out/r2_bench_<config>_n<N>.json
N=$1; CFG=$2; STEPS=${3:-10}
PORT=$((29500 + RANDOM % 400))
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
  bench.py --gpus $N --config $CFG --steps $STEPS --warmup 3 > gpurun_out/r2_bench_${CFG}_n${N}.json 2> gpurun_out/r2_bench_${CFG}_n${N}.err
echo "rc=$? $(head -c 400 gpurun_out/r2_bench_${CFG}_n${N}.json)"
