#!/bin/bash
# usage (inside `gpurun --gpus N`): bash benchmarks/run_scale.sh N
# Scaling evidence at N GPUs: medium all-reduce sweep vs NCCL, ResNet-50 bench (fused NVLS path, CUDA graph),
# fused all-reduce+SGD variant, sharded batch inference.
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
export NCCL_DEBUG=${NCCL_DEBUG:-WARN}
timeout 400 $TR --master-port 29701 benchmarks/allreduce_check.py --medium --max-mb 1024 > gpurun_out/allreduce_w$N.log 2>&1
echo "allreduce rc=$?"; grep -E "world=|FAIL|ALLREDUCE|f32 |bf16 " gpurun_out/allreduce_w$N.log | tail -16
timeout 300 $TR --master-port 29702 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_w$N.log 2>&1
echo "bench rc=$?"; tail -n 1 gpurun_out/bench_w$N.log | cut -c1-1000
timeout 300 $TR --master-port 29703 bench.py --gpus $N --steps 20 --warmup 5 --fused-update --no-e2e > gpurun_out/bench_w${N}_fusedsgd.log 2>&1
echo "bench(fused allreduce+sgd) rc=$?"; tail -n 1 gpurun_out/bench_w${N}_fusedsgd.log | cut -c1-400
timeout 300 $TR --master-port 29704 benchmarks/inference_bench.py --images 1000000 > gpurun_out/infer_w$N.log 2>&1
echo "inference rc=$?"; tail -n 1 gpurun_out/infer_w$N.log | cut -c1-400
