#!/bin/bash
# usage: bash benchmarks/run_multi_gpu.sh N [quick]   (inside `gpurun --gpus N`)
# 1) fused all-reduce correctness + bandwidth sweep vs NCCL, 2) ResNet-50 bench on N GPUs (ours + NCCL-path A/B),
# 3) torch + NCCL baseline on N GPUs.
cd "$(dirname "$0")/.."
N=${1:-2}
MODE=${2:-full}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
export NCCL_DEBUG=${NCCL_DEBUG:-WARN}
nvidia-smi topo -m > gpurun_out/topo_w$N.txt 2>&1
if [ "$MODE" = "quick" ]; then
  timeout 300 $TR --master-port 29601 benchmarks/allreduce_check.py --quick > gpurun_out/allreduce_w$N.log 2>&1
else
  timeout 600 $TR --master-port 29601 benchmarks/allreduce_check.py --max-mb 1024 > gpurun_out/allreduce_w$N.log 2>&1
fi
echo "allreduce rc=$?"; grep -E "world=|FAIL|ALLREDUCE|f32 |bf16 " gpurun_out/allreduce_w$N.log | tail -30
timeout 400 $TR --master-port 29602 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_w$N.log 2>&1
echo "bench rc=$?"; tail -n 1 gpurun_out/bench_w$N.log | cut -c1-900
# A/B of the communication path only (both eager, no CUDA graph): our fused kernel vs dist.all_reduce + div
timeout 200 $TR --master-port 29603 bench.py --gpus $N --steps 15 --warmup 4 --algo nccl --no-e2e --no-graph > gpurun_out/bench_w${N}_ncclpath.log 2>&1
echo "bench(nccl path, eager) rc=$?"; tail -n 1 gpurun_out/bench_w${N}_ncclpath.log | cut -c1-300
timeout 200 $TR --master-port 29605 bench.py --gpus $N --steps 15 --warmup 4 --no-e2e --no-graph > gpurun_out/bench_w${N}_eager.log 2>&1
echo "bench(fused path, eager) rc=$?"; tail -n 1 gpurun_out/bench_w${N}_eager.log | cut -c1-300
timeout 300 $TR --master-port 29604 baseline/torch_resnet50.py --steps 20 --warmup 5 > gpurun_out/base_w$N.log 2>&1
echo "torch baseline rc=$?"; tail -n 1 gpurun_out/base_w$N.log | cut -c1-300
