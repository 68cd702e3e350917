#!/bin/bash
# Round-2 last call (2 GPUs, ~3 min of box time left): the final tree's data-parallel step, then training from real JPEG bytes
# with GPU decode on both ranks (each GPU decodes its own shard: this path scales with the GPUs, the CPU path does not)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 120 $TR --master-port 29641 bench.py --gpus 2 --steps 20 --warmup 5 --no-baseline > $O/bench_w2_final.log 2>&1; echo "bench w2 rc=$?"
grep '^{' $O/bench_w2_final.log | tail -1 | cut -c1-700
timeout 110 $TR --master-port 29661 benchmarks/loader_jpeg_bench.py --images 1024 --stored 500x375 --batch 256 --steps 20 --warmup 3 --procs 24 --proc-threads 2 --modes gpu,procs > $O/loader_jpeg_w2.log 2>&1; echo "loader jpeg w2 rc=$?"
grep -E "LOADER_JPEG|Error" $O/loader_jpeg_w2.log | cut -c1-600
