"""Operative baseline: the reference's NCCL path reproduced in plain PyTorch.

ResNet-50 (torchvision architecture, random init), bf16 autocast, channels_last, cuDNN/cuBLAS kernels, SGD+momentum,
gradient all-reduce = ``torch.distributed.all_reduce`` per bucket (NCCL) followed by a separate divide kernel - what
Horovod's DistributedOptimizer does (reference: P1/03:301-302).  Nothing from b200ddl is on this path.

    python baseline/torch_resnet50.py --steps 20 --warmup 5 [--graph]
    torchrun --nproc-per-node N --master-addr 127.0.0.1 baseline/torch_resnet50.py ...
"""
from __future__ import annotations

import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def build(num_classes: int):
    import torchvision

    return torchvision.models.resnet50(weights=None, num_classes=num_classes)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--graph", action="store_true", help="capture the step in a CUDA graph")
    ap.add_argument("--bucket-mb", type=float, default=25.0)
    args = ap.parse_args(argv)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda", local)
    model = build(args.classes).to(dev).to(memory_format=torch.channels_last)
    opt = torch.optim.SGD(model.parameters(), lr=0.1 * world, momentum=0.9, weight_decay=1e-4)
    params = [p for p in model.parameters()]
    if world > 1:
        for p in params:
            dist.broadcast(p.data, 0)

    # Horovod-style fusion buckets in reverse registration order
    buckets, cur, cur_bytes = [], [], 0
    for p in reversed(params):
        cur.append(p)
        cur_bytes += p.numel() * 4
        if cur_bytes >= args.bucket_mb * 2 ** 20:
            buckets.append(cur)
            cur, cur_bytes = [], 0
    if cur:
        buckets.append(cur)

    x_u8 = torch.randint(0, 256, (args.batch, 224, 224, 3), dtype=torch.uint8, device=dev)
    labels = torch.randint(0, args.classes, (args.batch,), device=dev)

    def step():
        x = (x_u8.permute(0, 3, 1, 2).float() / 127.5 - 1.0).contiguous(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = torch.nn.functional.cross_entropy(model(x).float(), labels)
        opt.zero_grad(set_to_none=False)
        loss.backward()
        if world > 1:
            for b in buckets:
                flat = torch.cat([p.grad.reshape(-1) for p in b])
                dist.all_reduce(flat)
                flat.div_(world)
                off = 0
                for p in b:
                    p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
                    off += p.numel()
        opt.step()
        return loss

    if args.graph and world == 1:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            static_loss = step()
        run = g.replay
    else:
        run = step

    for _ in range(args.warmup):
        run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        ips = args.batch * world * args.steps / (ms / 1e3)
        print(json.dumps({"impl": "torch-nccl-baseline", "metric": "resnet50_train_images_per_sec", "value": ips,
                          "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms / args.steps, "graph": bool(args.graph), "dtype": "bf16",
                          "config": {"model": "resnet50", "global_batch": args.batch * world, "classes": args.classes}}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
