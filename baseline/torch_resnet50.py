"""Operative baseline: the reference's NCCL path reproduced in plain PyTorch - tuned, not a strawman.

ResNet-50 (torchvision architecture, random init), bf16 autocast, channels_last, cuDNN/cuBLAS kernels with
`cudnn.benchmark`, fused SGD+momentum, gradients living in ONE flat buffer (``p.grad`` are views, so there is no
concatenate / copy-back), Horovod-style fusion buckets all-reduced with ``torch.distributed.all_reduce`` (NCCL) on a
side stream AS SOON AS the bucket's last gradient has been accumulated (overlapped with the rest of backward),
followed by the separate divide kernel - what `hvd.DistributedOptimizer` does (reference: P1/03:301-302) - and the
whole step captured in a CUDA graph (``--graph``; bench.py retries without it if capture is not possible).
Nothing from b200ddl is on this path.  The SM clocks are sampled during the timed region (same sampler as bench.py).

    python baseline/torch_resnet50.py --steps 20 --warmup 5 [--graph]
    torchrun --nproc-per-node N --master-addr 127.0.0.1 baseline/torch_resnet50.py ...
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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
    ap.add_argument("--no-overlap", action="store_true", help="all-reduce after backward instead of from gradient hooks")
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
    params = [p for p in model.parameters()]
    try:
        opt = torch.optim.SGD(params, lr=0.1 * world, momentum=0.9, weight_decay=1e-4, fused=True)
        opt_kind = "fused"
    except (TypeError, RuntimeError):
        opt = torch.optim.SGD(params, lr=0.1 * world, momentum=0.9, weight_decay=1e-4, foreach=True)
        opt_kind = "foreach"
    if world > 1:
        for p in params:
            dist.broadcast(p.data, 0)

    # ONE flat gradient buffer in reverse registration (= backward completion) order; p.grad are views into it
    order = list(reversed(params))
    total = sum(p.numel() for p in order)
    flat = torch.zeros(total, device=dev)
    off = 0
    spans = {}
    for p in order:
        # same strides as the (channels_last) parameter: the fused / foreach optimizers require matching layouts
        p.grad = flat[off:off + p.numel()].as_strided(p.shape, p.stride())
        spans[p] = (off, off + p.numel())
        off += p.numel()
    # Horovod-style fusion buckets = contiguous slices of the flat buffer
    budget = int(args.bucket_mb * 2 ** 20 / 4)
    buckets = []  # [lo, hi, n_params]
    lo = 0
    cnt = 0
    for p in order:
        cnt += 1
        hi = spans[p][1]
        if hi - lo >= budget:
            buckets.append([lo, hi, cnt])
            lo, cnt = hi, 0
    if cnt:
        buckets.append([lo, total, cnt])
    bucket_of = {}
    bi = 0
    for p in order:
        while spans[p][0] >= buckets[bi][1]:
            bi += 1
        bucket_of[p] = bi
    pending = [b[2] for b in buckets]
    comm = torch.cuda.Stream(device=dev) if world > 1 else None

    def reduce_bucket(i):
        comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(comm):
            view = flat[buckets[i][0]:buckets[i][1]]
            dist.all_reduce(view)
            view.div_(world)  # separate scale kernel, as in Horovod

    if world > 1 and not args.no_overlap:
        def hook(p):
            i = bucket_of[p]
            pending[i] -= 1
            if pending[i] == 0:
                reduce_bucket(i)

        for p in params:
            p.register_post_accumulate_grad_hook(hook)

    x_u8 = torch.randint(0, 256, (args.batch, 224, 224, 3), dtype=torch.uint8, device=dev)
    labels = torch.randint(0, args.classes, (args.batch,), device=dev)

    def step():
        x = (x_u8.permute(0, 3, 1, 2).float() / 127.5 - 1.0).contiguous(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = torch.nn.functional.cross_entropy(model(x).float(), labels)
        flat.zero_()
        for i, b in enumerate(buckets):
            pending[i] = b[2]
        loss.backward()
        if world > 1:
            if args.no_overlap:
                for i in range(len(buckets)):
                    reduce_bucket(i)
            torch.cuda.current_stream().wait_stream(comm)
        opt.step()
        return loss

    # probe step: if the fused optimizer rejects the tensor layout on this torch build, fall back to the foreach one
    try:
        step()
        torch.cuda.synchronize()
    except RuntimeError as ex:
        if opt_kind != "fused":
            raise
        print(f"# fused SGD unavailable ({str(ex)[:120]}); using foreach", file=sys.stderr)
        opt = torch.optim.SGD(params, lr=0.1 * world, momentum=0.9, weight_decay=1e-4, foreach=True)
        opt_kind = "foreach"
        step()
        torch.cuda.synchronize()

    if args.graph:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            static_loss = step()  # noqa: F841
        run = g.replay
    else:
        run = step

    for _ in range(args.warmup):
        run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = None
    if rank == 0:
        try:
            sys.path.insert(0, ROOT)
            from bench import ClockSampler  # stdlib-only helper (nvidia-smi poller); nothing of the engine is imported

            sampler = ClockSampler(torch.cuda.current_device())
            sampler.start()
        except Exception:
            sampler = None
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler is not None else None
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        ips = args.batch * world * args.steps / (ms / 1e3)
        print(json.dumps({"impl": "torch-nccl-baseline", "metric": "resnet50_train_images_per_sec", "value": ips,
                          "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms / args.steps, "graph": bool(args.graph), "dtype": "bf16", "clocks": clocks,
                          "config": {"model": "resnet50", "global_batch": args.batch * world, "classes": args.classes,
                                     "memory_format": "channels_last", "optimizer": f"SGD({opt_kind})",
                                     "allreduce": "NCCL all_reduce per %.0f MB bucket + div, %s" % (
                                         args.bucket_mb, "after backward" if args.no_overlap else "overlapped from grad hooks"),
                                     "torch": torch.__version__}}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
