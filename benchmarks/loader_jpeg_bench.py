"""Training throughput when the input is REAL JPEG bytes in the converter's parquet cache (reference
P1/03:137-144 converter, :182-189 decode + resize, :332-348 make_tf_dataset + fit) - not the raw-tensor synthetic ring.

For each decode mode it reports (one JSON line per mode, prefix LOADER_JPEG):
  loader_only_img_s   batches pulled from the dataset with no training step (what the pipeline can deliver)
  train_img_s         `Trainer.fit` over the dataset, wall clock with a device sync on both sides (what the user gets)
  loader_wait_ms      time fit() spent waiting for a READY slot (pinned-ring datasets)
modes: 'cpu'  PIL decode (+draft-mode downscale) + resize on `--workers` threads into the pinned ring, side-stream H2D
       'procs' the same decode in `--procs` worker PROCESSES (loader/_decode_worker.py) driven by `--proc-threads` filler threads
       'gpu'  nvJPEG (library) decode on the device + our bilinear resize kernel; only compressed bytes cross PCIe
       'ring' the synthetic raw-tensor ring (upper bound of the host path; what bench.py's e2e arm uses)

    python benchmarks/loader_jpeg_bench.py --images 4096 --stored 320x256 --batch 256 --steps 30 --workers 32
"""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from b200ddl import Session, optim
from b200ddl.data import col, pandas_udf, synthetic_images
from b200ddl.loader import SyntheticDataset, make_converter
from b200ddl.models import CLASSES, build_model
from b200ddl.train import Trainer


def _sync(dev):
    if dev.type == "cuda":
        torch.cuda.synchronize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=4096)
    ap.add_argument("--stored", default="320x256", help="WxH of the stored JPEGs (tf_flowers photos are ~500x333 or smaller)")
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workers", type=int, default=32)
    ap.add_argument("--arch", default="resnet50")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--procs", type=int, default=96, help="decode processes of the 'procs' mode")
    ap.add_argument("--proc-threads", type=int, default=8, help="filler threads driving the decode processes")
    ap.add_argument("--modes", default="ring,cpu,procs,gpu")
    args = ap.parse_args()
    # under torchrun (WORLD_SIZE > 1): data-parallel training, every rank reads ITS shard of the same cache and decodes it itself
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = 0
    hvd = None
    if world > 1:
        import b200ddl.parallel as hvd

        hvd.init()
        rank = hvd.rank()
        args.device = str(hvd.device())
    dev = torch.device(args.device)
    sw, sh = (int(v) for v in args.stored.split("x"))
    root = tempfile.mkdtemp(prefix="b200ddl_jpeg_")
    Session(user="loader@example.com", root=root)

    t0 = time.time()
    raw = synthetic_images(args.images, size=(sh, sw), jpeg=True, seed=11)

    @pandas_udf("int")
    def label_idx(path):
        return path.map(lambda p: CLASSES.index(p.split("/")[-2]))

    table = raw.withColumn("label_idx", label_idx(col("path"))).select(["content", "label_idx"])
    conv = make_converter(table, os.path.join(root, "cache"))
    jpeg_bytes = sum(len(b) for b in table.limit(256).to_pandas()["content"]) / min(256, args.images)
    print(f"dataset: {args.images} JPEGs {sw}x{sh}, {jpeg_bytes / 1024:.1f} KB each, built in {time.time() - t0:.1f} s", flush=True)

    model = build_model(args.size, args.size, 3, len(CLASSES), arch=args.arch, batch_size=args.batch, freeze_base=False)
    opt = optim.SGD(learning_rate=0.01, momentum=0.9)
    if world > 1:
        opt = hvd.DistributedOptimizer(opt)
    trainer = Trainer(model, device=args.device).compile(optimizer=opt, loss="sparse_categorical_crossentropy", metrics=["accuracy"])
    shard = dict(cur_shard=rank, shard_count=world) if world > 1 else {}

    def make(mode):
        if mode == "ring":
            return SyntheticDataset(args.batch, num_classes=len(CLASSES), device=dev, threads=6, pool_images=2048, seed=3,
                                    image_size=(args.size, args.size), **shard) if dev.type == "cuda" else None
        if mode == "procs":
            return conv.make_dataset(batch_size=args.batch, num_epochs=None, workers_count=args.proc_threads,
                                     image_size=(args.size, args.size), device=dev, decode="cpu",
                                     decode_processes=max(1, args.procs // world), **shard)
        return conv.make_dataset(batch_size=args.batch, num_epochs=None, workers_count=args.workers,
                                 image_size=(args.size, args.size), device=dev, decode=mode, decode_processes=0, **shard)

    for mode in args.modes.split(","):
        if mode == "gpu" and dev.type != "cuda":
            continue
        ds = make(mode)
        if ds is None:
            continue
        with ds:
            it = iter(ds)
            for _ in range(args.warmup):
                next(it)
            _sync(dev)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                x, y = next(it)
            _sync(dev)
            loader_only = args.batch * args.steps / (time.perf_counter() - t0)
            trainer.fit(ds, steps_per_epoch=args.warmup, epochs=1, verbose=0)
            _sync(dev)
            ring = getattr(ds, "ring", None)
            w0 = ring.consumer_wait_ms if ring is not None else 0.0
            t0 = time.perf_counter()
            hist = trainer.fit(ds, steps_per_epoch=args.steps, epochs=1, verbose=0)
            _sync(dev)
            dt = time.perf_counter() - t0
            if world > 1:   # the job's rate: all ranks' images over the slowest rank's time
                dt = max(hvd.allgather_object(dt))
                loader_only = sum(hvd.allgather_object(loader_only))
            rec = {"mode": mode, "images": args.images, "stored": args.stored, "jpeg_kb": round(jpeg_bytes / 1024, 1),
                   "batch": args.batch, "steps": args.steps, "workers": args.workers if mode == "cpu" else (args.proc_threads if mode == "procs" else None),
                   "decode_processes": getattr(ds, "decode_processes", None),
                   "n_gpus": world, "loader_only_img_s": round(loader_only, 1), "train_img_s": round(world * args.batch * args.steps / dt, 1),
                   "train_ms_per_step": round(dt / args.steps * 1e3, 3),
                   "loader_wait_ms": round(ring.consumer_wait_ms - w0, 2) if ring is not None else None,
                   "gpu_decoded": getattr(ds, "gpu_decoded", None), "cpu_decoded": getattr(ds, "cpu_decoded", None),
                   "loss": hist.history["loss"][-1], "arch": args.arch, "cpus": os.cpu_count()}
            if rank == 0:
                print("LOADER_JPEG " + json.dumps(rec), flush=True)
    conv.delete()
    if world > 1:
        del trainer, model, opt
        hvd.shutdown()


if __name__ == "__main__":
    main()
