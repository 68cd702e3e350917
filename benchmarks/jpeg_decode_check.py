"""GPU JPEG decode: which nvJPEG backends this GPU offers, are their pixels right, how fast are they.

For every backend (hardware = NVJPG engines, gpu_hybrid, hybrid = CPU Huffman) of the native decoder (csrc/jpeg_decode.cpp)
and for the torchvision fallback: decode + resize a batch of synthetic JPEGs to 224x224, compare with the CPU loader's
PIL decode (mean / max absolute pixel difference), then time `--iters` batches (CUDA events, after warm-up).
One JSON line per backend, prefix JPEG_DECODE.

    python benchmarks/jpeg_decode_check.py --batch 256 --stored 500x375
"""
import argparse
import json
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from b200ddl import ops
from b200ddl.data import synthetic_images
from b200ddl.models import decode_image


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--stored", default="500x375")
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--backends", default="hardware,gpu_hybrid,hybrid,torchvision")
    args = ap.parse_args()
    sw, sh = (int(v) for v in args.stored.split("x"))
    pdf = synthetic_images(args.batch, size=(sh, sw), jpeg=True, seed=5).to_pandas()
    blobs = [np.frombuffer(b, dtype=np.uint8).copy() for b in pdf["content"]]   # own the memory: stable addresses
    ref = np.stack([decode_image(bytes(b), (args.size, args.size)) for b in blobs[:32]]).astype(np.int32)
    ptrs = [int(b.ctypes.data) for b in blobs]
    lens = [int(b.size) for b in blobs]
    dev = torch.device("cuda", 0)
    out = torch.empty(args.batch, args.size, args.size, 3, device=dev, dtype=torch.uint8)
    jpeg = ops.ext("_b200_jpeg")
    e = ops.ext("_b200_ops")

    def report(name, run, extra):
        rec = {"backend": name, "batch": args.batch, "stored": args.stored, **extra}
        try:
            out.zero_()
            run()
            torch.cuda.synchronize()
            d = np.abs(out[:32].cpu().numpy().astype(np.int32) - ref)
            rec.update(mean_abs_diff=round(float(d.mean()), 3), max_abs_diff=int(d.max()))
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            rec.update(ms_per_batch=round(ms, 3), images_per_sec=round(args.batch / ms * 1e3, 1))
        except Exception as ex:
            rec["error"] = str(ex).splitlines()[0][:300]
            traceback.print_exc()
        print("JPEG_DECODE " + json.dumps(rec), flush=True)

    want = args.backends.split(",")
    for name in [b for b in ("hardware", "gpu_hybrid", "hybrid") if b in want]:
        try:
            dec = jpeg.JpegDecoder(args.batch, name, 8, 0)
        except Exception as ex:
            print("JPEG_DECODE " + json.dumps({"backend": name, "error": "create: " + str(ex).splitlines()[0][:300]}), flush=True)
            continue
        report("nvjpeg:" + name, lambda: dec.decode_resize(ptrs, lens, out), {"hardware_info": dec.hardware_info()})
        del dec

    if "torchvision" not in want:
        return
    import torchvision

    data = [torch.from_numpy(b) for b in blobs]

    def tv():
        imgs = torchvision.io.decode_jpeg(data, device=dev, mode=torchvision.io.ImageReadMode.RGB)
        for i, img in enumerate(imgs):
            e.resize_bilinear_u8(img.unsqueeze(0), out[i:i + 1], True)

    report("torchvision+resize_bilinear_u8", tv, {})


if __name__ == "__main__":
    main()
