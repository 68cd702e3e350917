"""BatchNorm apply / backward-apply kernels with 1, 2 and 4 rows per loop iteration (`set_bn_rows_unroll`): the outputs
must be BIT-IDENTICAL (same per-element arithmetic, only the load schedule changes); timings with an L2 flush between launches.
Lines: BN_UNROLL ...      python benchmarks/bn_unroll_check.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from b200ddl import ops

e = ops.ext("_b200_ops")
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
l2 = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timed(fn, iters=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        l2.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


ok_all = True
tot = {1: 0.0, 2: 0.0, 4: 0.0}
# (rows, channels, how many such layers a ResNet-50 step has) - batch 256 activations, plus one ragged shape
for (M, C, weight) in ((802816, 64, 7), (802816, 256, 4), (200704, 128, 9), (200704, 512, 5), (50176, 256, 13), (50176, 1024, 7),
                       (12544, 512, 7), (12544, 2048, 4), (3001, 72, 0)):
    y = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
    res = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
    gr = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
    sc = torch.rand(C, device=dev, generator=g) + 0.5
    sh = torch.randn(C, device=dev, generator=g) * 0.3
    cA, cB, cC = (torch.randn(C, device=dev, generator=g) * 0.5 for _ in range(3))
    cases = {
        "apply_relu": lambda o, m: e.bn_apply(y, sc, sh, None, None, None, o, True, m),
        "apply_res_relu": lambda o, m: e.bn_apply(y, sc, sh, res, None, None, o, True, m),
        "apply_res2": lambda o, m: e.bn_apply(y, sc, sh, res, sc, sh, o, False, None),
        "bwd_apply_mask": lambda o, m: e.bn_bwd_apply(gr, y, sc, sh, cA, cB, cC, o),
        "bwd_apply": lambda o, m: e.bn_bwd_apply(gr, y, None, None, cA, cB, cC, o),
    }
    for name, fn in cases.items():
        outs, masks, times = {}, {}, {}
        for u in (1, 2, 4):
            e.set_bn_rows_unroll(u)
            o = torch.zeros_like(y)
            m = torch.zeros(M * C // 8, device=dev, dtype=torch.uint8)
            fn(o, m)
            torch.cuda.synchronize()
            outs[u], masks[u] = o, m
            times[u] = timed(lambda: fn(o, m))
            if name in ("apply_res_relu", "bwd_apply_mask"):
                tot[u] += times[u] * weight
        same = all(torch.equal(outs[1], outs[u]) and torch.equal(masks[1], masks[u]) for u in (2, 4))
        ok_all &= same
        gb = (3 if ("res" in name or "bwd" in name) else 2) * M * C * 2 / 1e9
        print("BN_UNROLL " + json.dumps({"case": name, "M": M, "C": C, "identical": same,
                                         "us": {str(u): round(t, 1) for u, t in times.items()},
                                         "TBs": {str(u): round(gb / t * 1e3, 2) for u, t in times.items()}}), flush=True)
e.set_bn_rows_unroll(1)
print("BN_UNROLL " + json.dumps({"all_identical": ok_all, "weighted_step_us(res_relu + bwd_mask)": {str(u): round(t) for u, t in tot.items()}}),
      flush=True)
sys.exit(0 if ok_all else 1)
