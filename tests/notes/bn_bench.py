"""Micro-benchmark (GPU): the streaming BatchNorm passes per layer size, GB/s against the HBM copy peak."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry
entry.build()
from jck_generation_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dt, dev = torch.bfloat16, "cuda"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()                      # evict L2 (126 MB) so every pass streams from HBM as in the real step
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for imgs in (B, 2 * B, 3 * B):
    for C, H in ((64, 32), (128, 16), (256, 8), (512, 4)):
        groups = imgs // B
        y = torch.randn(imgs, H, H, C, device=dev).to(dt)
        da = torch.randn_like(y)
        out = torch.empty_like(y)
        ss, mr = torch.randn(groups, 2 * C, device=dev), torch.rand(groups, 2 * C, device=dev) + 0.5
        gamma = torch.rand(C, device=dev) + 0.5
        sums = torch.zeros(groups, 2 * C, device=dev)
        n = y.numel()
        cnt = n // C // groups
        rows = {"fwd   ": (lambda: ops.bn_act_fwd(y, ss, out, C, groups, 0.2), 4 * n),
                "reduce": (lambda: ops.bn_act_bwd_reduce(da, y, ss, mr, sums, C, groups, 0.2), 4 * n),
                "apply ": (lambda: ops.bn_act_bwd_apply(da, y, ss, mr, gamma, sums, out, C, groups, cnt, 0.2), 6 * n)}
        for k, (fn, nbytes) in rows.items():
            us = timeit(fn)
            print(f"imgs={imgs:5d} C={C:3d} H={H:2d} {k}: {us:7.1f} us  {nbytes / us / 1e3:7.1f} GB/s  ({nbytes / 1e6:6.1f} MB)", flush=True)
