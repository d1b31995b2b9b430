"""Micro-benchmark (GPU): the BatchNorm streaming passes, 20 back-to-back launches per measurement (host launch cost hidden),
tensors larger than L2.  Usage: python tests/notes/bn_bench2.py [images]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry
entry.build()
from jck_generation_b200 import ops

dt, dev = torch.bfloat16, "cuda"
imgs_list = [int(a) for a in sys.argv[1:]] or [1024]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


for imgs in imgs_list:
    for C, H in ((64, 32), (128, 16), (256, 8)):
        y = torch.randn(imgs, H, H, C, device=dev).to(dt)
        da = torch.randn_like(y)
        out = torch.empty_like(y)
        ss, mr = torch.randn(1, 2 * C, device=dev), torch.rand(1, 2 * C, device=dev) + 0.5
        gamma = torch.rand(C, device=dev) + 0.5
        sums = torch.zeros(1, 2 * C, device=dev)
        n = y.numel()
        rows = {"fwd   ": (lambda: ops.bn_act_fwd(y, ss, out, C, 1, 0.2), 4 * n),
                "reduce": (lambda: ops.bn_act_bwd_reduce(da, y, ss, mr, sums, C, 1, 0.2), 4 * n),
                "apply ": (lambda: ops.bn_act_bwd_apply(da, y, ss, mr, gamma, sums, out, C, 1, n // C, 0.2), 6 * n)}
        for k, (fn, nbytes) in rows.items():
            us = timeit(fn)
            print(f"imgs={imgs:5d} C={C:3d} H={H:2d} {k}: {us:7.1f} us  {nbytes / us / 1e3:7.1f} GB/s  ({nbytes / 1e6:6.1f} MB)", flush=True)
