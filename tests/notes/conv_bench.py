"""Micro-benchmark (GPU): the tcgen05 conv kernels per layer shape at the bench batch, with CUDA events.
    python tests/notes/conv_bench.py [B]
Prints us per call and TFLOP/s for: plain, +stats epilogue, +fused BatchNorm-backward epilogue."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry
entry.build()
from jck_generation_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dt = torch.bfloat16
dev = "cuda"


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
    return e0.elapsed_time(e1) / n * 1e3


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, (Ca, Cb, Hs) in {"c2": (128, 64, 16), "c3": (256, 128, 8), "c4": (512, 256, 4)}.items():
    small = torch.randn(B, Hs, Hs, Ca, device=dev).to(dt)
    large = torch.randn(B, 2 * Hs, 2 * Hs, Cb, device=dev).to(dt)
    w4 = (torch.randn(Ca, Cb, 4, 4, device=dev) * 0.05)
    wd = torch.empty(Ca * 16 * Cb, dtype=dt, device=dev)
    wu = torch.empty(Ca * 16 * Cb, dtype=dt, device=dev)
    ops.pack_weights(w4, wd, wu)
    out_s, out_l = torch.empty_like(small), torch.empty_like(large)
    st_s, st_l = torch.zeros(1, 2 * Ca, device=dev), torch.zeros(1, 2 * Cb, device=dev)
    ss_s, mr_s = torch.randn(1, 2 * Ca, device=dev), torch.rand(1, 2 * Ca, device=dev)
    ss_l, mr_l = torch.randn(1, 2 * Cb, device=dev), torch.rand(1, 2 * Cb, device=dev)
    flop = 2.0 * B * Hs * Hs * 16 * Ca * Cb
    rows = {
        "down plain": lambda: ops.conv_down(large, wd, out_s, None, Ca, Cb),
        "down stats": lambda: ops.conv_down(large, wd, out_s, st_s, Ca, Cb),
        "down bnbwd": lambda: ops.conv_down_bnbwd(large, wd, small, ss_s, mr_s, 0.0, out_s, st_s, Ca, Cb),
        "up   plain": lambda: ops.conv_up(small, wu, out_l, None, Ca, Cb),
        "up   stats": lambda: ops.conv_up(small, wu, out_l, st_l, Ca, Cb),
        "up   bnbwd": lambda: ops.conv_up_bnbwd(small, wu, large, ss_l, mr_l, 0.2, out_l, st_l, Ca, Cb),
    }
    nb = ops.wgrad_workspace_bytes(B, Hs, Hs, Ca, Cb, dt)
    ws = torch.empty(nb // 4, device=dev)
    dw = torch.empty(Ca, Cb, 4, 4, device=dev)
    rows["wgrad     "] = lambda: ops.conv_wgrad(small, large, dw, ws, Ca, Cb, False)
    for k, fn in rows.items():
        us = timeit(fn)
        print(f"{name} Ca={Ca} Cb={Cb} Hs={Hs} B={B}  {k}: {us:8.1f} us  {flop / us / 1e6:7.1f} TFLOP/s", flush=True)
