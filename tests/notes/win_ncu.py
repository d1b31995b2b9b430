"""A few launches of the windowed 64-channel up conv (conv_up_win_kernel) for ncu.  Usage: python tests/notes/win_ncu.py [B]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry
entry.build()
from jck_generation_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dt, dev = torch.bfloat16, "cuda"
Ca, Cb, Hs = 128, 64, 16
small = torch.randn(B, Hs, Hs, Ca, device=dev).to(dt)
large = torch.randn(B, 2 * Hs, 2 * Hs, Cb, device=dev).to(dt)
w4 = torch.randn(Ca, Cb, 4, 4, device=dev) * 0.05
wd = torch.empty(Ca * 16 * Cb, dtype=dt, device=dev)
wu = torch.empty(Ca * 16 * Cb, dtype=dt, device=dev)
ops.pack_weights(w4, wd, wu)
out = torch.empty(B, 2 * Hs, 2 * Hs, Cb, dtype=dt, device=dev)
st = torch.zeros(1, 2 * Cb, device=dev)
ss, mr = torch.randn(1, 2 * Cb, device=dev), torch.rand(1, 2 * Cb, device=dev)
for _ in range(3):
    ops.conv_up(small, wu, out, None, Ca, Cb)
    ops.conv_up(small, wu, out, st, Ca, Cb)
    ops.conv_up_bnbwd(small, wu, large, ss, mr, 0.2, out, st, Ca, Cb)
torch.cuda.synchronize()
print("done")
