import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import parity
for nc, b in ((1, 4), (1, 8), (3, 4)):
    e = parity.dcgan_step_parity(torch.float32, batch=b, nc=nc)
    print("nc", nc, "batch", b)
    for k, v in sorted(e.items(), key=lambda kv: -kv[1])[:22]:
        print(f"  {v:.3e} {k}")
