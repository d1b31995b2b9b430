"""Two eager forwards of the Inception-v3 extractor at B=128 for an ncu capture of the conv_gemm launches:
  ncu --set full --clock-control none --import-source on -k regex:conv_gemm --launch-skip 95 --launch-count 40 \
      -o gpurun_out/prof_incep_r1 python tests/notes/incep_ncu.py
(95 conv_gemm launches per forward: 94 convolutions + fc; the first forward is the warm-up)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torchvision import models
from jck_generation_b200.inception import InceptionV3

torch.manual_seed(12345)
net = InceptionV3(models.inception_v3(weights=None, aux_logits=True, init_weights=False).state_dict(), feature="pool3",
                  device="cuda", use_graph=False)
fake = torch.tanh(torch.randn(int(sys.argv[1]) if len(sys.argv) > 1 else 128, 3, 64, 64, device="cuda"))
for _ in range(2):
    net.forward_generated(fake)
torch.cuda.synchronize()
print("done")
