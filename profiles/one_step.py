"""The profiled command: the bench workload (DCGAN 3x64x64, 512 images, bf16), N eager steps (no CUDA graph, so
every kernel is its own launch for ncu).  Usage: python profiles/one_step.py [steps]"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
entry.build()
from jck_generation_b200.model import DCGAN
from jck_generation_b200.train.dcgan_trainer import DCGANTrainer


class _Data:
    def get_data_loader(self):
        return [], None


steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
args = argparse.Namespace(epoch=1, max_learning_rate=2e-4, model_path="prof", log_file=0, batch_size=512, num_worker=0,
                          dtype="bf16", cuda_graph=0, metrics=0, save_path=os.path.join(ROOT, "gpurun_out", "prof_save"))
torch.manual_seed(12345)
tr = DCGANTrainer(args, DCGAN.Generator(), DCGAN.Discriminator(), _Data())
real = (torch.rand(512, 3, 64, 64, generator=torch.Generator().manual_seed(1)) * 2 - 1).cuda()
for _ in range(steps):
    tr.step.run(real)
torch.cuda.synchronize()
print("one_step done")
