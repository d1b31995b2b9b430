"""The "library bar" of SURVEY.md 8(d): the reference's own step (oracle port = the reference's torch operators in the
reference's order, train/dcgan_trainer.py:155-189) run UNCHANGED on the B200 through torch eager + cuDNN / cuBLAS, at the
headline batch.  Usage: python tests/notes/eager_bar.py [batch] [steps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import models, steps as osteps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    dev = torch.device("cuda")
    # (channels_last is not an option for the reference's code: compute_gradient_penalty's .view(B, -1), dcgan_trainer.py:125)
    for name, tf32, cl in (("fp32, TF32 off", False, False), ("fp32 storage, TF32 convs (torch default)", True, False)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        g, d = models.build("DCGAN", seed=12345)
        g, d = g.to(dev), d.to(dev)
        if cl:
            g, d = g.to(memory_format=torch.channels_last), d.to(memory_format=torch.channels_last)
        og, od = osteps.make_optimizers(g, d, 2e-4)
        real = osteps.make_real(B, n_steps=1)[0].to(dev)
        rng = {k: v.to(dev) for k, v in osteps.make_rng(B, n_steps=1, seed=1)[0].items()}
        if cl:
            real = real.contiguous(memory_format=torch.channels_last)
        with torch.device(dev):
            for _ in range(5):
                osteps.dcgan_step(g, d, og, od, real, rng)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                osteps.dcgan_step(g, d, og, od, real, rng)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(f"reference step, torch eager on B200 ({name}): B={B} {ms:.2f} ms/step {B / ms * 1e3:.0f} images/s", flush=True)


if __name__ == "__main__":
    main()
