"""Drop-in for the reference's train/dcgan_trainer.py: `DCGANTrainer(args, model_g, model_d, data_pre)`
with `.train()`, `.compute_gradient_penalty(real, fake)`, `.save_model(typ, iters, value, images)`.

The step loop keeps the reference's structure and logging (dcgan_trainer.py:130-239); the step body
(:155-189) is `DCGANStep` -- our kernels.  Differences that are deliberate and additive:
  * five `.item()` host syncs per step (:165,176,188,195,196) become one read-back per 100 steps;
  * `args.dtype` ('bf16' | 'fp32'), `args.cuda_graph`, `args.max_iters` are optional extras;
  * data parallel when launched under torchrun (rank r trains on its rows; SyncBN + averaged grads);
  * matplotlib / torchinfo are optional (absent from this image): plots are skipped without them.
"""
import argparse
import os
import time
from datetime import datetime

import torch
import torch.nn as nn

from ..logger.main_logger import MainLogger
from ..logger.utils import time_to_str
from ..model.DCGAN import weights_init
from ..utils import get_default_device
from .. import ops, parallel
from .dcgan_step import DCGANStep
from .optim import FusedAdam
from .prefetch import DevicePrefetcher
from .trainer import Trainer

try:  # optional, as in the reference's environment
    import matplotlib.pyplot as plt
except Exception:  # pragma: no cover
    plt = None


def _dtype_of(args):
    name = str(getattr(args, "dtype", "bf16")).lower()
    return torch.float32 if name in ("fp32", "float32", "f32") else torch.bfloat16


class DCGANTrainer(Trainer):
    def __init__(self, args: argparse.Namespace, model_g: nn.Module, model_d: nn.Module, data_pre):
        self.logger = MainLogger(args)
        self.device = get_default_device()
        if self.device.type != "cuda":
            raise RuntimeError("DCGANTrainer: no CUDA device; the B200 train step has no CPU fallback")
        self.comm = parallel.init_from_env()
        if self.comm.world_size > 1:
            self.device = torch.device("cuda", torch.cuda.current_device())

        self.epoch = args.epoch
        self.max_lr = args.max_learning_rate
        self.lambda_gp = 10.0

        self.model_g = model_g.to(self.device)
        self.model_d = model_d.to(self.device)
        dtype = _dtype_of(args)
        self.model_g.set_compute(dtype=dtype, comm=self.comm)
        self.model_d.set_compute(dtype=dtype, comm=self.comm)
        self.model_g.apply(weights_init)
        self.model_d.apply(weights_init)
        if self.comm.world_size > 1:   # identical replicas: rank 0's initial weights everywhere
            for t in list(self.model_g.state_dict().values()) + list(self.model_d.state_dict().values()):
                torch.distributed.broadcast(t, src=0)

        self.data_pre = data_pre
        self.train_loader, metric_loader = self.data_pre.get_data_loader()
        self.metric = None
        if metric_loader is not None and getattr(args, "metrics", 1):
            from ..metrics import Metrics
            self.metric = Metrics(metric_loader)

        self.flat_g = parallel.FlatParams(self.model_g)
        self.flat_d = parallel.FlatParams(self.model_d)
        self.optimizer_g = FusedAdam(self.model_g.parameters(), lr=self.max_lr, betas=[0.5, 0.999], flat=self.flat_g)
        self.optimizer_d = FusedAdam(self.model_d.parameters(), lr=self.max_lr, betas=[0.5, 0.999], flat=self.flat_d)
        self.criterion = nn.BCELoss()
        self.step = DCGANStep(self.model_g, self.model_d, self.optimizer_g, self.optimizer_d, self.flat_g,
                              self.flat_d, self.comm, self.lambda_gp, seed=int(getattr(args, "seed", 12345)))
        self.use_graph = bool(getattr(args, "cuda_graph", 0)) and (
            self.comm.world_size == 1 or bool(int(os.environ.get("JCK_DP_GRAPH", "1"))))
        self.max_iters = int(getattr(args, "max_iters", 0))

        datetime_now = args.model_path if getattr(args, "model_path", "") != "" else datetime.now().strftime("%Y%m%d_%H%M%S")
        self.model_save_path = os.path.join('.', 'save', 'dcgan', datetime_now)
        if self.comm.rank == 0:
            os.makedirs(self.model_save_path, exist_ok=True)
        self.logger.debug(f'save path: {self.model_save_path}')

    # -------------------------------------------------------------------------------------------------
    def save_model(self, typ, iters, value, images):
        if self.comm.rank != 0:
            return
        save_path = os.path.join(self.model_save_path, typ)
        os.makedirs(save_path, exist_ok=True)
        for filename in os.listdir(save_path):
            file_path = os.path.join(save_path, filename)
            if os.path.isfile(file_path) and filename.endswith('.pt'):
                os.remove(file_path)
        torch.save({
            'model_g': self.model_g.state_dict(),
            'model_d': self.model_d.state_dict(),
            'optimizer_g': self.optimizer_g.state_dict(),
            'optimizer_d': self.optimizer_d.state_dict()
        }, os.path.join(save_path, f'{iters}_{value:.04f}.pt'))
        if plt is not None:
            import numpy as np
            import torchvision.utils as vutils
            plt.clf()
            plt.axis("off")
            plt.title("fake images")
            plt.imshow(np.transpose(vutils.make_grid(images, padding=2, normalize=True), (1, 2, 0)))
            plt.savefig(os.path.join(save_path, f'{iters}_fake_image.png'))
        self.logger.debug(f'{iters} model save')

    def _gp_counter(self):
        if getattr(self, "_gp_rng_counter", None) is None:
            self._gp_rng_counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        return self._gp_rng_counter

    def compute_gradient_penalty(self, real_data, fake_data, alpha=None):
        """mean((||d D(x_hat)/d x_hat||_2 - 1)^2), x_hat = alpha*real + (1-alpha)*fake  (reference :110-127).
        Runs D forward (train mode) + the input-gradient sweep on our kernels.  The result is a plain
        tensor: in the DCGAN trainer the penalty is logged, never back-propagated (:179-180)."""
        ed = self.model_d.engine()
        B = real_data.size(0)
        if alpha is None:
            # a draw of this public method must neither collide with the step's Philox streams (ids 16*rank + 1..5) nor move
            # the training counter: private counter, stream ids from a disjoint range
            alpha = torch.empty(B, 1, 1, 1, device=self.device)
            ops.rand(alpha, self.step.seed, (1 << 20) + 1, self._gp_counter())
            ops.rng_advance(self._gp_counter(), (B + 3) // 4)
        x_hat = ops.img_alloc(B, ed.nc, 64, 64, ed.dtype, self.device, ed.img_layout)
        ops.prep_image(real_data.detach().contiguous().float(), out_nhwc=x_hat, a1=1.0,
                       x2=fake_data.detach().contiguous().float(), alpha=alpha.reshape(B).contiguous(),
                       layout=ed.img_layout)
        ctx = ed.trunk_forward(x_hat, groups=1)
        ed.head_forward(ctx)
        da4 = ed.head_backward(ctx, mode=1, wgrad=False)
        dx = ed.trunk_backward(ctx, da4, wgrad=False, input_grad=True)
        out = torch.zeros(2, dtype=torch.float32, device=self.device)
        ops.gp_penalty(dx, out)
        return out[0]

    def train_step(self, real_data, rng=None):
        """One G+D step on this rank's rows of the batch; returns the [4,2] device scalar block."""
        if self.use_graph and rng is None:
            if self.step._graph is None or self.step._static.shape[0] != real_data.shape[0]:
                self.step.capture(real_data.shape[0])
            return self.step.replay(real_data)
        return self.step.run(real_data, rng)

    # -------------------------------------------------------------------------------------------------
    def train(self):
        real_images_loader = self.train_loader
        losses_g, losses_d = [], []
        iters = 0
        fixed_noise = torch.empty(64, 100, 1, 1, device=self.device)
        ops.randn(fixed_noise, self.step.seed, 7, None)
        low_fid, high_is = 1e10, 0

        start = time.time()
        self.logger.debug("train start")
        pending = []          # (epoch, i, scalar block) not yet read back

        def flush():
            if not pending:
                return
            self.comm.check_health()
            block = torch.stack([p[2] for p in pending])             # one D2H for up to 100 steps
            self.comm.allreduce_mean_(block)
            host = block.cpu()
            for (ep, i, _), s in zip(pending, host):
                m = DCGANStep.summarize(s, self.lambda_gp)
                losses_g.append(m["loss_g"])
                losses_d.append(m["loss_d"])
                if i % 100 == 0:
                    self.logger.debug(f'[{ep}/{self.epoch}][{i}/{len(real_images_loader)}]\tloss_d: {m["loss_d"]:.4f}\tloss_g: {m["loss_g"]:.4f}'
                                      + f'\tD(x): {m["x_d"]:.4f}\tD(G(z)): {m["z1_gd"]:.4f} / {m["z2_gd"]:.4f}')
            pending.clear()

        done = False
        self.comm.barrier()        # ranks enter the first step (and its graph capture / peer-memory exchanges) together
        for epoch in range(self.epoch):
            for i, data in enumerate(DevicePrefetcher(real_images_loader, self.device)):
                real_data = data[0]
                real_data = parallel.shard_rows(real_data, self.comm) if getattr(self.data_pre, "global_batches", False) else real_data
                scal = self.train_step(real_data.contiguous().float())
                pending.append((epoch, i, scal.clone()))       # the block lives in the step's arena: keep a copy
                if len(pending) >= 100:
                    flush()

                last = (epoch == self.epoch - 1) and (i == len(real_images_loader) - 1)
                if self.metric is not None and ((iters % 500 == 0) or last):
                    flush()
                    with torch.no_grad():
                        fake = self.model_g(fixed_noise).detach()
                    inception_score, fid = self.metric.evaluate_generated(fake)
                    self.logger.debug(f'inception score: {inception_score}\tfid: {fid}')
                    if low_fid > fid:
                        low_fid = fid
                        self.logger.debug(f"{iters} lowest fid")
                        self.save_model('fid', iters, low_fid, fake.cpu())
                    if high_is < inception_score:
                        high_is = inception_score
                        self.logger.debug(f"{iters} highest is")
                        self.save_model('is', iters, high_is, fake.cpu())
                    self.comm.barrier()    # rank 0 alone wrote checkpoints / plots: do not let the others run ahead into SyncBN
                iters += 1
                if self.max_iters and iters >= self.max_iters:
                    done = True
                    break
            if done:
                break
        flush()
        end = time.time()
        self.logger.debug(f'train finish\ttiem: {time_to_str(end - start)}')
        self.losses_g, self.losses_d = losses_g, losses_d

        if plt is not None and self.comm.rank == 0:
            plt.clf()
            epoch_x = range(1, len(losses_g) + 1)
            plt.figure(figsize=(8, 6))
            plt.plot(epoch_x, losses_d, label='Discriminator Loss')
            plt.plot(epoch_x, losses_g, label='Generator Loss')
            plt.title('Discriminator and Generator Loss')
            plt.xlabel('Iterations')
            plt.ylabel('Loss')
            plt.legend()
            plt.savefig(os.path.join(self.model_save_path, 'loss.png'))
        return losses_d, losses_g
