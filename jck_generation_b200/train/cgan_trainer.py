"""Drop-in for the reference's train/cgan_trainer.py: `CGANTrainer(args, model_g, model_d, data_pre)` with
`.train()`, `.compute_gradient_penalty(real, fake, labels)`, `.save_model(typ, iters, inception_score, fid,
intra_fid, images)`, `.save_image(path, iters, images)`.

The loop keeps the reference's structure (cgan_trainer.py:134-270): labels threaded through G and D, one
combined `error_d.backward()` that includes 10*GP (:200-204), fixed evaluation set of 100 classes x 10 noise
vectors (:144-153), IS / FID / intra-FID every 500 iterations when a Metrics source exists.  The step body
(:173-213) is `CGANStep` -- our kernels, including the explicit second-order sweep for the penalty."""
import argparse
import os
import time

import torch
import torch.nn as nn

from .. import ops, parallel
from ..logger.main_logger import MainLogger
from ..logger.utils import time_to_str
from ..model.CGAN import weights_init
from ..utils import get_default_device
from .cgan_step import CGANStep
from .dcgan_step import DCGANStep
from .dcgan_trainer import _dtype_of
from .optim import FusedAdam
from .prefetch import DevicePrefetcher
from .trainer import Trainer

try:
    import matplotlib.pyplot as plt
except Exception:  # pragma: no cover
    plt = None


class CGANTrainer(Trainer):
    def __init__(self, args: argparse.Namespace, model_g: nn.Module, model_d: nn.Module, data_pre):
        self.logger = MainLogger(args)
        self.device = get_default_device()
        if self.device.type != "cuda":
            raise RuntimeError("CGANTrainer: no CUDA device; the B200 train step has no CPU fallback")
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
        if self.comm.world_size > 1:
            for t in list(self.model_g.state_dict().values()) + list(self.model_d.state_dict().values()):
                torch.distributed.broadcast(t, src=0)
        self.data_pre = data_pre
        self.train_loader, metric_source = self.data_pre.get_data_loader()
        self.metric = None
        if metric_source is not None and getattr(args, "metrics", 1):
            from ..metrics import Metrics
            self.metric = Metrics(metric_source)
        self.flat_g = parallel.FlatParams(self.model_g)
        self.flat_d = parallel.FlatParams(self.model_d)
        self.optimizer_g = FusedAdam(self.model_g.parameters(), lr=self.max_lr, betas=[0.5, 0.999], flat=self.flat_g)
        self.optimizer_d = FusedAdam(self.model_d.parameters(), lr=self.max_lr, betas=[0.5, 0.999], flat=self.flat_d)
        self.criterion = nn.BCELoss()
        self.step = CGANStep(self.model_g, self.model_d, self.optimizer_g, self.optimizer_d, self.flat_g, self.flat_d,
                             self.comm, self.lambda_gp, seed=int(getattr(args, "seed", 12345)))
        self.use_graph = bool(getattr(args, "cuda_graph", 0)) and (
            self.comm.world_size == 1 or bool(int(os.environ.get("JCK_DP_GRAPH", "1"))))
        self.max_iters = int(getattr(args, "max_iters", 0))
        self.model_save_path = args.save_path
        if self.comm.rank == 0:
            os.makedirs(self.model_save_path, exist_ok=True)
        self.logger.debug(f'save path: {self.model_save_path}')

    def save_model(self, typ, iters, inception_score, fid, intra_fid, images):
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
        }, os.path.join(save_path, f'{iters}_{inception_score:.04f}_{fid:.04f}_{intra_fid:.04f}.pt'))
        self.save_image(save_path, iters, images)

    def save_image(self, path, iters, images):
        if plt is None or self.comm.rank != 0:
            return
        import numpy as np
        plt.clf()
        fig = plt.figure(figsize=(10, 10))
        for i in range(min(100, len(images))):
            fig.add_subplot(10, 10, i + 1)
            plt.title(self.data_pre.idx_to_labels[i])
            plt.axis('off')
            plt.imshow(np.transpose(images[i].cpu().numpy(), (1, 2, 0)))
        plt.savefig(os.path.join(path, f'{iters}_fake_image.png'))
        plt.close()

    def _gp_counter(self):
        if getattr(self, "_gp_rng_counter", None) is None:
            self._gp_rng_counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        return self._gp_rng_counter

    def compute_gradient_penalty(self, real_data, fake_data, labels_data, alpha=None, dropout_mask=None):
        """mean((||d D(x_hat, y)/d x_hat||_2 - 1)^2) (reference :114-131) as a plain tensor.  Inside `train()`
        the penalty's gradient comes from CGANStep's second-order sweep, not from this method."""
        ed = self.model_d.engine()
        B = real_data.size(0)
        if alpha is None:
            # a draw of this public method must neither collide with the step's Philox streams (ids 16*rank + 1..5) nor move
            # the training counter: private counter, stream ids from a disjoint range
            alpha = torch.empty(B, 1, 1, 1, device=self.device)
            ops.rand(alpha, self.step.seed, (1 << 20) + 1, self._gp_counter())
            ops.rng_advance(self._gp_counter(), (B + 3) // 4)
        if dropout_mask is None:
            dropout_mask = torch.empty(B, 256, dtype=torch.float32, device=self.device)
            ops.dropout_mask(dropout_mask, 0.25, self.step.seed, (1 << 20) + 2, self._gp_counter())
        x_hat = ops.img_alloc(B, ed.nc, 64, 64, ed.dtype, self.device, ed.img_layout)
        ops.prep_image(real_data.detach().contiguous().float(), out_nhwc=x_hat, a1=1.0,
                       x2=fake_data.detach().contiguous().float(), alpha=alpha.reshape(B).contiguous(), layout=ed.img_layout)
        ctx = ed.trunk_forward(x_hat, groups=1)
        ed.head_forward(ctx, labels_data, dropout_mask)
        v = ed.trunk_backward(ctx, ed.head_gp_seed(ctx), wgrad=False, input_grad=True)
        out = torch.zeros(2, dtype=torch.float32, device=self.device)
        ops.gp_seed(v, None, out, 0.0)
        return out[0]

    def train_step(self, real_data, labels_data, rng=None):
        """One G+D step on this rank's rows; returns the [4,2] device scalar block.  With args.cuda_graph the step is
        captured once per (batch size, label dtype) and replayed."""
        if self.use_graph and rng is None:
            st = self.step
            if (st._graph is None or st._static.shape[0] != real_data.shape[0] or st._static_labels.dtype != labels_data.dtype):
                st.capture(real_data.shape[0], label_dtype=labels_data.dtype)
            return st.replay(real_data, labels_data)
        return self.step.run(real_data, labels_data, rng)

    def train(self):
        real_images_loader = self.train_loader
        losses_g, losses_d = [], []
        iters = 0
        n_cls = self.model_g.n_classes
        fixed_noise = torch.empty(n_cls * 10, self.model_g.nz, 1, 1, device=self.device)          # :144-153
        ops.randn(fixed_noise, self.step.seed, 7, None)
        fixed_labels = torch.eye(n_cls, device=self.device).repeat_interleave(10, dim=0)
        low_fid = low_intra_fid = 1e10
        high_is = 0
        start = time.time()
        self.logger.debug("train start")
        pending = []

        def flush():
            if not pending:
                return
            self.comm.check_health()
            block = torch.stack([p[2] for p in pending])
            self.comm.allreduce_mean_(block)
            for (ep, i, _), s in zip(pending, block.cpu()):
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
                real_data, labels_data = data
                if getattr(self.data_pre, "global_batches", False):      # host DataLoader under torchrun: this rank's rows
                    real_data, labels_data = parallel.shard_rows(real_data, self.comm), parallel.shard_rows(labels_data, self.comm)
                real_data = real_data.contiguous().float()
                scal = self.train_step(real_data, labels_data)
                pending.append((epoch, i, scal.clone()))       # the block lives in the step's arena: keep a copy
                if len(pending) >= 100:
                    flush()
                last = (epoch == self.epoch - 1) and (i == len(real_images_loader) - 1)
                if self.metric is not None and ((iters % 500 == 0) or last):
                    flush()
                    with torch.no_grad():
                        fake = self.model_g(fixed_noise, fixed_labels).detach()
                    inception_score, fid, intra_fid = self.metric.evaluate_generated(fake, intra=True)
                    self.logger.debug(f'inception score: {inception_score}\tfid: {fid}\tintra fid: {intra_fid}')
                    denorm = (0.5 * fake + 0.5)[::10]
                    if low_fid > fid:
                        low_fid = fid
                        self.save_model('fid', iters, inception_score, fid, intra_fid, denorm)
                    if low_intra_fid > intra_fid:            # cgan_trainer.py:242-245 (never true for a NaN intra-FID)
                        low_intra_fid = intra_fid
                        self.save_model('intra_fid', iters, inception_score, fid, intra_fid, denorm)
                    if high_is < inception_score:
                        high_is = inception_score
                        self.save_model('is', iters, inception_score, fid, intra_fid, denorm)
                    self.comm.barrier()    # rank 0 alone wrote checkpoints / plots: do not let the others run ahead into SyncBN
                iters += 1
                if self.max_iters and iters >= self.max_iters:
                    done = True
                    break
            if done:
                break
        flush()
        self.logger.debug(f'train finish\ttiem: {time_to_str(time.time() - start)}')
        self.losses_g, self.losses_d = losses_g, losses_d
        return losses_d, losses_g
