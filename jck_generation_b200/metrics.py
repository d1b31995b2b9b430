"""The reference's metrics.py (secondary path) on the jck kernels.

Same public surface -- `Metrics(real_images)`, `.inception_score(loader, splits=10)`,
`.fid(loader, intra_fid=False, label=0)`, `.intra_fid(tensor)` -- and the same definitions
(metrics.py:80-141): features are the 100-d logits of an Inception-v3 whose fc is Linear(2048,100)
(:46-52, :87); IS = exp(mean KL(p(y|x) || p(y))) per split; FID = |mu1-mu2|^2 + tr(S1+S2-2 sqrtm(S1 S2));
intra-FID sums the 20 CIFAR-100 superclass FIDs and divides by 100 (sic, :141 -- kept for parity).

What runs where:
  * the Inception-v3 forward (`self.inception_model(image)`, :87): `inception.InceptionV3` -- 94 tcgen05 implicit-GEMM
    launches + pooling kernels in one CUDA graph; the torchvision module built here is only the parameter container the
    reference's checkpoint loads into (same state_dict keys), its forward is never called;
  * the eval branch's pre-processing (dcgan_trainer.py:203-207): fused into the stem kernel (`evaluate_generated`);
  * softmax + split marginals + KL (:96-110): `jck_inception_score`;
  * np.mean / np.cov (:118-124): `ops.feature_moments` (tcgen05 Gram matrix, all-reducible across ranks);
  * the d x d matrix square root stays scipy's `sqrtm` on the host, as in the reference (:128).
There is no CPU path: without a CUDA device the constructor raises.  Differences forced by the environment:
  * ./save/iception_v3/loss_bset.pt (:51; spelling is the on-disk contract) is loaded when present,
    otherwise the network is seeded random-init -- there is no network to fetch weights;
  * `real_images` may be a dataset with `.targets` (the CGAN preprocessor, as the reference expects), a
    DataLoader (what the reference's DCGAN preprocessor actually passes, which crashes the reference at
    :56) or None; `feature='pool3'` gives the 2048-d features BASELINE configs[4] names.
"""
import os
import pickle

import numpy as np
import torch
import torch.nn as nn

from .utils import get_default_device

SUPERCLASS = [
    [4, 30, 55, 72, 95], [1, 32, 67, 73, 91], [54, 62, 70, 82, 92], [9, 10, 16, 28, 61], [0, 51, 53, 57, 83],
    [22, 39, 40, 86, 87], [5, 20, 25, 84, 94], [6, 7, 14, 18, 24], [3, 42, 43, 88, 97], [12, 17, 37, 68, 76],
    [23, 33, 49, 60, 71], [15, 19, 21, 31, 38], [34, 63, 64, 66, 75], [26, 45, 77, 79, 99], [2, 11, 35, 46, 98],
    [27, 29, 44, 78, 93], [36, 50, 65, 74, 80], [47, 52, 56, 59, 96], [8, 13, 48, 58, 90], [41, 69, 81, 85, 89]]


class Metrics:
    def __init__(self, real_images=None, feature="logits", checkpoint=os.path.join('./save/iception_v3', 'loss_bset.pt'),
                 cache=os.path.join('./data', 'metric_data.pikl'), batch=128, comm=None, allow_random_weights=None,
                 precision=None):
        from torchvision import models
        from .inception import InceptionV3
        if not torch.cuda.is_available():
            raise RuntimeError("Metrics: the Inception-v3 forward runs on the jck sm_100a kernels; there is no CPU path")
        self.device = get_default_device()
        self.feature, self.batch = feature, batch
        self.comm = comm                      # data-parallel evaluation (BASELINE configs[4]): see evaluate_generated
        self.class_to_superclass = {c: s for s, cs in enumerate(SUPERCLASS) for c in cs}
        # parameter container only (the reference's checkpoint format); the forward below is ours.  Built under a forked RNG:
        # constructing the container must not disturb the global torch RNG the loaders' shuffle order comes from.
        with torch.random.fork_rng(devices=[]):
            torch.manual_seed(12345)
            self.inception_model = models.inception_v3(weights=None, aux_logits=True, init_weights=False)
            self.inception_model.aux_logits = False
            self.inception_model.fc = nn.Sequential(nn.Linear(self.inception_model.fc.in_features, 100))
        if os.path.exists(checkpoint):
            self.inception_model.load_state_dict(torch.load(checkpoint, map_location="cpu"))
        else:
            # the reference raises here (metrics.py:51 torch.load of a missing file).  Scores from a random-weight network
            # are meaningless for model selection, so that is opt-in only (throughput benchmarks, plumbing tests).
            if allow_random_weights is None:
                allow_random_weights = os.environ.get("JCK_METRICS_RANDOM_WEIGHTS", "0") == "1"
            if not allow_random_weights:
                raise FileNotFoundError(f"Metrics: Inception-v3 checkpoint {checkpoint} not found (the reference's metrics.py:51 "
                                        "loads it unconditionally); pass allow_random_weights=True / JCK_METRICS_RANDOM_WEIGHTS=1 "
                                        "to run the evaluation plumbing on seeded random weights")
            import warnings
            warnings.warn(f"Metrics: {checkpoint} missing -- IS / FID below come from a RANDOM-WEIGHT Inception-v3 and are "
                          "meaningless as quality scores")
        self.inception_model.eval()
        # "split" (default): activations / weights as hi + lo bf16 pairs, fp32-grade features -- the reference runs the network
        # in fp32 (metrics.py:87), and this is the mode pinned against torchvision fp32 free running (tests/test_gpu_incep.py:
        # <= 2e-3 at logits / pool3 of a random-weight network, where plain bf16 is 15 % off).  "bf16" (JCK_INCEPTION_PRECISION
        # =bf16): one third of the tensor work, for throughput runs.
        if precision is None:
            precision = os.environ.get("JCK_INCEPTION_PRECISION", "split")
        self.precision = precision
        self.extractor = InceptionV3(self.inception_model.state_dict(), feature=feature, device=self.device, precision=precision)

        real_targets = getattr(real_images, "targets", None)
        fake_targets = [i for i in range(100) for _ in range(10)]
        self.real_superclass_idx, self.fake_superclass_idx = {}, {}
        for sidx in range(20):
            if real_targets is not None:
                self.real_superclass_idx[sidx] = [i for i, t in enumerate(real_targets) if self.class_to_superclass[int(t)] == sidx]
            self.fake_superclass_idx[sidx] = [i for i, t in enumerate(fake_targets) if self.class_to_superclass[t] == sidx]

        self.real_features = None
        if os.path.exists(cache) and feature == "logits":
            with open(cache, 'rb') as f:
                self.real_features = pickle.load(f)
        elif real_images is not None:
            is_loader = isinstance(real_images, torch.utils.data.DataLoader) or not isinstance(real_images, torch.utils.data.Dataset) \
                and hasattr(real_images, "__iter__") and hasattr(real_images, "batch_size")    # e.g. preprocess.DeviceImageLoader
            loader = real_images if is_loader else \
                torch.utils.data.DataLoader(real_images, batch, shuffle=False, num_workers=0, pin_memory=True)
            self.real_features = self._extract(loader, real=True).cpu().numpy()
            if feature == "logits" and (comm is None or getattr(comm, "rank", 0) == 0):
                # the reference pickles the real-image features after the first extraction (metrics.py:70-77)
                try:
                    os.makedirs(os.path.dirname(cache) or ".", exist_ok=True)
                    with open(cache, 'wb') as f:
                        pickle.dump(self.real_features, f)
                except OSError:
                    pass

    @torch.no_grad()
    def _extract(self, images, real=False, generated=False):
        """features [n, d] fp32 on the device (metrics.py:80-93 without the per-batch .cpu().numpy())"""
        feats = []
        for image in images:
            if real or isinstance(image, (list, tuple)):
                image = image[0]
            image = image.to(self.device, non_blocking=True).float()
            feats.append(self.extractor.forward_generated(image) if generated else self.extractor.forward(image))
        return torch.cat(feats).float().contiguous()

    def _moments(self, feats, sharded=False):
        """(mean, covariance) as float64 numpy, taken on the device (ops.feature_moments).  `sharded`: `feats` holds this
        rank's rows only; the column sums and the Gram matrix are all-reduced, every rank gets the moments of ALL rows."""
        from . import ops
        if not torch.is_tensor(feats):
            feats = torch.as_tensor(np.ascontiguousarray(feats), dtype=torch.float32)
        feats = feats.to(self.device).float().contiguous()
        mean, cov = ops.feature_moments(feats, self.comm if sharded else None)
        return mean.double().cpu().numpy(), cov.double().cpu().numpy()

    def _score(self, logits, n, splits):
        from . import ops
        per = n // splits
        if per == 0:
            return float("nan")
        scores = torch.zeros(splits, dtype=torch.float32, device=self.device)
        ops.inception_score(logits[:per * splits].contiguous(), splits, scores)
        return float(scores.mean().item())

    def inception_score(self, images, splits=10):
        n = len(images.dataset)
        return self._score(self._extract(images), n, splits)

    def _fid_from(self, generated_features, intra_fid=False, label=0, sharded=False):
        from scipy.linalg import sqrtm
        real = self.real_features
        if real is None:
            raise RuntimeError("Metrics.fid: no real-image features (construct Metrics with a dataset or loader)")
        if intra_fid:
            real = real[self.real_superclass_idx[label]]
        mu1, sigma1 = self._moments(real)
        mu2, sigma2 = self._moments(generated_features, sharded)
        diff = np.sum((mu1 - mu2) ** 2.0)
        covmean = sqrtm(sigma1.dot(sigma2))
        if np.iscomplexobj(covmean):
            covmean = covmean.real
        return float(diff + np.trace(sigma1 + sigma2 - 2.0 * covmean))

    def fid(self, generated_images, intra_fid=False, label=0):
        return self._fid_from(self._extract(generated_images), intra_fid, label)

    def intra_fid(self, generated_images):
        total = 0.0
        for sidx in range(20):
            loader = torch.utils.data.DataLoader(generated_images[self.fake_superclass_idx[sidx]], 128,
                                                 pin_memory=False, num_workers=0, shuffle=False)
            total += self.fid(loader, intra_fid=True, label=sidx)
        return total / 100

    def evaluate_generated(self, fake, intra=False):
        """The reference's eval branch (dcgan_trainer.py:198-211, cgan_trainer.py:221-235) on the device: `fake` is the
        generator's output in [-1, 1]; de-normalise / resize to 299 / ImageNet-normalise are fused into the Inception stem
        kernel, and the features are extracted ONCE and feed the score, the FID and -- with `intra` (CGAN: `fake` holds the
        1000 class-ordered samples of cgan_trainer.py:146-153) -- the intra-FID; the reference runs the network 22 times.
        Returns (score, fid) or (score, fid, intra_fid)."""
        fake = fake.to(self.device).float()
        feats = self._extract(fake.split(self.batch), generated=True)
        score = self._score(feats, feats.shape[0], 10) if self.feature == "logits" else float("nan")
        fid = self._fid_from(feats) if self.real_features is not None else float("nan")
        if not intra:
            return score, fid
        return score, fid, self._intra_fid_from(feats)

    def _intra_fid_from(self, feats):
        """metrics.py:133-141 on already extracted features: the FID of every CIFAR-100 superclass's generated rows against
        that superclass's real rows, summed and divided by 100 (sic)."""
        if self.real_features is None or not self.real_superclass_idx:
            return float("nan")
        total = 0.0
        for sidx in range(20):
            rows = self.fake_superclass_idx[sidx]
            if not rows or max(rows) >= feats.shape[0] or not self.real_superclass_idx.get(sidx):
                return float("nan")           # not the 1000-sample class-ordered set / a superclass without real rows
            total += self._fid_from(feats[rows], intra_fid=True, label=sidx)
        return total / 100

    def evaluate_generated_sharded(self, fake_local):
        """Data-parallel evaluation (BASELINE configs[4]: 50 k generated samples over 8 GPUs): every rank passes ITS rows of
        the generated set (rank r = rows [r n/W, (r+1) n/W), n/W equal on all ranks); features are extracted locally, their
        sums and Gram matrix all-reduced (2 x d + d x d floats), the d-wide logits all-gathered for the split scores.  Every
        rank returns the same (score, fid) as one rank evaluating all n rows."""
        import torch.distributed as dist
        comm = self.comm
        assert comm is not None and comm.world_size > 1, "construct Metrics(comm=...) under torchrun"
        feats = self._extract(fake_local.to(self.device).float().split(self.batch), generated=True)
        score = float("nan")
        if self.feature == "logits":
            parts = [torch.empty_like(feats) for _ in range(comm.world_size)]
            dist.all_gather(parts, feats)
            allf = torch.cat(parts)
            score = self._score(allf, allf.shape[0], 10)
        fid = self._fid_from(feats, sharded=True) if self.real_features is not None else float("nan")
        return score, fid
