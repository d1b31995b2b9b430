"""Interface-compatible stand-in for the reference's metrics.py (secondary path).

Same public surface -- `Metrics(real_images)`, `.inception_score(loader, splits=10)`,
`.fid(loader, intra_fid=False, label=0)`, `.intra_fid(tensor)` -- and the same definitions
(metrics.py:80-141): features are the 100-d logits of an Inception-v3 whose fc is Linear(2048,100)
(:46-52, :87); IS = exp(mean KL(p(y|x) || p(y))) per split; FID = |mu1-mu2|^2 + tr(S1+S2-2 sqrtm(S1 S2));
intra-FID sums the 20 CIFAR-100 superclass FIDs and divides by 100 (sic, :141 -- kept for parity).

Status in this round: PARTLY on our kernels.  The feature moments of the generated set (np.mean / np.cov of the
reference, metrics.py:118-124) are taken on the device by ops.feature_moments -- column sums, a centred hi/lo bf16
split and the Gram matrix as three tcgen05 GEMMs with fp32 accumulation (jck_gemm_tc, both operands MN-major),
all-reducible across ranks -- and only the d x d result goes to the host for scipy's sqrtm.  The Inception forward
is still torchvision's (library code) and the IS reduction numpy.  Differences forced by the environment:
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


def _entropy(pk, qk):
    """scipy.stats.entropy(pk, qk): KL divergence after normalising both."""
    pk = pk / pk.sum()
    qk = qk / qk.sum()
    mask = pk > 0
    return float(np.sum(pk[mask] * np.log(pk[mask] / qk[mask])))


class Metrics:
    def __init__(self, real_images=None, feature="logits", checkpoint=os.path.join('./save/iception_v3', 'loss_bset.pt'),
                 cache=os.path.join('./data', 'metric_data.pikl')):
        from torchvision import models
        self.device = get_default_device()
        self.feature = feature
        self.class_to_superclass = {c: s for s, cs in enumerate(SUPERCLASS) for c in cs}
        torch.manual_seed(12345)
        self.inception_model = models.inception_v3(weights=None, aux_logits=True, init_weights=False)
        self.inception_model.aux_logits = False
        self.inception_model.fc = nn.Sequential(nn.Linear(self.inception_model.fc.in_features, 100))
        if os.path.exists(checkpoint):
            self.inception_model.load_state_dict(torch.load(checkpoint, map_location="cpu"))
        if feature == "pool3":
            self.inception_model.fc = nn.Identity()
        self.inception_model.to(self.device).eval()

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
            loader = real_images if isinstance(real_images, torch.utils.data.DataLoader) else \
                torch.utils.data.DataLoader(real_images, 128, shuffle=False, num_workers=0, pin_memory=True)
            self.real_features = self._extract(loader, real=True)

    @torch.no_grad()
    def _extract(self, images, real=False, softmax=False, on_device=False):
        feats = []
        for image in images:
            if real or isinstance(image, (list, tuple)):
                image = image[0]
            out = self.inception_model(image.to(self.device, non_blocking=True).float())
            feats.append(nn.functional.softmax(out, dim=1) if softmax else out)
        feats = torch.cat(feats)
        return feats.float().contiguous() if on_device else feats.double().cpu().numpy()

    def _moments(self, feats):
        """(mean, covariance) as float64 numpy: on our kernels for device-resident features, numpy otherwise."""
        if torch.is_tensor(feats) and feats.is_cuda and feats.shape[0] > 1:
            from . import ops
            mean, cov = ops.feature_moments(feats)
            return mean.double().cpu().numpy(), cov.double().cpu().numpy()
        feats = feats.double().cpu().numpy() if torch.is_tensor(feats) else feats
        return np.mean(feats, axis=0), np.cov(feats, rowvar=False)

    def inception_score(self, images, splits=10):
        n = len(images.dataset)
        preds = self._extract(images, softmax=True)
        split_scores = []
        for k in range(splits):
            part = preds[k * (n // splits): (k + 1) * (n // splits), :]
            if part.shape[0] == 0:
                continue
            py = np.mean(part, axis=0)
            split_scores.append(np.exp(np.mean([_entropy(part[i, :], py) for i in range(part.shape[0])])))
        return float(np.mean(split_scores))

    def fid(self, generated_images, intra_fid=False, label=0):
        from scipy.linalg import sqrtm
        generated_features = self._extract(generated_images, on_device=torch.cuda.is_available())
        real = self.real_features
        if real is None:
            raise RuntimeError("Metrics.fid: no real-image features (construct Metrics with a dataset or loader)")
        if intra_fid:
            real = real[self.real_superclass_idx[label]]
        mu1, sigma1 = self._moments(real)
        mu2, sigma2 = self._moments(generated_features)
        diff = np.sum((mu1 - mu2) ** 2.0)
        covmean = sqrtm(sigma1.dot(sigma2))
        if np.iscomplexobj(covmean):
            covmean = covmean.real
        return float(diff + np.trace(sigma1 + sigma2 - 2.0 * covmean))

    def intra_fid(self, generated_images):
        total = 0.0
        for sidx in range(20):
            loader = torch.utils.data.DataLoader(generated_images[self.fake_superclass_idx[sidx]], 128,
                                                 pin_memory=False, num_workers=0, shuffle=False)
            total += self.fid(loader, intra_fid=True, label=sidx)
        return total / 100

    def evaluate_generated(self, fake):
        """The reference's eval branch (dcgan_trainer.py:203-211) without the GPU->CPU->GPU round trip:
        de-normalise, resize to 299, ImageNet-normalise on the device, then IS and FID."""
        x = 0.5 * fake.float() + 0.5
        x = nn.functional.interpolate(x, size=(299, 299), mode="bilinear", align_corners=False, antialias=False)
        mean = torch.tensor([0.485, 0.456, 0.406], device=x.device).view(1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225], device=x.device).view(1, 3, 1, 1)
        x = (x - mean) / std
        loader = torch.utils.data.DataLoader(x, batch_size=64)
        score = self.inception_score(loader)
        fid = self.fid(loader) if self.real_features is not None else float("nan")
        return score, fid
