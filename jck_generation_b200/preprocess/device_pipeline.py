"""Device-resident input pipeline.

The reference feeds the step from `DataLoader(CIFAR100(transform=Resize(64) -> ToTensor -> Normalize), batch, shuffle=True)`
(preprocess/dcgan_data_preprocessor.py:37-49,69-75): one PIL resize + two tensor passes per SAMPLE on the host, ~10 k
images/s on a few workers -- 20x short of one B200's step rate.  Here the uint8 dataset (CIFAR-100 train: 150 MB) is
uploaded once; a batch is ONE kernel that gathers the rows of a seeded permutation, resizes them exactly as Pillow does
(fixed-point separable bilinear, uint8 after each pass), applies ToTensor / Normalize in fp32 and writes the NCHW fp32
tensor the trainers expect; CGAN's OneHotEncoder (cgan_data_preprocessor.py:11-16) is a second tiny kernel.

Results are bit-identical to the reference's transform (tests/test_input_pipeline.py, against Pillow + torchvision), and the
batch COMPOSITION is the reference's too: the index batches come from a torch DataLoader over range(N) with the same
batch size / shuffle flag, so the global torch RNG is consumed exactly as the reference's loader consumes it.
"""
import math

import numpy as np
import torch

PRECISION_BITS = 22


def bilinear_tables(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc (src/libImaging/Resample.c) for the bilinear filter over a whole
    axis, vectorised in float64: (bounds int32 [out, 2] = (first, count), coef int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = filterscale                       # bilinear support 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum(np.trunc(center - support + 0.5).astype(np.int64), 0)
    xmax = np.minimum(np.trunc(center + support + 0.5).astype(np.int64), in_size)
    count = xmax - xmin
    k = np.arange(ksize, dtype=np.float64)[None, :]
    t = np.abs((k + xmin[:, None] - center[:, None] + 0.5) * (1.0 / filterscale))
    w = np.where(t < 1.0, 1.0 - t, 0.0)
    w = np.where(np.arange(ksize)[None, :] < count[:, None], w, 0.0)
    ww = np.zeros(out_size, dtype=np.float64)
    for j in range(ksize):                      # left-to-right accumulation, as the C loop
        ww = ww + w[:, j]
    w = np.where(ww[:, None] != 0.0, w / np.where(ww[:, None] != 0.0, ww[:, None], 1.0), w)
    coef = np.trunc(0.5 + w * float(1 << PRECISION_BITS)).astype(np.int32)
    bounds = np.stack([xmin, count], axis=1).astype(np.int32)
    return bounds, coef


class _Dataset:
    """what `len(loader.dataset)` / `dataset.targets` see (metrics.py:56-66,97)"""

    def __init__(self, n, targets):
        self.n, self.targets = n, targets

    def __len__(self):
        return self.n


class _Indices(torch.utils.data.Dataset):
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return i


class DeviceImageLoader:
    """Iterable of (images fp32 [B, C, size_h, size_w] on the device, labels) with the reference loader's contract.
    labels: int64 one-hot [B, n_classes] when `n_classes` is given (CGAN), else int64 class indices [B]."""

    def __init__(self, data_u8, targets, batch_size, size, mean, std, shuffle=True, n_classes=None, device="cuda", rank=0, world=1):
        """rank / world: data parallel -- `batch_size` is the GLOBAL batch; every rank draws the same index batch (same torch
        seed on all ranks, main.py:34) and gathers only its own rows (parallel.local_slice), so no row is processed twice."""
        from .. import _lib
        self._lib = _lib
        self.rank, self.world = rank, world
        data_u8 = np.ascontiguousarray(data_u8)
        assert data_u8.dtype == np.uint8 and data_u8.ndim == 4, "dataset must be uint8 [N, H, W, C]"
        self.N, self.Hi, self.Wi, self.C = data_u8.shape
        self.Ho, self.Wo = (size, size) if isinstance(size, int) else size
        if isinstance(size, int):               # tt.Resize(int): the SMALLER edge becomes `size`, aspect kept
            if self.Hi <= self.Wi:
                self.Ho, self.Wo = size, int(size * self.Wi / self.Hi)
            else:
                self.Ho, self.Wo = int(size * self.Hi / self.Wi), size
        self.device = torch.device(device)
        self.batch_size, self.shuffle, self.n_classes = batch_size, shuffle, n_classes
        self.data = torch.from_numpy(data_u8).to(self.device)
        targets = list(targets) if targets is not None else [0] * self.N
        self.labels = torch.as_tensor(targets, dtype=torch.int64, device=self.device)
        self.dataset = _Dataset(self.N, targets)
        self.targets = targets
        hb, hk = bilinear_tables(self.Wi, self.Wo)
        vb, vk = bilinear_tables(self.Hi, self.Ho)
        self._hb, self._hk = torch.from_numpy(hb).to(self.device), torch.from_numpy(hk).to(self.device)
        self._vb, self._vk = torch.from_numpy(vb).to(self.device), torch.from_numpy(vk).to(self.device)
        self._mean = [float(m) for m in mean]
        self._std = [float(s) for s in std]
        self._index_loader = torch.utils.data.DataLoader(_Indices(self.N), batch_size, shuffle=shuffle, num_workers=0)

    def __len__(self):
        return len(self._index_loader)

    def batch(self, index):
        """images / labels of the dataset rows `index` (int64 tensor)"""
        import ctypes
        lib, check = self._lib.load(), self._lib.check
        index = index.to(self.device, dtype=torch.int64, non_blocking=True).contiguous()
        B = index.numel()
        out = torch.empty(B, self.C, self.Ho, self.Wo, dtype=torch.float32, device=self.device)
        st = torch.cuda.current_stream().cuda_stream
        mean = (ctypes.c_float * self.C)(*self._mean)
        std = (ctypes.c_float * self.C)(*self._std)
        check(lib.jck_u8_resize_norm(self.data.data_ptr(), index.data_ptr(), out.data_ptr(), B, self.Hi, self.Wi, self.C, self.Ho,
                                     self.Wo, self._hb.data_ptr(), self._hk.data_ptr(), self._hk.shape[1], self._vb.data_ptr(),
                                     self._vk.data_ptr(), self._vk.shape[1], ctypes.cast(mean, ctypes.c_void_p),
                                     ctypes.cast(std, ctypes.c_void_p), st), "u8_resize_norm")
        if self.n_classes is None:
            return out, self.labels[index]
        onehot = torch.empty(B, self.n_classes, dtype=torch.int64, device=self.device)
        check(lib.jck_one_hot_i64(self.labels.data_ptr(), index.data_ptr(), onehot.data_ptr(), B, self.n_classes, st), "one_hot")
        return out, onehot

    def __iter__(self):
        from ..parallel import local_slice
        for index in self._index_loader:
            if self.world > 1:
                index = index[local_slice(index.numel(), self.rank, self.world)]
                if index.numel() == 0:
                    continue
            yield self.batch(index)
