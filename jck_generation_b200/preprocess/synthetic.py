"""Synthetic image source with the reference preprocessors' `get_data_loader()` contract.

There is no network in the build / GPU environment, so the CIFAR-100 download the reference performs
(preprocess/dcgan_data_preprocessor.py:20-21, download=True) cannot run.  The synthetic source yields
batches shaped and scaled like the reference's transformed data: float32 [B,nc,64,64] in [-1,1)
(Resize(64) -> ToTensor -> Normalize(.5,.5), dcgan_data_preprocessor.py:38-43) and, for CGAN, int64
one-hot [B,n_classes] targets (cgan_data_preprocessor.py:11-16)."""
import torch


class SyntheticLoader:
    """Iterable of `n_batches` (images[, one-hot labels]) batches; deterministic per (seed, epoch, index).
    Batches are produced in pinned host memory so `.to(device, non_blocking=True)` overlaps."""

    def __init__(self, batch_size, n_batches, nc=3, hw=64, n_classes=None, seed=12345, pin=True, rank=0, world=1):
        """rank / world: data parallel -- `batch_size` is the GLOBAL batch, drawn identically on every rank; a rank yields its own
        rows of it (parallel.local_slice)."""
        self.batch_size, self.n_batches = batch_size, n_batches
        self.rank, self.world = rank, world
        self.nc, self.hw, self.n_classes, self.seed = nc, hw, n_classes, seed
        self.pin = pin and torch.cuda.is_available()
        self.epoch = 0
        self.targets = None

    def __len__(self):
        return self.n_batches

    def __iter__(self):
        gen = torch.Generator().manual_seed(self.seed + 7919 * self.epoch)
        self.epoch += 1
        from ..parallel import local_slice
        rows = local_slice(self.batch_size, self.rank, self.world)
        for _ in range(self.n_batches):
            x = torch.rand(self.batch_size, self.nc, self.hw, self.hw, generator=gen) * 2 - 1
            if self.world > 1:
                x = x[rows].contiguous()
            if self.pin:
                x = x.pin_memory()
            if self.n_classes is None:
                yield (x, torch.zeros(x.shape[0], dtype=torch.int64))
            else:
                idx = torch.randint(0, self.n_classes, (self.batch_size,), generator=gen)
                y = torch.nn.functional.one_hot(idx, self.n_classes).to(torch.int64)
                yield (x, y[rows].contiguous() if self.world > 1 else y)
