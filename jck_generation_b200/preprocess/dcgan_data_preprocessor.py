"""Drop-in for the reference's preprocess/dcgan_data_preprocessor.py: `DCGANDataPreprocessor(args)`,
`.transform_data()`, `.get_data_loader() -> (train_loader, inception_loader)`.

The behaviour is the reference's (torchvision CIFAR-100 under ./data, downloaded when missing,
Resize(64)/ToTensor/Normalize(.5,.5) for training, Resize(299)/ImageNet-normalise for the metric
loader, shuffle=True, pin_memory=True; dcgan_data_preprocessor.py:20-21,37-75); a missing dataset that
cannot be downloaded raises.  Only with args.synthetic=1 (an explicit request: benchmarks, smoke runs) a
synthetic source with the same contract is used (preprocess/synthetic.py)."""
import os

import torch

from ..logger.main_logger import MainLogger
from .synthetic import SyntheticLoader

IMAGENET_MEAN, IMAGENET_STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def synthetic_u8(n, n_classes=100, hw=32, seed=12345):
    """CIFAR-shaped stand-in: uint8 [n, hw, hw, 3] + class indices (no network here to download the real set)"""
    import numpy as np
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n, hw, hw, 3), dtype=np.uint8), rng.integers(0, n_classes, n).tolist()


def u8_source(pre, args, n_classes=100):
    """(uint8 [N,H,W,3], targets) for the device pipeline, or None: CIFAR-100 when it is on disk, a synthetic uint8 set with
    args.synthetic_u8=1; None (-> the host loaders) without a CUDA device or with args.device_pipeline=0."""
    if not torch.cuda.is_available() or not int(getattr(args, "device_pipeline", 1)):
        return None
    if int(getattr(args, "synthetic_u8", 0)):
        return synthetic_u8(pre.synthetic_batches * pre.batch_size, n_classes)
    if not pre.synthetic:
        return pre._trainset.data, pre._trainset.targets
    return None


def _cifar_available(root="./data"):
    return os.path.isdir(os.path.join(root, "cifar-100-python"))


def _cifar100(torchvision, root="./data"):
    """torchvision.datasets.CIFAR100(root, train=True, download=True) as the reference builds it
    (preprocess/dcgan_data_preprocessor.py:20-21); a failed download is an error that names the way out."""
    try:
        return torchvision.datasets.CIFAR100(root, train=True, download=not _cifar_available(root), transform=None)
    except Exception as e:      # noqa: BLE001 -- no network, bad archive ...
        raise RuntimeError(f"CIFAR-100 is not under {root}/cifar-100-python and could not be downloaded ({e}); put the dataset "
                           "there, or pass --synthetic 1 to train on synthetic images (benchmarks / smoke runs only)") from e


class DCGANDataPreprocessor:
    def __init__(self, args):
        self._logger = MainLogger(args)
        self.batch_size = args.batch_size
        self.num_worker = getattr(args, "num_worker", 0)
        # synthetic images ONLY on request (args.synthetic / --synthetic); otherwise the reference's behaviour: CIFAR-100 from
        # ./data, downloaded when missing (download=True, :20-21), and an error when that fails -- never a silent stand-in
        self.synthetic = bool(getattr(args, "synthetic", 0))
        self.synthetic_batches = int(getattr(args, "synthetic_batches", 391))   # 50000 / 128
        self._trainset = self._inceptionset = None
        if not self.synthetic:
            import torchvision
            self._trainset = _cifar100(torchvision)
            self._inceptionset = _cifar100(torchvision)
        self._u8 = u8_source(self, args)
        self._logger.debug('data preprocessor init' + (' (synthetic source)' if self.synthetic else '') +
                           (' (device pipeline)' if self._u8 is not None else ''))

    def transform_data(self):
        if self.synthetic or self._u8 is not None:
            return
        import torchvision.transforms as tt
        self._trainset.transform = tt.Compose([
            tt.Resize(64), tt.ToTensor(),
            tt.Normalize(mean=[0.5, 0.5, 0.5], std=[0.5, 0.5, 0.5], inplace=True)])
        self._inceptionset.transform = tt.Compose([
            tt.Resize((299, 299)), tt.ToTensor(),
            tt.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        self._logger.debug('data transform')

    def get_data_loader(self):
        # data parallel (torchrun): args.batch_size is the GLOBAL batch.  The device and synthetic loaders hand every rank its
        # own rows; the host DataLoader yields global batches and the trainer takes the rank's rows (`global_batches`).
        from ..parallel import env_rank_world
        rank, world = env_rank_world()
        self.global_batches = self._u8 is None and not self.synthetic
        if self._u8 is not None:
            # the reference's two transforms (:37-49) and loaders (:69-75) on the device, bit-identical results
            from .device_pipeline import DeviceImageLoader
            data, targets = self._u8
            self.trainloader = DeviceImageLoader(data, targets, self.batch_size, 64, [0.5] * 3, [0.5] * 3, shuffle=True,
                                                 rank=rank, world=world)
            self.inceptionloader = DeviceImageLoader(data, targets, self.batch_size * 2, (299, 299), IMAGENET_MEAN, IMAGENET_STD,
                                                     shuffle=False)
        elif self.synthetic:
            self.trainloader = SyntheticLoader(self.batch_size, self.synthetic_batches, rank=rank, world=world)
            self.inceptionloader = None
        else:
            self.trainloader = torch.utils.data.DataLoader(self._trainset, self.batch_size, shuffle=True,
                                                           num_workers=self.num_worker, pin_memory=True)
            self.inceptionloader = torch.utils.data.DataLoader(self._inceptionset, self.batch_size * 2,
                                                               pin_memory=True, num_workers=0)
        return self.trainloader, self.inceptionloader
