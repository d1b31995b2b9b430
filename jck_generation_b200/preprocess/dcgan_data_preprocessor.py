"""Drop-in for the reference's preprocess/dcgan_data_preprocessor.py: `DCGANDataPreprocessor(args)`,
`.transform_data()`, `.get_data_loader() -> (train_loader, inception_loader)`.

With CIFAR-100 already on disk under ./data the behaviour is the reference's (torchvision dataset,
Resize(64)/ToTensor/Normalize(.5,.5) for training, Resize(299)/ImageNet-normalise for the metric
loader, shuffle=True, pin_memory=True; dcgan_data_preprocessor.py:37-75).  Without it -- there is no
network here, and the reference's download=True cannot work -- or with args.synthetic=1, a synthetic
source with the same contract is used (preprocess/synthetic.py)."""
import os

import torch

from ..logger.main_logger import MainLogger
from .synthetic import SyntheticLoader


def _cifar_available(root="./data"):
    return os.path.isdir(os.path.join(root, "cifar-100-python"))


class DCGANDataPreprocessor:
    def __init__(self, args):
        self._logger = MainLogger(args)
        self.batch_size = args.batch_size
        self.num_worker = getattr(args, "num_worker", 0)
        self.synthetic = bool(getattr(args, "synthetic", 0)) or not _cifar_available()
        self.synthetic_batches = int(getattr(args, "synthetic_batches", 391))   # 50000 / 128
        self._trainset = self._inceptionset = None
        if not self.synthetic:
            import torchvision
            self._trainset = torchvision.datasets.CIFAR100("./data", train=True, download=False, transform=None)
            self._inceptionset = torchvision.datasets.CIFAR100("./data", train=True, download=False, transform=None)
        self._logger.debug('data preprocessor init' + (' (synthetic source)' if self.synthetic else ''))

    def transform_data(self):
        if self.synthetic:
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
        if self.synthetic:
            self.trainloader = SyntheticLoader(self.batch_size, self.synthetic_batches)
            self.inceptionloader = None
        else:
            self.trainloader = torch.utils.data.DataLoader(self._trainset, self.batch_size, shuffle=True,
                                                           num_workers=self.num_worker, pin_memory=True)
            self.inceptionloader = torch.utils.data.DataLoader(self._inceptionset, self.batch_size * 2,
                                                               pin_memory=True, num_workers=0)
        return self.trainloader, self.inceptionloader
