"""Drop-in for the reference's preprocess/cgan_data_preprocessor.py: `OneHotEncoder`,
`CGANDataPreprocessor(args)` with `.idx_to_labels`, `.transform_data()`, `.get_data_loader()` (which,
as in the reference :90, returns the inception *dataset*, not a loader)."""
import torch

from ..logger.main_logger import MainLogger
from .dcgan_data_preprocessor import _cifar100, u8_source, IMAGENET_MEAN, IMAGENET_STD
from .synthetic import SyntheticLoader


class OneHotEncoder:
    def __init__(self, label_count):
        self.label_count = label_count

    def __call__(self, label):
        out = torch.zeros(self.label_count, dtype=torch.int64)
        out[label] = 1
        return out


class CGANDataPreprocessor:
    def __init__(self, args):
        self._logger = MainLogger(args)
        self.batch_size = args.batch_size
        self.num_worker = getattr(args, "num_worker", 0)
        self.n_classes = int(getattr(args, "n_classes", 100))
        # synthetic images ONLY on request (args.synthetic / --synthetic); otherwise the reference's behaviour: CIFAR-100 from
        # ./data, downloaded when missing (download=True, :20-21), and an error when that fails -- never a silent stand-in
        self.synthetic = bool(getattr(args, "synthetic", 0))
        self.synthetic_batches = int(getattr(args, "synthetic_batches", 391))
        self._trainset = self._inceptionset = None
        self.idx_to_labels = {i: str(i) for i in range(self.n_classes)}
        if not self.synthetic:
            import torchvision
            self._trainset = _cifar100(torchvision)
            self._inceptionset = _cifar100(torchvision)
            self.idx_to_labels = {v: k for k, v in self._trainset.class_to_idx.items()}
        self._u8 = u8_source(self, args, self.n_classes)
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
        self._trainset.target_transform = OneHotEncoder(label_count=len(self._trainset.classes))
        self._logger.debug('data transform')

    def get_data_loader(self):
        # data parallel (torchrun): args.batch_size is the GLOBAL batch.  The device and synthetic loaders hand every rank its
        # own rows; the host DataLoader yields global batches and the trainer takes the rank's rows (`global_batches`).
        from ..parallel import env_rank_world
        rank, world = env_rank_world()
        self.global_batches = self._u8 is None and not self.synthetic
        if self._u8 is not None:
            # Resize(64) / ToTensor / Normalize + OneHotEncoder (:49-62) and the loaders (:82-92) on the device
            from .device_pipeline import DeviceImageLoader
            data, targets = self._u8
            n_cls = self.n_classes if self.synthetic else len(self._trainset.classes)
            self.trainloader = DeviceImageLoader(data, targets, self.batch_size, 64, [0.5] * 3, [0.5] * 3, shuffle=True,
                                                 n_classes=n_cls, rank=rank, world=world)
            self.inceptionloader = DeviceImageLoader(data, targets, 128, (299, 299), IMAGENET_MEAN, IMAGENET_STD, shuffle=False)
        elif self.synthetic:
            self.trainloader = SyntheticLoader(self.batch_size, self.synthetic_batches, n_classes=self.n_classes, rank=rank, world=world)
            self.inceptionloader = None
        else:
            self.trainloader = torch.utils.data.DataLoader(self._trainset, self.batch_size, shuffle=True,
                                                           num_workers=self.num_worker, pin_memory=True)
            self.inceptionloader = self._inceptionset
        return self.trainloader, self.inceptionloader
