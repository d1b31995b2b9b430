import torch


def get_default_device() -> torch.device:
    """reference utils.py:4-8.  The train step itself only exists for CUDA devices."""
    return torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')
