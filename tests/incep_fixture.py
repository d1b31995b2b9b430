"""Test fixture: the reference's Inception-v3 (metrics.py:46-52: torchvision inception_v3, fc -> Linear(2048, 100)) with
seeded random weights whose BatchNorm running statistics are CALIBRATED on a random batch.  Without the reference's
checkpoint (./save/iception_v3/loss_bset.pt is not in the repository and there is no network) an un-calibrated random
network maps every image to the same logits to 1e-10, which would make any comparison vacuous; after calibration every
layer's activations are O(1) and image dependent."""
import torch
import torch.nn as nn


def calibrated_inception(seed=1, n_out=100, calib_batch=2):
    from torchvision import models
    torch.manual_seed(seed)
    m = models.inception_v3(weights=None, aux_logits=True, init_weights=False)
    m.aux_logits = False                             # metrics.py:48; AuxLogits stays in the state_dict, as in the reference's checkpoint
    m.fc = nn.Sequential(nn.Linear(m.fc.in_features, n_out))
    g = torch.Generator().manual_seed(seed + 1)
    for mod in m.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.momentum = 1.0                       # running stats <- the calibration batch's statistics
            with torch.no_grad():
                mod.weight.copy_(0.75 + 0.5 * torch.rand(mod.weight.shape, generator=g))
                mod.bias.copy_(0.2 * torch.rand(mod.bias.shape, generator=g) - 0.05)
    m.train()
    with torch.no_grad():
        m(torch.randn(calib_batch, 3, 299, 299, generator=g))
    m.eval()
    return m
