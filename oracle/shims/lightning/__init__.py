"""TEST INFRASTRUCTURE ONLY -- import shim for the absent `lightning` package.

The reference hot path (models/net.py etc.) only needs `LightningModule` to
behave like `torch.nn.Module` with `save_hyperparameters`, `log` and a
`device` property.  Nothing here is shipped or measured.
"""
import torch


class LightningModule(torch.nn.Module):
    def save_hyperparameters(self, *a, **k):
        return None

    def log(self, *a, **k):
        return None

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")


class LightningDataModule:  # placeholder, never instantiated on the hot path
    pass


class Trainer:  # placeholder
    def __init__(self, *a, **k):
        raise RuntimeError("lightning shim: training is out of scope")


def seed_everything(seed, workers=False):
    torch.manual_seed(seed)
    return seed
