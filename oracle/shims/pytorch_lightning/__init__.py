"""Import-only stub for pytorch_lightning (PL 1.6.5 in the reference env, not installed). TEST INFRASTRUCTURE ONLY."""
import torch.nn as nn


class LightningModule(nn.Module):
    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass


class LightningDataModule:
    pass


def seed_everything(seed):
    import torch
    torch.manual_seed(seed)
