"""ORACLE / TEST INFRASTRUCTURE ONLY — the slice of lightning.pytorch the reference's DuETT module touches
(duett/duett.py:5,48,371,459-487; models/main_architecture_duett.py:106-118)."""
import torch
import torch.nn as nn


class LightningModule(nn.Module):
    current_epoch = 0

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def log(self, *a, **k):
        pass

    def on_load_checkpoint(self, checkpoint):
        pass

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, strict=True, map_location="cpu", **kwargs):
        ckpt = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
        model = cls(**kwargs)
        model.on_load_checkpoint(ckpt)
        model.load_state_dict(ckpt["state_dict"], strict=strict)
        return model


class Callback:
    pass


class Trainer:
    def __init__(self, *a, **k):
        raise NotImplementedError("shim: pl.Trainer is not part of the oracle")


def seed_everything(seed, workers=False):
    import random
    import numpy as np
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    return seed
