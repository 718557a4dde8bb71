"""ORACLE / TEST INFRASTRUCTURE ONLY — AUROC / AveragePrecision stand-ins (sklearn-backed) for duett/duett.py:135-140."""
import numpy as np
import torch
import torch.nn as nn


class _Metric(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        self._p, self._y = [], []

    def update(self, preds, target):
        self._p.append(preds.detach().float().cpu()); self._y.append(target.detach().cpu())

    def reset(self):
        self._p, self._y = [], []

    def _cat(self):
        return torch.cat(self._p).numpy(), torch.cat(self._y).numpy()


class AUROC(_Metric):
    def compute(self):
        from sklearn.metrics import roc_auc_score
        p, y = self._cat()
        return torch.tensor(roc_auc_score(y, p))


class AveragePrecision(_Metric):
    def compute(self):
        from sklearn.metrics import average_precision_score
        p, y = self._cat()
        return torch.tensor(average_precision_score(y, p))
