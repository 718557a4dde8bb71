"""B200-native drop-in for the reference's loss/losses_duett.py: same classes, constructor arguments, buffers
(`label_weights`, `pos_weight`, `eps`) and returned dict keys; every loss is one fused CUDA kernel that also yields
the gradient w.r.t. the logits (csrc/dx_loss.cu)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..functional import BCELogitsFn, KDLossFn, MaskedBceColsFn


class VanillaKLKD(nn.Module):
    """Binary KL(P_teacher || P_student) with temperature T, scaled by T^2 (loss/losses_duett.py:8-25)."""

    def __init__(self, T: float = 4.0, eps: float = 1e-7):
        super().__init__()
        self.T = T
        self.eps = eps

    def forward(self, z_s: torch.Tensor, z_t: torch.Tensor) -> torch.Tensor:
        total, _, _ = KDLossFn.apply(z_s, z_t.detach(), torch.zeros_like(z_s, dtype=torch.float32), self.T, 0.0, None, self.eps)
        return total


KD_LOSSES = {
    "vanilla_kl": VanillaKLKD,
}


def build_kd_loss(name: str, **kwargs) -> nn.Module:
    if name not in KD_LOSSES:
        raise ValueError(f"unknown KD loss: {name!r}. available: {list(KD_LOSSES)}")
    return KD_LOSSES[name](**kwargs)


class StudentKDLoss(nn.Module):
    """total = alpha * BCE(z_s, y) + (1 - alpha) * L_kd(z_s, z_t)  (loss/losses_duett.py:39-57)."""

    def __init__(self, kd_name: str = "vanilla_kl", kd_T: float = 4.0, kd_alpha: float = 0.5,
                 pos_weight: float | None = None):
        super().__init__()
        self.alpha = kd_alpha
        self.kd = build_kd_loss(kd_name, T=kd_T)
        self.pos_weight_value = None if pos_weight is None else float(pos_weight)
        pw = None if pos_weight is None else torch.tensor([pos_weight], dtype=torch.float32)
        self.bce = nn.BCEWithLogitsLoss(pos_weight=pw)      # holder for the pos_weight buffer (state-dict parity)

    def forward(self, z_s: torch.Tensor, z_t: torch.Tensor, y: torch.Tensor) -> dict:
        if isinstance(self.kd, VanillaKLKD):
            total, bce, kd = KDLossFn.apply(z_s, z_t.detach(), y.float(), self.kd.T, self.alpha, self.pos_weight_value,
                                            self.kd.eps)
            return {"total": total, "bce": bce.detach(), "kd": kd.detach()}
        loss_kd = self.kd(z_s, z_t)
        pw = 1.0 if self.pos_weight_value is None else self.pos_weight_value
        loss_bce = BCELogitsFn.apply(z_s, y.float(), pw, 1.0)
        total = self.alpha * loss_bce + (1.0 - self.alpha) * loss_kd
        return {"total": total, "bce": loss_bce.detach(), "kd": loss_kd.detach()}


class _MaskedMultiLabel(nn.Module):
    def __init__(self, label_weights, pos_weight, eps):
        super().__init__()
        self.register_buffer("label_weights", label_weights.float())
        if pos_weight is not None:
            self.register_buffer("pos_weight", pos_weight.float())
        else:
            self.pos_weight = None
        self.eps = eps
        self.n_pathologies = int(label_weights.numel())

    def _per_pathology_bce(self, logits, y, mask):
        """logits/y/mask: [B,K] -> [K]: sum_b(bce*m) / (sum_b m + eps), all K columns in one kernel."""
        pw = None if self.pos_weight is None else self.pos_weight.contiguous()
        return MaskedBceColsFn.apply(logits, y, mask, pw, self.eps)

    def _weighted(self, per):
        # sum_k w_k * per_k on a [K] vector (K = 7): scalar glue
        return (self.label_weights * per).sum()


class PathologyMultiLabelLoss(_MaskedMultiLabel):
    """loss/losses_duett.py:63-129."""

    def __init__(self, label_weights: torch.Tensor, pos_weight: torch.Tensor | None = None, alpha_stage2: float = 0.5,
                 alpha_stage4: float = 1.0, eps: float = 1e-6):
        super().__init__(label_weights, pos_weight, eps)
        self.alpha_stage2, self.alpha_stage4 = float(alpha_stage2), float(alpha_stage4)

    def forward(self, stage2_logits, stage4_logits, y_multi, y_multi_mask) -> dict:
        s2_per = self._per_pathology_bce(stage2_logits, y_multi, y_multi_mask)
        s4_per = self._per_pathology_bce(stage4_logits, y_multi, y_multi_mask)
        s2_total, s4_total = self._weighted(s2_per), self._weighted(s4_per)
        total = self.alpha_stage2 * s2_total + self.alpha_stage4 * s4_total
        return {"total": total, "stage2_total": s2_total.detach(), "stage4_total": s4_total.detach(),
                "stage2_per": s2_per.detach(), "stage4_per": s4_per.detach()}


class DualPathologyLoss(_MaskedMultiLabel):
    """loss/losses_duett.py:135-194."""

    def __init__(self, label_weights: torch.Tensor, pos_weight: torch.Tensor | None = None, alpha_img: float = 0.5,
                 alpha_ts: float = 0.5, alpha_fus: float = 1.0, eps: float = 1e-6):
        super().__init__(label_weights, pos_weight, eps)
        self.alpha_img, self.alpha_ts, self.alpha_fus = float(alpha_img), float(alpha_ts), float(alpha_fus)

    def forward(self, img_logits, ts_logits, fusion_logits, y_multi, y_multi_mask) -> dict:
        img_per = self._per_pathology_bce(img_logits, y_multi, y_multi_mask)
        ts_per = self._per_pathology_bce(ts_logits, y_multi, y_multi_mask)
        fus_per = self._per_pathology_bce(fusion_logits, y_multi, y_multi_mask)
        img_total, ts_total, fus_total = self._weighted(img_per), self._weighted(ts_per), self._weighted(fus_per)
        total = self.alpha_img * img_total + self.alpha_ts * ts_total + self.alpha_fus * fus_total
        return {"total": total, "img_total": img_total.detach(), "ts_total": ts_total.detach(),
                "fus_total": fus_total.detach(), "img_per": img_per.detach(), "ts_per": ts_per.detach(),
                "fus_per": fus_per.detach()}
