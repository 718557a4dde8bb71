"""Binary AUROC/AUPRC evaluation for teacher/student — same surface as the reference's training_duett/evaluator.py:10-78
(`evaluate_binary(model, loader, device, forward_fn)`, `make_teacher_forward`, `make_teacher_aux_forward`,
`make_student_forward`), scored on the device.

Differences from the reference, both deliberate:
  * logits and labels never leave the GPU: they are concatenated on the device and ranked by `dx_binary_auc`
    (bitonic sort + threshold scan; sklearn's roc_auc_score / average_precision_score definitions, ties included);
  * with torch.distributed initialised every rank's shard is gathered first, so all ranks report the metric of the WHOLE
    loader (the reference scores rank 0's shard only).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .. import ops


def _gather_all(t: torch.Tensor) -> torch.Tensor:
    """Concatenate the 1-D shards of every rank (shards may differ in length)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    world = dist.get_world_size()
    n = torch.tensor([t.numel()], device=t.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s) for s in sizes]
    pad = torch.zeros(max(sizes), device=t.device, dtype=t.dtype)
    pad[: t.numel()] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])


def binary_metrics(logits: torch.Tensor, y: torch.Tensor) -> dict:
    """{"auroc","auprc","n","pos_frac"} of sigmoid(logits) against y, computed on the device (one D2H read of 4 doubles)."""
    logits = _gather_all(logits.detach().reshape(-1).float())
    y = _gather_all(y.detach().reshape(-1).float())
    auroc, auprc, n_pos, n = ops.binary_auc(logits, y, apply_sigmoid=True).tolist()
    return {"auroc": float(auroc), "auprc": float(auprc), "n": int(n), "pos_frac": float(n_pos / n)}


@torch.no_grad()
def evaluate_binary(model, loader, device, forward_fn):
    """Aggregate logits/labels across a loader and compute AUROC/AUPRC.

    forward_fn(model, batch, device) -> dict with keys `logits`, `y` (teacher and student loaders go through the same
    indirection as in the reference)."""
    model.eval()
    logits_all, y_all = [], []
    for batch in loader:
        out = forward_fn(model, batch, device)
        logits_all.append(out["logits"].detach().reshape(-1).float())
        y_all.append(out["y"].detach().reshape(-1).float().to(out["logits"].device))
    if not logits_all:
        return {"auroc": float("nan"), "auprc": float("nan"), "n": 0, "pos_frac": float("nan")}
    return binary_metrics(torch.cat(logits_all), torch.cat(y_all))


def _main_logit(out):
    """Teacher output -> main-head logits: dict (pathology modes) -> "main_logit", legacy tuple -> first entry."""
    if isinstance(out, dict):
        return out["main_logit"]
    return out[0] if isinstance(out, tuple) else out


def _aux_logit(out):
    if not isinstance(out, tuple):
        raise RuntimeError("aux forward is only available with TeacherModel(use_aux_cxr=True)")
    return out[1]


def _make_forward(with_image: bool, pick):
    """forward_fn factory for evaluate_binary: moves the collate-format batch, runs the model, picks the scored logits."""
    from .engine import _move_lists

    @torch.no_grad()
    def _fwd(model, batch, device):
        b = _move_lists(batch, device)
        inputs = (b["x_ts"], b["x_static"], b["bin_ends"]) + ((b["pixel_values"],) if with_image else ())
        return {"logits": pick(model(*inputs)), "y": b["y"]}

    return _fwd


def make_teacher_forward():
    """Main-head evaluation of a teacher (training_duett/evaluator.py:40-60)."""
    return _make_forward(True, _main_logit)


def make_teacher_aux_forward():
    """Auxiliary CXR-only head; valid only when the teacher returns a tuple (training_duett/evaluator.py:63-76)."""
    return _make_forward(True, _aux_logit)


def make_student_forward():
    """Student evaluation (training_duett/evaluator.py:79-88)."""
    return _make_forward(False, lambda z: z)
