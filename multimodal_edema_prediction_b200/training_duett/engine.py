"""B200-native drop-in for the reference's training_duett/engine.py: the same one-step train/eval functions with the
same signatures and returned dict keys (training_duett/engine.py:7-308).  The steps call the B200 modules
(models/main_architecture_duett.py, loss/losses_duett.py); the only arithmetic done here — the auxiliary residual KL of
engine.py:149-165 — is a fused CUDA kernel too."""
from __future__ import annotations

import torch

from ..functional import AuxResidualKLFn


def _set_train_with_frozen_eval(teacher, accelerator=None):
    """teacher.train(), but fully-frozen submodules go to eval() (frozen BN then uses running statistics) —
    training_duett/engine.py:7-20."""
    teacher.train()
    unwrapped = accelerator.unwrap_model(teacher) if accelerator is not None else teacher
    for attr in ("duett", "cxr", "pretrained_cxr_head"):
        mod = getattr(unwrapped, attr, None)
        if mod is None:
            continue
        params = list(mod.parameters(recurse=True))
        if params and not any(p.requires_grad for p in params):
            mod.eval()


def _move_lists(batch: dict, device: torch.device) -> dict:
    """Host->device boundary of the collate format (training_duett/engine.py:23-36).  The reference moves the 3*B per-sample
    tensors one by one; here the per-sample tuples stay where they are (host tensors, or StayRows records from
    MIMICDataset) and the modules' feats_to_input stacks each of them into pinned staging memory and uploads it with ONE
    asynchronous copy (Model._upload) — 3 H2D copies per batch instead of 3*B.  Samples that are already on the device are
    passed through.  The batch-level tensors go up non-blocking as in the reference."""
    out = {"x_ts": tuple(batch["x_ts"]), "x_static": tuple(batch["x_static"]), "bin_ends": tuple(batch["bin_ends"]),
           "y": batch["y"].to(device, non_blocking=True)}
    for k in ("pixel_values", "y_multi", "y_multi_mask"):
        if k in batch:
            out[k] = batch[k].to(device, non_blocking=True)
    return out


def _backward_step(loss, optimizer, accelerator):
    optimizer.zero_grad()
    if accelerator is not None:
        accelerator.backward(loss)
    else:
        loss.backward()
    optimizer.step()


def _aux_residual(out, b, device, aux_residual_alpha):
    if aux_residual_alpha > 0.0 and "scaled_correction" in out:
        return AuxResidualKLFn.apply(out["img_logits"], out["scaled_correction"], b["y_multi"].float(),
                                     b["y_multi_mask"].float(), 0.05)
    return torch.zeros((), device=device)


def train_teacher_batch(batch, teacher, loss_fn, optimizer, device, accelerator=None, aux_alpha: float = 0.0):
    """Binary teacher step for a teacher that returns the main logit, or a (main, aux) tuple whose auxiliary CXR-only logit
    is scored with the same loss at weight aux_alpha (training_duett/engine.py:41-74).  The legacy TeacherModel modes that
    return these are not built here (models/main_architecture_duett.py), the step itself is model-agnostic."""
    teacher.train()
    b = _move_lists(batch, device)
    out = teacher(b["x_ts"], b["x_static"], b["bin_ends"], b["pixel_values"])
    main_logit, aux_logit = out if isinstance(out, tuple) else (out, None)
    y = b["y"].float()
    main_loss = loss_fn(main_logit, y)
    loss, aux_value = main_loss, 0.0
    if aux_logit is not None:
        aux_loss = loss_fn(aux_logit, y)
        loss = main_loss + aux_alpha * aux_loss
        aux_value = aux_loss.detach().item()
    _backward_step(loss, optimizer, accelerator)
    return {"loss": loss.detach().item(), "main_loss": main_loss.detach().item(), "aux_loss": aux_value,
            "logits": main_logit.detach(), "y": b["y"].detach()}


def train_teacher_pathology_batch(batch, teacher, path_loss_fn, optimizer, device, accelerator=None):
    """Step of a `pathology_mode` teacher returning dict(main_logit, stage2_logits, stage4_logits), scored with
    PathologyMultiLabelLoss (training_duett/engine.py:93-130).  Model-agnostic like train_teacher_batch."""
    teacher.train()
    b = _move_lists(batch, device)
    out = teacher(b["x_ts"], b["x_static"], b["bin_ends"], b["pixel_values"])
    if not isinstance(out, dict):
        raise RuntimeError("pathology mode but TeacherModel did not return a dict")
    losses = path_loss_fn(out["stage2_logits"], out["stage4_logits"], b["y_multi"], b["y_multi_mask"])
    _backward_step(losses["total"], optimizer, accelerator)
    res = {"loss": losses["total"].detach().item(), "stage2_total": losses["stage2_total"].item(),
           "stage4_total": losses["stage4_total"].item(), "stage2_per": losses["stage2_per"].cpu(),
           "stage4_per": losses["stage4_per"].cpu()}
    res.update({k: out[k].detach() for k in ("main_logit", "stage2_logits", "stage4_logits")})
    res.update({k: b[k].detach() for k in ("y", "y_multi", "y_multi_mask")})
    return res


def train_teacher_dual_pathology_batch(batch, teacher, path_loss_fn, optimizer, device, accelerator=None,
                                       aux_residual_alpha: float = 0.0):
    """training_duett/engine.py:135-190."""
    _set_train_with_frozen_eval(teacher, accelerator)
    b = _move_lists(batch, device)
    out = teacher(b["x_ts"], b["x_static"], b["bin_ends"], b["pixel_values"])
    if not isinstance(out, dict):
        raise RuntimeError("dual_pathology mode but TeacherModel did not return a dict")
    losses = path_loss_fn(out["img_logits"], out["ts_logits"], out["fusion_logits"], b["y_multi"], b["y_multi_mask"])
    total = losses["total"]
    aux_residual_loss = _aux_residual(out, b, device, aux_residual_alpha)
    if aux_residual_alpha > 0.0:
        total = total + aux_residual_alpha * aux_residual_loss
    _backward_step(total, optimizer, accelerator)
    return {
        "loss": total.detach().item(), "img_total": losses["img_total"].item(), "ts_total": losses["ts_total"].item(),
        "fus_total": losses["fus_total"].item(), "aux_residual": float(aux_residual_loss.detach().item()),
        "img_per": losses["img_per"].cpu(), "ts_per": losses["ts_per"].cpu(), "fus_per": losses["fus_per"].cpu(),
        "main_logit": out["main_logit"].detach(), "img_logits": out["img_logits"].detach(),
        "ts_logits": out["ts_logits"].detach(), "fusion_logits": out["fusion_logits"].detach(),
        "y": b["y"].detach(), "y_multi": b["y_multi"].detach(), "y_multi_mask": b["y_multi_mask"].detach(),
    }


def train_teacher_dual_pathology_lp_batch(batch, teacher, path_loss_fn, optimizer, device, accelerator=None,
                                          beta_l2: float = 0.0, corr_l2: float = 0.0, aux_residual_alpha: float = 0.0):
    """LP stage: everything eval() except perceiver.correction_head (training_duett/engine.py:196-264)."""
    teacher.eval()
    unwrapped = accelerator.unwrap_model(teacher) if accelerator is not None else teacher
    unwrapped.perceiver.correction_head.train()
    b = _move_lists(batch, device)
    out = teacher(b["x_ts"], b["x_static"], b["bin_ends"], b["pixel_values"])
    if not isinstance(out, dict):
        raise RuntimeError("dual_pathology LP mode but TeacherModel did not return a dict")
    losses = path_loss_fn(out["img_logits"], out["ts_logits"], out["fusion_logits"], b["y_multi"], b["y_multi_mask"])
    reg_beta = torch.zeros((), device=device)
    reg_corr = torch.zeros((), device=device)
    if beta_l2 > 0.0:
        reg_beta = beta_l2 * (unwrapped.perceiver.beta ** 2).mean()          # [K]-vector regulariser (scalar glue)
    if corr_l2 > 0.0:
        reg_corr = corr_l2 * (out["scaled_correction"] ** 2).mean()
    aux_residual_loss = _aux_residual(out, b, device, aux_residual_alpha)
    total = losses["total"] + reg_beta + reg_corr + aux_residual_alpha * aux_residual_loss
    _backward_step(total, optimizer, accelerator)
    return {
        "loss": total.detach().item(), "img_total": losses["img_total"].item(), "ts_total": losses["ts_total"].item(),
        "fus_total": losses["fus_total"].item(), "img_per": losses["img_per"].cpu(), "ts_per": losses["ts_per"].cpu(),
        "fus_per": losses["fus_per"].cpu(), "reg_beta_l2": float(reg_beta.detach().item()),
        "reg_corr_l2": float(reg_corr.detach().item()), "aux_residual": float(aux_residual_loss.detach().item()),
        "main_logit": out["main_logit"].detach(), "img_logits": out["img_logits"].detach(),
        "ts_logits": out["ts_logits"].detach(), "fusion_logits": out["fusion_logits"].detach(),
        "y": b["y"].detach(), "y_multi": b["y_multi"].detach(), "y_multi_mask": b["y_multi_mask"].detach(),
    }


@torch.no_grad()
def eval_teacher_batch(batch, teacher, loss_fn, device):
    """training_duett/engine.py:77-89."""
    teacher.eval()
    b = _move_lists(batch, device)
    out = teacher(b["x_ts"], b["x_static"], b["bin_ends"], b["pixel_values"])
    main_logit = out["main_logit"] if isinstance(out, dict) else (out[0] if isinstance(out, tuple) else out)
    loss = loss_fn(main_logit, b["y"].float())
    return {"loss": loss.item(), "logits": main_logit, "y": b["y"]}


def train_student_batch(batch_stu, batch_tea, student, teacher, kd_loss_fn, optimizer, device, accelerator=None):
    """Student KD step (training_duett/engine.py:270-301): frozen teacher forward under no_grad, student forward,
    StudentKDLoss, backward, optimizer step."""
    student.train()
    teacher.eval()
    b_s = _move_lists(batch_stu, device)
    b_t = _move_lists(batch_tea, device)
    with torch.no_grad():
        z_t = teacher(b_t["x_ts"], b_t["x_static"], b_t["bin_ends"], b_t["pixel_values"])["main_logit"]
    z_s = student(b_s["x_ts"], b_s["x_static"], b_s["bin_ends"])
    losses = kd_loss_fn(z_s, z_t, b_s["y"])
    _backward_step(losses["total"], optimizer, accelerator)
    return {"loss": losses["total"].detach().item(), "bce": losses["bce"].item(), "kd": losses["kd"].item(),
            "logits": z_s.detach(), "y": b_s["y"].detach()}


@torch.no_grad()
def eval_student_batch(batch, student, device):
    student.eval()
    b = _move_lists(batch, device)
    z = student(b["x_ts"], b["x_static"], b["bin_ends"])
    return {"logits": z, "y": b["y"]}
