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


def _ap_like_sklearn(auprc: float, n_pos: float) -> float:
    """dx_binary_auc reports the average precision of a label set without positives as NaN (undefined).  The reference scores
    with sklearn's average_precision_score, which since scikit-learn 1.1 warns and returns 0.0 there ("no positive class
    found in y_true, recall is set to one for all thresholds") — and that 0.0 enters the reference's macro means
    (training_duett/evaluator.py:322-331).  Follow the installed-sklearn behaviour of the reference."""
    return 0.0 if n_pos == 0 else auprc


def binary_metrics(logits: torch.Tensor, y: torch.Tensor) -> dict:
    """{"auroc","auprc","n","pos_frac"} of sigmoid(logits) against y, computed on the device (one D2H read of 4 doubles)."""
    logits = _gather_all(logits.detach().reshape(-1).float())
    y = _gather_all(y.detach().reshape(-1).float())
    if logits.numel() == 0:
        return {"auroc": float("nan"), "auprc": float("nan"), "n": 0, "pos_frac": float("nan")}
    auroc, auprc, n_pos, n = ops.binary_auc(logits, y, apply_sigmoid=True).tolist()
    return {"auroc": float(auroc), "auprc": float(_ap_like_sklearn(auprc, n_pos)), "n": int(n), "pos_frac": float(n_pos / n)}


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
        # an empty shard still takes part in the cross-rank gather (returning early here would leave the other ranks waiting)
        empty = torch.empty(0, device=torch.device(device), dtype=torch.float32)
        return binary_metrics(empty, empty.clone())
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


def _bce_mean(logits: torch.Tensor, y: torch.Tensor) -> float:
    """mean of max(l,0) - l*y + log1p(exp(-|l|)) (training_duett/evaluator.py:181-183)."""
    if logits.numel() == 0:
        return float("nan")
    l, t = logits.double(), y.double()
    return float((l.clamp_min(0) - l * t + torch.log1p(torch.exp(-l.abs()))).mean())


def _pearson(a: torch.Tensor, b: torch.Tensor) -> float:
    """Population Pearson correlation, NaN for fewer than two samples or zero variance (evaluator.py:186-194)."""
    if a.numel() < 2:
        return float("nan")
    a, b = a.double(), b.double()
    da, db = a - a.mean(), b - b.mean()
    va, vb = float((da * da).mean()), float((db * db).mean())
    if va == 0 or vb == 0:
        return float("nan")
    return float((da * db).mean() / (va ** 0.5 * vb ** 0.5))


@torch.no_grad()
def evaluate_dual_pathology(model, loader, device, pathology_labels, *, query_ref=None) -> dict:
    """Per-pathology evaluation of a (Patch)DualPathologyPerceiver teacher — same result dict as the reference's
    training_duett/evaluator.py:197-335: for every label the masked AUROC / AUPRC of the image, time-series and fusion
    branches with their gaps, per-branch BCE and delta, residual-correction usage and its correlation with the image
    branch's error, beta; macro means as main_auroc / main_auprc.  Logits stay on the device (every rank sees the whole
    loader, see binary_metrics); ranking runs in dx_binary_auc."""
    from .engine import _move_lists

    model.eval()
    cols = {k: [] for k in ("img_logits", "ts_logits", "fusion_logits", "scaled_correction", "y", "mask")}
    for batch in loader:
        b = _move_lists(batch, device)
        out = model(b["x_ts"], b["x_static"], b["bin_ends"], b["pixel_values"])
        if not isinstance(out, dict) or "fusion_logits" not in out:
            raise RuntimeError("evaluate_dual_pathology needs a dual_pathology_mode teacher")
        for k in ("img_logits", "ts_logits", "fusion_logits"):
            cols[k].append(out[k].detach().float())
        if "scaled_correction" in out:
            cols["scaled_correction"].append(out["scaled_correction"].detach().float())
        cols["y"].append(b["y_multi"].float())
        cols["mask"].append(b["y_multi_mask"].float())
    K = len(pathology_labels)

    def table(k):
        t = torch.cat(cols[k])
        return torch.stack([_gather_all(t[:, j].contiguous()) for j in range(K)], 1)

    img, ts, fus, y, mk = (table(k) for k in ("img_logits", "ts_logits", "fusion_logits", "y", "mask"))
    corr = table("scaled_correction") if cols["scaled_correction"] else None
    unwrapped = model.module if hasattr(model, "module") else model
    perceiver = getattr(unwrapped, "perceiver", None)
    beta = perceiver.beta.detach().float().cpu() if perceiver is not None and hasattr(perceiver, "beta") else None
    nan = float("nan")

    def ranked(logits, yk):
        if yk.numel() == 0:
            return nan, nan
        r = ops.binary_auc(logits, yk, apply_sigmoid=True).tolist()
        return float(r[0]), float(_ap_like_sklearn(r[1], r[2]))

    per_label = []
    for k in range(K):
        m = mk[:, k] > 0.5
        yk = y[m, k].contiguous()
        li, lt, lf = (t[m, k].contiguous() for t in (img, ts, fus))
        (ai, ri), (at, rt), (af, rf) = ranked(li, yk), ranked(lt, yk), ranked(lf, yk)
        ib, tb, fb = _bce_mean(li, yk), _bce_mean(lt, yk), _bce_mean(lf, yk)
        if corr is not None and yk.numel():
            ck = corr[m, k]
            mean_abs_corr, corr_r = float(ck.abs().double().mean()), _pearson(ck, yk - torch.sigmoid(li.double()).float())
        else:
            mean_abs_corr, corr_r = nan, nan
        per_label.append({
            "name": pathology_labels[k], "n_valid": int(m.sum()), "pos_frac": float(yk.double().mean()) if yk.numel() else nan,
            "img_auroc": ai, "ts_auroc": at, "fus_auroc": af, "gap_i2f": af - ai, "gap_t2f": af - at,
            "img_auprc": ri, "ts_auprc": rt, "fus_auprc": rf, "gap_i2f_pr": rf - ri, "gap_t2f_pr": rf - rt,
            "img_bce": ib, "ts_bce": tb, "fus_bce": fb, "delta_bce": fb - ib,
            "mean_abs_corr": mean_abs_corr, "corr_residual": corr_r, "beta": float(beta[k]) if beta is not None else nan,
        })

    def macro(key):
        vals = [r[key] for r in per_label if r[key] == r[key]]
        return sum(vals) / len(vals) if vals else nan

    return {"labels": list(pathology_labels), "n": int(y.shape[0]), "main_auroc": macro("fus_auroc"),
            "main_auprc": macro("fus_auprc"), "per_label": per_label}


# ---- console tables (training_duett/trainer.py:24-25 imports these next to the evaluators) ---------------------------------
def _cell(v, spec):
    """One table cell; NaN / non-numeric -> right-aligned '--' of the column's width."""
    width = int(spec.lstrip("+").split(".")[0])
    try:
        if v != v:
            return "--".rjust(width)
        return format(v, spec)
    except (TypeError, ValueError):
        return "--".rjust(width)


# (header, result key, format spec, separator AFTER the column) of format_dual_pathology_gap_table, left to right
_DUAL_COLS = (("imgROC", "img_auroc", "7.3f", " "), ("tsROC", "ts_auroc", "7.3f", " "), ("fusROC", "fus_auroc", "7.3f", " "),
              ("gain", "gap_i2f", "+7.3f", "  "), ("imgAP", "img_auprc", "6.3f", " "), ("tsAP", "ts_auprc", "6.3f", " "),
              ("fusAP", "fus_auprc", "6.3f", "  "), ("dBCE", "delta_bce", "+7.4f", "  "), ("|corr|", "mean_abs_corr", "7.4f", " "),
              ("corr_r", "corr_residual", "+7.3f", "  "), ("beta", "beta", "6.3f", ""))


def format_dual_pathology_gap_table(result: dict) -> str:
    """The residual-fusion summary table of evaluate_dual_pathology's result (training_duett/evaluator.py:350-395): one row per
    pathology (image / time-series / fusion AUROC, fusion gain, the three APs, BCE delta, correction usage, beta) and a
    closing macro-mAP row; NaNs print as '--'."""
    def row(label, cells):
        out = f"{label:<12s} "
        for (_, _, spec, sep), c in zip(_DUAL_COLS, cells):
            out += c + sep
        return out

    width = lambda spec: int(spec.lstrip("+").split(".")[0])
    header = row("label", [h.rjust(width(spec)) for h, _, spec, _ in _DUAL_COLS])
    rule = "-" * len(header)
    lines = [header, rule]
    per = result["per_label"]
    for r in per:
        lines.append(row(r["name"].replace("label_", ""), [_cell(r[key], spec) for _, key, spec, _ in _DUAL_COLS]))

    def macro(key):
        vals = [r[key] for r in per if not (isinstance(r[key], float) and r[key] != r[key])]
        return sum(vals) / len(vals) if vals else float("nan")

    # closing row: blanks under the four AUROC columns, the macro means under the three AP columns (the row ends there)
    tail = f"{'mAP (macro)':<12s} "
    for i, (_, key, spec, sep) in enumerate(_DUAL_COLS[:7]):
        tail += (_cell(macro(key), spec) if key.endswith("_auprc") else " " * width(spec)) + (sep if i < 6 else "")
    lines += [rule, tail]
    return "\n".join(lines)


def evaluate_pathology(model, loader, device, pathology_labels) -> dict:
    """Evaluator of the legacy `pathology_mode` teacher (stage2 / stage4 logits, training_duett/evaluator.py:100-160).  That
    TeacherModel branch is dead in the reference snapshot and not built here (TeacherModel raises for it), so this raises too
    instead of scoring something else."""
    raise NotImplementedError("evaluate_pathology scores the legacy pathology_mode teacher, which is not implemented "
                              "(only patch_dual_pathology_mode is live in the reference): use evaluate_dual_pathology")


def format_pathology_gap_table(result: dict) -> str:
    """Table of an evaluate_pathology result dict (training_duett/evaluator.py:163-178); kept for import compatibility."""
    cols = (("s2_auroc", "stage2_auroc", "10.4f"), ("s4_auroc", "stage4_auroc", "10.4f"), ("gap_ro", "gap_auroc", "+8.4f"),
            ("s2_auprc", "stage2_auprc", "10.4f"), ("s4_auprc", "stage4_auprc", "10.4f"), ("gap_pr", "gap_auprc", "+8.4f"))
    width = lambda spec: int(spec.lstrip("+").split(".")[0])
    lines = [f"{'label':<22s} {'n':>6s} {'pos':>7s} " + " ".join(h.rjust(width(s)) for h, _, s in cols)]
    for r in result["per_label"]:
        lines.append(f"{r['name']:<22s} {r['n_valid']:>6d} {r['pos_frac']:>7.4f} " + " ".join(format(r[k], s) for _, k, s in cols))
    return "\n".join(lines)
