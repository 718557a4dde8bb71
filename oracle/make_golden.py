"""ORACLE / TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the REFERENCE'S OWN FILES in place.

Run in the authoring container only (needs /root/reference, read-only):   python oracle/make_golden.py
The reference's duett/duett.py, models/main_architecture_duett.py, loss/losses_duett.py and training_duett/engine.py
are imported unmodified through oracle/shims (lightning / torchmetrics stand-ins, restated x_transformers.Encoder).
Each fixture stores: the reference module's state dict ("param/…"), the synthetic inputs ("in/…"), the reference's
outputs ("out/…") and parameter gradients ("grad/…").  tests/test_oracle_golden.py replays them through
oracle/duett_oracle.py; the GPU parity tests replay them through the CUDA path.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DUETT_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "shims"), REF, ROOT]

from oracle import duett_oracle as O  # noqa: E402


def np_(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def save(name, blobs):
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **{k: np_(v) for k, v in blobs.items()})
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB, {len(blobs)} arrays)")


def grads_of(module, prefix="grad/"):
    return {prefix + k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in module.named_parameters()}


def self_dev(module, blobs, step_fn):
    """How far the REFERENCE'S OWN bf16-autocast run lands from its own fp32 run on this fixture (global relative L2
    over all parameter gradients, and the loss).  The bf16-mode parity tests bound the CUDA path's deviation by this."""
    module.load_state_dict({k[len("param/"):]: v for k, v in blobs.items() if k.startswith("param/")})
    module.zero_grad()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        loss = step_fn()
    loss.backward()
    got = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).flatten().double()
                     for _, p in module.named_parameters()])
    want = torch.cat([blobs["grad/" + n].flatten().double() for n, _ in module.named_parameters()])
    return {"selfdev/global_grad": (got - want).norm() / want.norm(), "selfdev/loss_bf16": loss.detach().double()}


def params_of(module, prefix="param/"):
    return {prefix + k: v.clone() for k, v in module.state_dict().items()}


def main():
    torch.set_num_threads(4)
    from models.main_architecture_duett import (DuettFeatureExtractor, PatchDualPathologyPerceiver, StudentModel,
                                                TeacherModel)
    from duett.duett import Model
    from loss.losses_duett import DualPathologyLoss, StudentKDLoss
    from training_duett import engine
    torch.set_float32_matmul_precision("highest")   # reference sets 'high' (TF32) — irrelevant on CPU

    cfg = O.DuettConfig(d_static_num=3, d_time_series_num=5, n_timesteps=4, d_embedding=8, n_layers=2, d_feedforward=96)
    kw = dict(d_static_num=cfg.d_static_num, d_time_series_num=cfg.V, d_target=1, d_embedding=cfg.d_embedding,
              masked_transform_timesteps=cfg.T, max_len=cfg.T, n_duett_layers=cfg.n_layers, d_feedforward=cfg.d_feedforward)

    # ---- G1: student KD step (encode + StudentModel + StudentKDLoss), engine.train_student_batch semantics --------
    torch.manual_seed(0)
    duett = DuettFeatureExtractor(pretrain=False, **kw)
    student = StudentModel(duett, pool="mean", head_hidden=16, head_dropout=0.0)
    student.train()
    batch = O.synth_batch(cfg, B=6, seed=1234)
    blobs = params_of(student)          # snapshot BEFORE the forward (BN running stats change in train mode)
    z_t = torch.randn(6, generator=torch.Generator().manual_seed(7)) * 1.5
    tokens = duett.encode(duett.feats_to_input((batch["x_ts"], batch["x_static"], list(batch["bin_ends"])), 6))
    student.load_state_dict({k[len("param/"):]: v for k, v in blobs.items()})   # rewind BN running stats
    z_s = student(batch["x_ts"], batch["x_static"], list(batch["bin_ends"]))
    losses = StudentKDLoss(kd_T=4.0, kd_alpha=0.5, pos_weight=2.0)(z_s, z_t, batch["y"])
    student.zero_grad()
    losses["total"].backward()
    blobs.update(grads_of(student))
    blobs.update({"after/" + k: v.clone() for k, v in student.state_dict().items() if "running" in k or "tracked" in k})
    blobs.update({"in/x_ts": torch.stack(batch["x_ts"]), "in/x_static": torch.stack(batch["x_static"]),
                  "in/bin_ends": torch.stack(batch["bin_ends"]), "in/y": batch["y"], "in/z_t": z_t,
                  "out/tokens": tokens, "out/z_s": z_s, "out/total": losses["total"], "out/bce": losses["bce"],
                  "out/kd": losses["kd"]})
    blobs.update(self_dev(student, blobs, lambda: StudentKDLoss(kd_T=4.0, kd_alpha=0.5, pos_weight=2.0)(
        student(batch["x_ts"], batch["x_static"], list(batch["bin_ends"])), z_t, batch["y"])["total"]))
    save("g1_student_kd", blobs)

    # ---- G2: supervised Lightning step (Model.training_step, pretrain=False, rep_token, pos_frac) -----------------
    torch.manual_seed(1)
    model = Model(pretrain=False, fusion_method="rep_token", pos_frac=0.3, **kw)
    model.train()
    batch = O.synth_batch(cfg, B=6, seed=1235)
    # ragged lengths exercise feats_to_input padding (duett/duett.py:177-181)
    lens = [4, 3, 4, 2, 4, 4]
    x_ts = tuple(t[:n] for t, n in zip(batch["x_ts"], lens))
    times = [t[:n] for t, n in zip(batch["bin_ends"], lens)]
    blobs = params_of(model)
    x = (x_ts, batch["x_static"], times)
    xin = model.feats_to_input(x, 6)
    y_hat = model.forward(tuple(t.clone() if torch.is_tensor(t) else t for t in xin))
    model2_state = {k: v.clone() for k, v in model.state_dict().items()}
    # the actual training_step (runs its own forward: BN running stats advance twice overall; grads come from this one)
    model.load_state_dict({k[len("param/"):]: v for k, v in blobs.items()})
    loss = model.training_step(((x_ts, batch["x_static"], times), tuple(batch["y"].tolist())), 0)
    model.zero_grad()
    loss.backward()
    blobs.update(grads_of(model))
    blobs.update({"in/xs_static": xin[0], "in/xs_ts": xin[1], "in/xs_times": xin[2],
                  "in/n_timesteps": np.array(xin[3]), "in/y": batch["y"], "out/y_hat": y_hat, "out/loss": loss})
    blobs.update(self_dev(model, blobs, lambda: model.training_step(
        ((tuple(t.clone() for t in x_ts), batch["x_static"], [t.clone() for t in times]), tuple(batch["y"].tolist())), 0)))
    save("g2_supervised", blobs)

    # ---- G3: SSL step (pretrain_prep_batch with numpy RNG seed 42 + forward(pretrain=True) + loss) ----------------
    torch.manual_seed(2)
    model = Model(pretrain=True, seed=42, **kw)
    model.train()
    batch = O.synth_batch(cfg, B=6, seed=1236, density=0.5)
    blobs = params_of(model)
    x = (batch["x_ts"], batch["x_static"], list(batch["bin_ends"]))
    x_pre, y, mask, y_events, y_events_mask = model.pretrain_prep_batch(x, 6)
    model.rng = np.random.default_rng(42)   # rewind so training_step draws the same masks
    outs = model.forward(tuple(t.clone() if torch.is_tensor(t) else t for t in x_pre), pretrain=True)
    model.load_state_dict({k[len("param/"):]: v for k, v in blobs.items()})
    loss = model.training_step((x, tuple(batch["y"].tolist())), 0)
    model.zero_grad()
    loss.backward()
    blobs.update(grads_of(model))
    blobs.update({"in/x_ts": torch.stack(batch["x_ts"]), "in/x_static": torch.stack(batch["x_static"]),
                  "in/bin_ends": torch.stack(batch["bin_ends"]), "out/xs_ts_clipped": x_pre[1], "out/y": y,
                  "out/mask": mask, "out/y_events": y_events, "out/y_events_mask": y_events_mask,
                  "out/y_hat_value": outs[0], "out/y_hat_presence": outs[1], "out/y_hat_events": outs[2],
                  "out/y_hat_events_presence": outs[3], "out/loss": loss})
    def _ssl_step():
        model.rng = np.random.default_rng(42)
        return model.training_step((x, tuple(batch["y"].tolist())), 0)
    blobs.update(self_dev(model, blobs, _ssl_step))
    save("g3_ssl", blobs)

    # ---- G4: teacher patch_dual step (TeacherModel + PatchDualPathologyPerceiver + DualPathologyLoss + aux KL) ------
    torch.manual_seed(3)
    duett = DuettFeatureExtractor(pretrain=False, **kw)
    K, d_lat, d_img, n_patch = 7, 32, 16, 10

    class StubCXR(torch.nn.Module):      # CXR embeddings ride in the pixel_values slot
        d_out = d_img
        def forward(self, pv):
            return pv[:, 0], pv[:, 1:]

    perceiver = PatchDualPathologyPerceiver(K, duett.d_representation, d_latent=d_lat, n_heads=4, dropout=0.0,
                                            head_hidden=16, head_dropout=0.0)
    with torch.no_grad():
        perceiver.correction_head[-1].weight.normal_(0, 0.05)   # zero-init would hide the correction path
    teacher = TeacherModel(duett, StubCXR(), perceiver, patch_dual_pathology_mode=True, d_img=d_img)
    teacher.train()
    batch = O.synth_batch(cfg, B=6, seed=1237)
    g = torch.Generator().manual_seed(11)
    pv = torch.randn(6, 1 + n_patch, d_img, generator=g)
    y_multi = (torch.rand(6, K, generator=g) < 0.2).float()
    y_mask = (torch.rand(6, K, generator=g) < 0.9).float()
    lw = torch.tensor([1.0, 0.2, 0.2, 0.2, 0.2, 0.2, 0.2])
    pw = torch.tensor([2.0, 1.5, 1.0, 3.0, 1.0, 2.5, 1.2])
    blobs = params_of(teacher)
    out = teacher(batch["x_ts"], batch["x_static"], list(batch["bin_ends"]), pv)
    loss_fn = DualPathologyLoss(lw, pw, alpha_img=0.5, alpha_ts=0.5, alpha_fus=1.0)
    losses = loss_fn(out["img_logits"], out["ts_logits"], out["fusion_logits"], y_multi, y_mask)
    aux = O.aux_residual_kl(out["img_logits"], out["scaled_correction"], y_multi, y_mask)
    # cross-check the oracle's aux KL against engine.train_teacher_dual_pathology_batch (which also steps an optimizer)
    teacher.load_state_dict({k[len("param/"):]: v for k, v in blobs.items()})
    opt = torch.optim.SGD(teacher.parameters(), lr=0.0)
    eb = {"x_ts": batch["x_ts"], "x_static": batch["x_static"], "bin_ends": batch["bin_ends"], "y": batch["y"],
          "pixel_values": pv, "y_multi": y_multi, "y_multi_mask": y_mask}
    res = engine.train_teacher_dual_pathology_batch(eb, teacher, loss_fn, opt, torch.device("cpu"),
                                                    aux_residual_alpha=0.3)
    blobs.update(grads_of(teacher))      # grads of total + 0.3*aux, left in .grad by the engine step (lr = 0)
    blobs.update({"in/x_ts": torch.stack(batch["x_ts"]), "in/x_static": torch.stack(batch["x_static"]),
                  "in/bin_ends": torch.stack(batch["bin_ends"]), "in/pixel_values": pv, "in/y_multi": y_multi,
                  "in/y_multi_mask": y_mask, "in/label_weights": lw, "in/pos_weight": pw,
                  "out/loss": torch.tensor(res["loss"]), "out/aux_residual": torch.tensor(res["aux_residual"]),
                  "out/oracle_aux": aux, "out/total_noaux": losses["total"]})
    for k in ("main_logit", "img_logits", "ts_logits", "fusion_logits", "ts_correction", "scaled_correction"):
        blobs["out/" + k] = out[k]
    for k in ("img_per", "ts_per", "fus_per", "img_total", "ts_total", "fus_total"):
        blobs["out/" + k] = losses[k]
    def _teacher_step():
        o = teacher(batch["x_ts"], batch["x_static"], list(batch["bin_ends"]), pv)
        l = loss_fn(o["img_logits"].float(), o["ts_logits"].float(), o["fusion_logits"].float(), y_multi, y_mask)
        return l["total"] + 0.3 * O.aux_residual_kl(o["img_logits"].float(), o["scaled_correction"].float(), y_multi, y_mask)
    blobs.update(self_dev(teacher, blobs, _teacher_step))
    save("g4_teacher", blobs)


if __name__ == "__main__":
    main()
