"""End-to-end parity on a real B200 (through the reference-facing nn.Module API -> C ABI -> CUDA kernels):

  * replay of the golden fixtures generated from the REFERENCE'S OWN FILES (tests/golden, oracle/make_golden.py):
    logits, every loss term and every parameter gradient — fp32 mode within 1e-3 relative (north-star bound),
    bf16 mode within 2e-2;
  * a MIMIC-shaped and a C1-shaped (BASELINE.json config 1, reduced batch) random problem against the CPU oracle;
  * AUROC on a fixed synthetic eval set, scored like training_duett/evaluator.py:10-37 (sigmoid -> sklearn).
Relative error is ||a-b||2/||b||2 per tensor (SURVEY §7 "error metric"); cancellation-dominated gradients get an
absolute floor tied to the largest gradient entry, stated in `_grad_check`.
"""
import numpy as np
import pytest
import torch

from golden_util import golden_cfg, load, rel
from oracle import duett_oracle as O

pytestmark = pytest.mark.gpu

KW = dict(d_static_num=3, d_time_series_num=5, d_target=1, d_embedding=8, masked_transform_timesteps=4, max_len=4,
          n_duett_layers=2, d_feedforward=96)
MODES = [("fp32", 1e-3), ("bf16", 2e-2)]


def dev(t):
    return t.cuda() if torch.is_tensor(t) else t


def _ref_keyed_grads(module):
    from multimodal_edema_prediction_b200 import state_keys
    from multimodal_edema_prediction_b200.duett.duett import Model
    sd = {n: (p.grad if p.grad is not None else torch.zeros_like(p)).detach().cpu() for n, p in module.named_parameters()}
    for name, m in module.named_modules():
        if isinstance(m, Model):
            state_keys.to_reference(sd, name + "." if name else "")
    return sd


def _grad_check(got, want, tol, floor=1e-1, selfdev=None, min_numel=1):
    """fp32 mode / oracle comparisons: per tensor ||got-want|| <= tol*||want|| + floor*tol*max|grad|*sqrt(numel) (the
    floor covers cancellation-dominated gradients such as a bias feeding tanh->BatchNorm or a ScaleNorm gain under a
    scale-invariant BatchNorm head, whose magnitude is 1e-3 of their siblings).  min_numel: the floor of tensors with fewer
    elements is computed as if they had that many — a single-element ScaleNorm gain is a sum over every token of the batch,
    its absolute error does not shrink with its own element count.
    bf16 mode on the tiny golden fixtures (BatchNorm over 6-24 rows is ill-conditioned): the reference's OWN bf16-autocast
    run deviates from its fp32 run by `selfdev` (global relative L2 over all gradients, recorded by oracle/make_golden.py:
    3-8.5 %); the CUDA path must stay within max(5e-2, 1.5 x that)."""
    if selfdev is not None:
        g = torch.cat([got[k].flatten().double() for k in want])
        w = torch.cat([want[k].flatten().double() for k in want])
        err = float((g - w).norm() / w.norm())
        assert err <= max(5e-2, 1.5 * float(selfdev)), (err, float(selfdev))
        return
    bad = []
    gscale = max(float(v.abs().max()) for v in want.values())
    for k, w in want.items():
        g = got.get(k)
        if g is None:
            bad.append((k, "missing")); continue
        if (g.double() - w.double()).norm() > tol * w.double().norm() + floor * tol * gscale * max(w.numel(), min_numel) ** 0.5 + 1e-7:
            bad.append((k, rel(g, w), float(w.abs().max())))
    assert not bad, (len(bad), bad[:10])


@pytest.mark.parametrize("mode,tol", MODES)
def test_golden_student_kd(mode, tol):
    from multimodal_edema_prediction_b200.loss.losses_duett import StudentKDLoss
    from multimodal_edema_prediction_b200.models.main_architecture_duett import DuettFeatureExtractor, StudentModel
    G = load("g1_student_kd")
    duett = DuettFeatureExtractor(pretrain=False, precision=mode, **KW)
    student = StudentModel(duett, pool="mean", head_hidden=16, head_dropout=0.0)
    student.load_state_dict(G["param"], strict=True)
    student.cuda().train()
    I = G["in"]
    x_ts, x_static, bin_ends = tuple(I["x_ts"]), tuple(I["x_static"]), list(I["bin_ends"])
    tokens = duett.encode(duett.feats_to_input((x_ts, x_static, bin_ends), 6))
    assert rel(tokens.float().cpu(), G["out"]["tokens"]) < tol
    student.load_state_dict(G["param"])
    z_s = student(x_ts, x_static, bin_ends)
    assert rel(z_s.cpu(), G["out"]["z_s"]) < tol * 3            # logits ~0.1 in magnitude: bf16 noise is relatively larger
    losses = StudentKDLoss(kd_T=4.0, kd_alpha=0.5, pos_weight=2.0)(z_s, dev(I["z_t"]), dev(I["y"]))
    for k in ("total", "bce", "kd"):
        assert rel(losses[k].cpu(), G["out"][k]) < tol, k
    losses["total"].backward()
    _grad_check(_ref_keyed_grads(student), G["grad"], tol, selfdev=None if mode == "fp32" else G["selfdev"]["global_grad"])
    if mode == "fp32":
        sd = student.state_dict()
        for k, want in G["after"].items():
            got = sd[k].cpu()
            assert torch.equal(got, want) if want.dtype == torch.long else (rel(got, want) < 1e-4 or (got - want).abs().max() < 1e-6), k


@pytest.mark.parametrize("mode,tol", MODES)
def test_golden_supervised_ragged(mode, tol):
    from multimodal_edema_prediction_b200.duett.duett import Model
    G = load("g2_supervised")
    model = Model(pretrain=False, fusion_method="rep_token", pos_frac=0.3, precision=mode, **KW)
    model.load_state_dict(G["param"], strict=True)
    model.cuda().train()
    I = G["in"]
    lens = I["n_timesteps"].tolist()
    x_ts = tuple(I["xs_ts"][i, :n, :-1] for i, n in enumerate(lens))
    times = [I["xs_times"][i, :n] for i, n in enumerate(lens)]
    loss = model.training_step(((x_ts, tuple(I["xs_static"]), times), tuple(I["y"].tolist())), 0)
    assert loss.dtype == torch.float64 and rel(loss.cpu(), G["out"]["loss"]) < tol
    loss.backward()
    _grad_check(_ref_keyed_grads(model), G["grad"], tol, selfdev=None if mode == "fp32" else G["selfdev"]["global_grad"])


@pytest.mark.parametrize("mode,tol", MODES)
def test_golden_ssl(mode, tol):
    from multimodal_edema_prediction_b200.duett.duett import Model
    G = load("g3_ssl")
    model = Model(pretrain=True, seed=42, precision=mode, **KW)
    model.load_state_dict(G["param"], strict=True)
    model.cuda().train()
    I = G["in"]
    x = (tuple(I["x_ts"]), tuple(I["x_static"]), list(I["bin_ends"]))
    x_pre, y, mask, y_ev, y_ev_mask = model.pretrain_prep_batch(x, 6)
    assert torch.equal(x_pre[1].cpu(), G["out"]["xs_ts_clipped"])       # index / mask work: bit-exact
    assert torch.equal(y.cpu(), G["out"]["y"]) and torch.equal(mask.cpu(), G["out"]["mask"])
    assert torch.equal(y_ev.cpu(), G["out"]["y_events"]) and torch.equal(y_ev_mask.cpu(), G["out"]["y_events_mask"])
    model.rng = np.random.default_rng(42)
    outs = model.forward(x_pre, pretrain=True)
    for got, key in zip(outs, ("y_hat_value", "y_hat_presence", "y_hat_events", "y_hat_events_presence")):
        assert rel(got.cpu(), G["out"][key]) < tol * 2, key
    model.load_state_dict(G["param"])
    model.cuda()
    loss = model.training_step((x, tuple([0.0] * 6)), 0)
    assert rel(loss.cpu(), G["out"]["loss"]) < tol
    loss.backward()
    _grad_check(_ref_keyed_grads(model), G["grad"], tol, selfdev=None if mode == "fp32" else G["selfdev"]["global_grad"])


@pytest.mark.parametrize("mode,tol", MODES)
def test_golden_ssl_two_masked_steps(mode, tol):
    """pretrain_masked_steps = 2 against the reference's own step (fixture g7, oracle/make_golden_masked_steps.py): host-RNG
    draws with replacement -> dx_ssl_mask with K = 2 (bit-exact), distinct masked rows gathered and zero-padded
    (dx_gather_vec with negative offsets), heads on B*2 rows, loss = mean of the per-step losses, every gradient."""
    from multimodal_edema_prediction_b200.duett.duett import Model
    G = load("g7_ssl_masked_steps")
    model = Model(pretrain=True, seed=42, pretrain_masked_steps=2, precision=mode, **KW)
    model.load_state_dict(G["param"], strict=True)
    model.cuda().train()
    I = G["in"]
    x = (tuple(I["x_ts"]), tuple(I["x_static"]), list(I["bin_ends"]))
    x_pre, y, mask, y_ev, y_ev_mask = model.pretrain_prep_batch(x, 6)
    assert y.shape == (6, 2, 5) and torch.equal(x_pre[1].cpu(), G["out"]["xs_ts_clipped"])       # index / mask work: bit-exact
    assert torch.equal(y.cpu(), G["out"]["y"]) and torch.equal(mask.cpu(), G["out"]["mask"])
    assert torch.equal(y_ev.cpu(), G["out"]["y_events"]) and torch.equal(y_ev_mask.cpu(), G["out"]["y_events_mask"])
    outs = model.forward(x_pre, pretrain=True)
    for got, key in zip(outs, ("y_hat_value", "y_hat_presence", "y_hat_events", "y_hat_events_presence")):
        assert got.shape == G["out"][key].shape and rel(got.cpu(), G["out"][key]) < tol * 2, key
    model.rng = np.random.default_rng(42)
    model.load_state_dict(G["param"])
    model.cuda()
    loss = model.training_step((x, tuple([0.0] * 6)), 0)
    assert rel(loss.cpu(), G["out"]["loss"]) < tol
    loss.backward()
    _grad_check(_ref_keyed_grads(model), G["grad"], tol, selfdev=None if mode == "fp32" else G["selfdev"]["global_grad"])


@pytest.mark.parametrize("mode,tol", MODES)
def test_golden_teacher_patch_dual(mode, tol):
    from multimodal_edema_prediction_b200.loss.losses_duett import DualPathologyLoss
    from multimodal_edema_prediction_b200.models.main_architecture_duett import (DuettFeatureExtractor,
                                                                                 PatchDualPathologyPerceiver, TeacherModel)
    from multimodal_edema_prediction_b200.training_duett import engine
    G = load("g4_teacher")

    class StubCXR(torch.nn.Module):
        d_out = 16
        def forward(self, pv):
            return pv[:, 0], pv[:, 1:]

    duett = DuettFeatureExtractor(pretrain=False, precision=mode, **KW)
    perceiver = PatchDualPathologyPerceiver(7, duett.d_representation, d_latent=32, n_heads=4, dropout=0.0, head_hidden=16,
                                            head_dropout=0.0)
    teacher = TeacherModel(duett, StubCXR(), perceiver, patch_dual_pathology_mode=True, d_img=16)
    teacher.load_state_dict(G["param"], strict=True)
    teacher.cuda()
    I = G["in"]
    loss_fn = DualPathologyLoss(I["label_weights"], I["pos_weight"], 0.5, 0.5, 1.0).cuda()
    opt = torch.optim.SGD(teacher.parameters(), lr=0.0)
    batch = {"x_ts": tuple(I["x_ts"]), "x_static": tuple(I["x_static"]), "bin_ends": tuple(I["bin_ends"]),
             "y": torch.zeros(6), "pixel_values": I["pixel_values"], "y_multi": I["y_multi"], "y_multi_mask": I["y_multi_mask"]}
    res = engine.train_teacher_dual_pathology_batch(batch, teacher, loss_fn, opt, torch.device("cuda"), aux_residual_alpha=0.3)
    assert abs(res["loss"] - float(G["out"]["loss"])) < tol * abs(float(G["out"]["loss"]))
    for k in ("main_logit", "img_logits", "ts_logits", "fusion_logits"):
        assert rel(res[k].cpu(), G["out"][k]) < tol * 2, k
    for k in ("img_per", "ts_per", "fus_per"):
        assert rel(res[k], G["out"][k]) < tol, k
    _grad_check(_ref_keyed_grads(teacher), G["grad"], tol, selfdev=None if mode == "fp32" else G["selfdev"]["global_grad"])


@pytest.mark.parametrize("mode,tol", MODES)
def test_golden_teacher_return_attn(mode, tol):
    """Eval-mode TeacherModel.forward(return_attn=True) against the reference's own call (fixture g6,
    oracle/make_golden_attn.py): head-averaged attention maps of the two cross-attention blocks, latent tokens, logits."""
    from multimodal_edema_prediction_b200.models.main_architecture_duett import (DuettFeatureExtractor,
                                                                                 PatchDualPathologyPerceiver, TeacherModel)
    G4, G = load("g4_teacher"), load("g6_teacher_attn")

    class StubCXR(torch.nn.Module):
        d_out = 16
        def forward(self, pv):
            return pv[:, 0], pv[:, 1:]

    duett = DuettFeatureExtractor(pretrain=False, precision=mode, **KW)
    perceiver = PatchDualPathologyPerceiver(7, duett.d_representation, d_latent=32, n_heads=4, dropout=0.1, head_hidden=16,
                                            head_dropout=0.0)
    teacher = TeacherModel(duett, StubCXR(), perceiver, patch_dual_pathology_mode=True, d_img=16)
    teacher.load_state_dict(G4["param"], strict=True)
    teacher.cuda().eval()
    I = G4["in"]
    with torch.no_grad():
        out = teacher(tuple(I["x_ts"]), tuple(I["x_static"]), tuple(I["bin_ends"]), I["pixel_values"].cuda(), return_attn=True)
    assert out["img_attn"].shape == (6, 7, 10) and out["ts_attn"].shape == (6, 7, 4)
    for k in ("img_attn", "ts_attn", "img_tokens", "ts_tokens"):
        assert rel(out[k].float().cpu(), G["out"][k]) < tol, k
    for k in ("main_logit", "img_logits", "ts_logits", "fusion_logits"):
        assert rel(out[k].float().cpu(), G["out"][k]) < tol * 2, k
    assert float((out["img_attn"].sum(-1) - 1).abs().max()) < 1e-2 and float((out["ts_attn"].sum(-1) - 1).abs().max()) < 1e-2


def _oracle_vs_cuda_student(cfg, B, mode, tol, seed=0, pool="mean"):
    from multimodal_edema_prediction_b200.models.main_architecture_duett import DuettFeatureExtractor, StudentModel
    P = O.init_params(cfg, seed=seed)
    H = O.init_student_head(cfg, seed=seed + 1)
    batch = O.synth_batch(cfg, B, seed=1234 + seed)
    x_static, xs_ts, xs_times, _ = O.feats_to_input(batch["x_ts"], batch["x_static"], batch["bin_ends"], cfg.T)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    Pl = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in P.items()}
    Hl = {k: v.clone().requires_grad_(True) for k, v in H.items()}
    tokens_ref = O.encode(Pl, cfg, x_static, xs_ts, xs_times, training=True)
    z_ref = O.student_forward(Pl, Hl, cfg, x_static, xs_ts, xs_times, pool=pool)
    z_t = torch.randn(B, generator=torch.Generator().manual_seed(5)) * 1.5
    L_ref = O.student_kd_loss(z_ref, z_t, batch["y"], 4.0, 0.5, None)
    L_ref["total"].backward()
    duett = DuettFeatureExtractor(cfg.d_static_num, cfg.V, 1, d_embedding=cfg.d_embedding, n_duett_layers=cfg.n_layers,
                                  masked_transform_timesteps=cfg.T, max_len=cfg.T, d_feedforward=cfg.d_feedforward,
                                  pretrain=False, precision=mode)
    student = StudentModel(duett, pool=pool, head_hidden=128, head_dropout=0.0)
    sd = {"duett." + k: v for k, v in P.items()}
    sd.update({k: v for k, v in H.items()})
    student.load_state_dict(sd, strict=True)
    student.cuda().train()
    from multimodal_edema_prediction_b200.loss.losses_duett import StudentKDLoss
    x = (batch["x_ts"], batch["x_static"], list(batch["bin_ends"]))
    tokens = duett.encode(duett.feats_to_input(x, B))
    assert rel(tokens.float().cpu(), tokens_ref) < tol
    student.load_state_dict(sd)
    student.cuda()
    z = student(*x)
    losses = StudentKDLoss(kd_T=4.0, kd_alpha=0.5)(z, z_t.cuda(), batch["y"].cuda())
    # logits: the head's BatchNorm over B=4..16 samples amplifies the bf16 rounding of the pooled token ~10x (tokens above
    # agree to 5e-3) and the fp32 atomics of the row reductions make it vary run to run: 200 runs of the C1 case gave
    # 0.030 / 0.048 / 0.066 (min / median / max, tests/flake_check.py with FLAKE_ORACLE=1), so bf16 gets 6 x tol.
    assert rel(z.cpu(), z_ref) < tol * (3 if mode == "fp32" else 6)
    assert rel(losses["total"].cpu(), L_ref["total"]) < tol
    losses["total"].backward()
    want = {"duett." + k: v.grad for k, v in Pl.items() if torch.is_tensor(v) and v.requires_grad and v.grad is not None}
    want.update({k: v.grad for k, v in Hl.items()})
    got = _ref_keyed_grads(student)
    _grad_check({k: got[k] for k in want}, want, tol * (1 if mode == "fp32" else 2.5), floor=5e-2)


@pytest.mark.parametrize("mode,tol", MODES)
def test_mimic_default_shape_vs_oracle(mode, tol):
    """The reference's real working point: S=24, V=34, T=24, d=24, L=2, F=512 (E=600, E'=840: exercises K tails,
    non-multiple-of-64 widths and the padded time-embedding hidden (h=28))."""
    cfg = O.DuettConfig(d_static_num=24, d_time_series_num=34, n_timesteps=24, d_embedding=24, n_layers=2)
    _oracle_vs_cuda_student(cfg, B=16, mode=mode, tol=tol)


@pytest.mark.parametrize("mode,tol", MODES)
def test_c1_shape_vs_oracle(mode, tol):
    """BASELINE.json config 1 model (d=64, 2+2 layers, T=32, V=128) at a reduced batch so the CPU oracle takes seconds."""
    cfg = O.DuettConfig(d_static_num=24, d_time_series_num=128, n_timesteps=32, d_embedding=64, n_layers=2)
    _oracle_vs_cuda_student(cfg, B=4, mode=mode, tol=tol, seed=3, pool="rep_token")


@pytest.mark.parametrize("mode,tol", MODES)
def test_c1_shape_ssl_step_vs_oracle(mode, tol):
    """BASELINE.json config 3's workload (self-supervised pre-training: masked-step + masked-event reconstruction) on the
    config-1 model shape at a reduced batch: Model.training_step(pretrain=True) incl. the host numpy-RNG masking of
    pretrain_prep_batch (same seed -> bit-identical masks), loss and every parameter gradient against the CPU oracle."""
    from multimodal_edema_prediction_b200.duett.duett import Model
    cfg = O.DuettConfig(d_static_num=24, d_time_series_num=128, n_timesteps=32, d_embedding=64, n_layers=2)
    B = 6
    P = O.init_params(cfg, seed=11)
    batch = O.synth_batch(cfg, B, seed=4321)
    xs_static, xs_ts, xs_times, n_ts = O.feats_to_input(batch["x_ts"], batch["x_static"], batch["bin_ends"], cfg.T)
    x_m, y, mask, y_ev, y_ev_mask = O.pretrain_prep_batch(np.random.default_rng(42), cfg, xs_ts, n_ts, pretrain_dropout=0.5)
    Pl = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in P.items()}
    outs = O.model_forward_pretrain(Pl, cfg, xs_static, x_m, xs_times, training=True)
    L_ref = O.ssl_loss(*outs, y, mask, y_ev, y_ev_mask)
    L_ref.backward()
    model = Model(cfg.d_static_num, cfg.V, 1, d_embedding=cfg.d_embedding, n_duett_layers=cfg.n_layers,
                  masked_transform_timesteps=cfg.T, max_len=cfg.T, d_feedforward=cfg.d_feedforward, pretrain=True, seed=42,
                  precision=mode)
    model.load_state_dict(P, strict=True)
    model.cuda().train()
    x = (batch["x_ts"], batch["x_static"], list(batch["bin_ends"]))
    x_pre, y2, mask2, y_ev2, y_ev_mask2 = model.pretrain_prep_batch(x, B)
    assert torch.equal(x_pre[1].cpu(), x_m) and torch.equal(y2.cpu(), y) and torch.equal(mask2.cpu(), mask)
    assert torch.equal(y_ev2.cpu(), y_ev) and torch.equal(y_ev_mask2.cpu(), y_ev_mask)
    model.rng = np.random.default_rng(42)
    loss = model.training_step((x, tuple([0.0] * B)), 0)
    assert rel(loss.cpu(), L_ref) < tol
    loss.backward()
    want = {k: v.grad for k, v in Pl.items() if torch.is_tensor(v) and v.requires_grad and v.grad is not None}
    got = _ref_keyed_grads(model)
    _grad_check({k: got[k] for k in want}, want, tol * (1 if mode == "fp32" else 2.5), floor=5e-2)


@pytest.mark.parametrize("mode,dauc", [("fp32", 1e-4), ("bf16", 5e-3)])
def test_auroc_on_fixed_synthetic_eval_set(mode, dauc):
    """evaluate_binary semantics (training_duett/evaluator.py:10-37): eval-mode logits -> sigmoid -> sklearn AUROC, one
    process scoring the whole fixed set (seed 999)."""
    from sklearn.metrics import roc_auc_score
    from multimodal_edema_prediction_b200.models.main_architecture_duett import DuettFeatureExtractor, StudentModel
    cfg = O.DuettConfig(d_static_num=24, d_time_series_num=34, n_timesteps=24, d_embedding=24, n_layers=2)
    P, H = O.init_params(cfg, seed=7), O.init_student_head(cfg, seed=8)
    H["head.3.weight"] = H["head.3.weight"] * 40.0       # O(1) logits so ranking is meaningful
    for k in P:                                           # non-trivial running statistics for eval-mode BatchNorm
        if k.endswith("running_mean"):
            P[k] = 0.05 * torch.randn(P[k].shape, generator=torch.Generator().manual_seed(1))
        if k.endswith("running_var"):
            P[k] = 0.5 + torch.rand(P[k].shape, generator=torch.Generator().manual_seed(2))
    N, bs = 1024, 256
    data = O.synth_batch(cfg, N, seed=999)
    duett = DuettFeatureExtractor(24, 34, 1, d_embedding=24, n_duett_layers=2, masked_transform_timesteps=24, max_len=24,
                                  pretrain=False, precision=mode)
    student = StudentModel(duett, pool="mean", head_dropout=0.0)
    sd = {"duett." + k: v for k, v in P.items()}
    sd.update(H)
    student.load_state_dict(sd, strict=True)
    student.cuda().eval()
    zs, zr = [], []
    with torch.no_grad():
        for i in range(0, N, bs):
            sl = slice(i, i + bs)
            x = (data["x_ts"][sl], data["x_static"][sl], list(data["bin_ends"][sl]))
            zs.append(student(*x).float().cpu())
            xs, xt, tm, _ = O.feats_to_input(*x, cfg.T)
            zr.append(O.student_forward(P, H, cfg, xs, xt, tm, pool="mean", training=False))
    zs, zr = torch.cat(zs), torch.cat(zr)
    y = (zr + torch.randn(N, generator=torch.Generator().manual_seed(3)) * zr.std() > zr.median()).float().numpy()
    a_cuda = roc_auc_score(y, torch.sigmoid(zs).numpy())
    a_ref = roc_auc_score(y, torch.sigmoid(zr).numpy())
    assert abs(a_cuda - a_ref) < dauc, (a_cuda, a_ref)
    # the same set through the device-side evaluator (evaluate_binary surface of training_duett/evaluator.py:10-37)
    from multimodal_edema_prediction_b200.training_duett import evaluator
    loader = [{"x_ts": data["x_ts"][i:i + bs], "x_static": data["x_static"][i:i + bs], "bin_ends": data["bin_ends"][i:i + bs],
               "y": torch.from_numpy(y[i:i + bs])} for i in range(0, N, bs)]
    res = evaluator.evaluate_binary(student, loader, torch.device("cuda"), evaluator.make_student_forward())
    # (a second forward pass of the same model: its logits repeat only to ~1e-6 - atomic row sums - so near-ties may flip;
    # one pair of the 492 x 532 is 3.8e-6 of AUROC.  The bound is the same parity bound as above, against the oracle.)
    assert res["n"] == N and abs(res["auroc"] - a_ref) < dauc, (res, a_ref, a_cuda)
    # logits: fp32 within the north-star 1e-3.  In bf16 the default-init logits are ~0.1 with a large common-mode part, so
    # ||dz||/||z|| is dominated by cancellation (the reference's own bf16 run is off by 37 % on this measure);
    # compare the sample-to-sample variation that AUROC actually ranks on.
    if mode == "fp32":
        assert rel(zs, zr) < 1e-3
    else:
        assert float(((zs - zs.mean()) - (zr - zr.mean())).norm() / (zr - zr.mean()).norm()) < 0.1


@pytest.mark.parametrize("mode,tol", MODES)
def test_student_step_with_dropout_vs_oracle(mode, tol):
    """transformer_dropout (attention probabilities + FFN hidden of every axis encoder) and head dropout in training mode:
    the product's dropout sites log their (probability, seed); the oracle replays the step with masks drawn from the same
    counter-based generator, so logits, loss and every gradient are compared on identical masks."""
    from multimodal_edema_prediction_b200 import ops
    from multimodal_edema_prediction_b200.loss.losses_duett import StudentKDLoss
    from multimodal_edema_prediction_b200.models.main_architecture_duett import DuettFeatureExtractor, StudentModel
    cfg = O.DuettConfig(d_static_num=24, d_time_series_num=34, n_timesteps=24, d_embedding=32, n_layers=2)
    B = 8
    P, H = O.init_params(cfg, seed=21), O.init_student_head(cfg, seed=22)
    batch = O.synth_batch(cfg, B, seed=777)
    duett = DuettFeatureExtractor(cfg.d_static_num, cfg.V, 1, d_embedding=cfg.d_embedding, n_duett_layers=cfg.n_layers,
                                  masked_transform_timesteps=cfg.T, max_len=cfg.T, d_feedforward=cfg.d_feedforward,
                                  pretrain=False, precision=mode, transformer_dropout=0.25)
    student = StudentModel(duett, pool="mean", head_hidden=128, head_dropout=0.1)
    sd = {"duett." + k: v for k, v in P.items()}
    sd.update(H)
    student.load_state_dict(sd, strict=True)
    student.cuda().train()
    x = (batch["x_ts"], batch["x_static"], list(batch["bin_ends"]))
    z_t = torch.randn(B, generator=torch.Generator().manual_seed(5)) * 1.5
    ops.reset_drop_seeds(1234)
    ops.DROP_LOG = []
    try:
        z = student(*x)
        log = list(ops.DROP_LOG)
    finally:
        ops.DROP_LOG = None
    drop = {("head" if tag == "dropout" else tag): (p, seed) for tag, p, seed, _ in log}
    assert len(drop) == 4 * cfg.n_layers + 1 and all(p > 0 for p, _ in drop.values()), sorted(drop)
    losses = StudentKDLoss(kd_T=4.0, kd_alpha=0.5)(z, z_t.cuda(), batch["y"].cuda())
    losses["total"].backward()
    xs_static, xs_ts, xs_times, _ = O.feats_to_input(batch["x_ts"], batch["x_static"], batch["bin_ends"], cfg.T)
    Pl = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in P.items()}
    Hl = {k: v.clone().requires_grad_(True) for k, v in H.items()}
    z_ref = O.student_forward(Pl, Hl, cfg, xs_static, xs_ts, xs_times, pool="mean", drop=drop)
    L_ref = O.student_kd_loss(z_ref, z_t, batch["y"], 4.0, 0.5, None)
    L_ref["total"].backward()
    assert rel(z.cpu(), z_ref) < tol * (3 if mode == "fp32" else 6)
    assert rel(losses["total"].cpu(), L_ref["total"]) < tol
    want = {"duett." + k: v.grad for k, v in Pl.items() if torch.is_tensor(v) and v.requires_grad and v.grad is not None}
    want.update({k: v.grad for k, v in Hl.items()})
    got = _ref_keyed_grads(student)
    _grad_check({k: got[k] for k in want}, want, tol * (1 if mode == "fp32" else 2.5), floor=5e-2)
    # eval mode: dropout is the identity and no site draws a seed
    student.eval()
    ops.DROP_LOG = []
    try:
        with torch.no_grad():
            student(*x)
        assert ops.DROP_LOG == []
    finally:
        ops.DROP_LOG = None
