"""Host-side logic of the B200 package on CPU: the kernel wrappers are replaced by the torch contract emulator
(tests/ops_emulator.py), everything else — fused-backbone orchestration and its ScaleNorm-backward algebra, autograd
Functions, nn.Module surface, reference state-dict key mapping, engine steps — is the product code, checked against the
golden fixtures generated from the REFERENCE'S OWN FILES."""
import numpy as np
import pytest
import torch

import ops_emulator
from golden_util import golden_cfg, load, rel

KW = dict(d_static_num=3, d_time_series_num=5, d_target=1, d_embedding=8, masked_transform_timesteps=4, max_len=4,
          n_duett_layers=2, d_feedforward=96)
TOL = 5e-5


@pytest.fixture
def emu(monkeypatch):
    ops_emulator.install(monkeypatch)


def _grad_check(named_params, ggold, prefix="", tol=3e-4):
    """named_params yield internal names; map through the module's own state_dict hook by comparing in reference keys."""
    bad = []
    gscale = max(float(v.abs().max()) for v in ggold.values())
    for k, want in ggold.items():
        got = named_params.get(k)
        if got is None:
            bad.append((k, "missing"))
            continue
        if (got.double() - want.double()).norm() > tol * want.double().norm() + 5e-5 * gscale * want.numel() ** 0.5:
            bad.append((k, rel(got, want)))
    assert not bad, bad[:8]


def _ref_keyed_grads(module):
    """Gradients keyed like the reference's named_parameters()."""
    from multimodal_edema_prediction_b200 import state_keys
    sd = {}
    for n, p in module.named_parameters():
        sd[n] = p.grad if p.grad is not None else torch.zeros_like(p)
    # find prefixes of Model instances to apply the key mapping
    from multimodal_edema_prediction_b200.duett.duett import Model
    for name, m in module.named_modules():
        if isinstance(m, Model):
            state_keys.to_reference(sd, name + "." if name else "")
    return sd


def test_student_kd_step_matches_reference(emu):
    from multimodal_edema_prediction_b200.loss.losses_duett import StudentKDLoss
    from multimodal_edema_prediction_b200.models.main_architecture_duett import DuettFeatureExtractor, StudentModel
    G = load("g1_student_kd")
    duett = DuettFeatureExtractor(pretrain=False, **KW)
    student = StudentModel(duett, pool="mean", head_hidden=16, head_dropout=0.0)
    missing, unexpected = student.load_state_dict(G["param"], strict=True)
    assert not missing and not unexpected
    # round trip: our state_dict() exposes exactly the reference's keys
    assert set(student.state_dict().keys()) == set(G["param"].keys())
    student.train()
    I = G["in"]
    x_ts, x_static, bin_ends = tuple(I["x_ts"]), tuple(I["x_static"]), list(I["bin_ends"])
    tokens = duett.encode(duett.feats_to_input((x_ts, x_static, bin_ends), 6))
    assert rel(tokens, G["out"]["tokens"]) < TOL
    student.load_state_dict(G["param"])      # rewind BN running stats
    z_s = student(x_ts, x_static, bin_ends)
    assert rel(z_s, G["out"]["z_s"]) < TOL
    losses = StudentKDLoss(kd_T=4.0, kd_alpha=0.5, pos_weight=2.0)(z_s, I["z_t"], I["y"])
    for k in ("total", "bce", "kd"):
        assert rel(losses[k], G["out"][k]) < TOL, k
    losses["total"].backward()
    _grad_check(_ref_keyed_grads(student), G["grad"])
    # BatchNorm running statistics / counters advance exactly like the reference's modules
    sd = student.state_dict()
    for k, want in G["after"].items():
        if want.dtype == torch.long:
            assert torch.equal(sd[k], want), k
        else:
            assert rel(sd[k], want) < 1e-5 or (sd[k] - want).abs().max() < 1e-6, k


def test_supervised_step_matches_reference(emu):
    from multimodal_edema_prediction_b200.duett.duett import Model
    G = load("g2_supervised")
    model = Model(pretrain=False, fusion_method="rep_token", pos_frac=0.3, **KW)
    model.load_state_dict(G["param"], strict=True)
    model.train()
    I = G["in"]
    lens = I["n_timesteps"].tolist()
    x_ts = tuple(I["xs_ts"][i, :n, :-1] for i, n in enumerate(lens))
    times = [I["xs_times"][i, :n] for i, n in enumerate(lens)]
    loss = model.training_step(((x_ts, tuple(I["xs_static"]), times), tuple(I["y"].tolist())), 0)
    assert loss.dtype == torch.float64
    assert rel(loss, G["out"]["loss"]) < TOL
    loss.backward()
    _grad_check(_ref_keyed_grads(model), G["grad"])


def test_ssl_step_matches_reference(emu):
    from multimodal_edema_prediction_b200.duett.duett import Model
    G = load("g3_ssl")
    model = Model(pretrain=True, seed=42, **KW)
    model.load_state_dict(G["param"], strict=True)
    model.train()
    I = G["in"]
    x = (tuple(I["x_ts"]), tuple(I["x_static"]), list(I["bin_ends"]))
    x_pre, y, mask, y_ev, y_ev_mask = model.pretrain_prep_batch(x, 6)
    assert torch.equal(x_pre[1], G["out"]["xs_ts_clipped"])           # host RNG masking: bit-exact
    assert torch.equal(y, G["out"]["y"]) and torch.equal(mask, G["out"]["mask"])
    assert torch.equal(y_ev, G["out"]["y_events"]) and torch.equal(y_ev_mask, G["out"]["y_events_mask"])
    model.rng = np.random.default_rng(42)
    model.load_state_dict(G["param"])
    loss = model.training_step((x, tuple([0.0] * 6)), 0)
    assert rel(loss, G["out"]["loss"]) < TOL
    loss.backward()
    _grad_check(_ref_keyed_grads(model), G["grad"])


def test_ssl_step_with_two_masked_steps_matches_reference(emu):
    """pretrain_masked_steps = 2 against the reference's own step (fixture g7, oracle/make_golden_masked_steps.py): draws with
    replacement bit-exact, [B,2,V] targets, distinct masked rows zero-padded to two, loss and every gradient."""
    from multimodal_edema_prediction_b200.duett.duett import Model
    G = load("g7_ssl_masked_steps")
    model = Model(pretrain=True, seed=42, pretrain_masked_steps=2, **KW)
    model.load_state_dict(G["param"], strict=True)
    model.train()
    I = G["in"]
    x = (tuple(I["x_ts"]), tuple(I["x_static"]), list(I["bin_ends"]))
    x_pre, y, mask, y_ev, y_ev_mask = model.pretrain_prep_batch(x, 6)
    assert y.shape == (6, 2, 5) and torch.equal(x_pre[1], G["out"]["xs_ts_clipped"])
    assert torch.equal(y, G["out"]["y"]) and torch.equal(mask, G["out"]["mask"])
    assert torch.equal(y_ev, G["out"]["y_events"]) and torch.equal(y_ev_mask, G["out"]["y_events_mask"])
    outs = model.forward(x_pre, pretrain=True)
    for got, key in zip(outs, ("y_hat_value", "y_hat_presence", "y_hat_events", "y_hat_events_presence")):
        assert got.shape == G["out"][key].shape and rel(got, G["out"][key]) < TOL, key
    z = model.forward(x_pre, representation=True)
    assert z.shape == (6, 2, model.d_embedding * 6)
    dup = G["out"]["n_masked"] == 1                        # a repeated draw: the second row is the zero padding
    assert bool(dup.any()) and float(z[dup, 1].abs().max()) == 0.0 and float(z[~dup, 1].abs().min()) > 0.0
    model.rng = np.random.default_rng(42)
    model.load_state_dict(G["param"])
    loss = model.training_step((x, tuple([0.0] * 6)), 0)
    assert rel(loss, G["out"]["loss"]) < TOL
    loss.backward()
    _grad_check(_ref_keyed_grads(model), G["grad"])
    short = (tuple(t[:n] for t, n in zip(I["x_ts"], (4, 4, 1, 4, 4, 4))), x[1], [t[:n] for t, n in zip(I["bin_ends"], (4, 4, 1, 4, 4, 4))])
    with pytest.raises(ValueError):
        model.pretrain_prep_batch(short, 6)
    with pytest.raises(ValueError):
        Model(pretrain=True, pretrain_masked_steps=0, **KW)


def test_teacher_step_matches_reference(emu):
    from multimodal_edema_prediction_b200.loss.losses_duett import DualPathologyLoss
    from multimodal_edema_prediction_b200.models.main_architecture_duett import (DuettFeatureExtractor,
                                                                                 PatchDualPathologyPerceiver, TeacherModel)
    from multimodal_edema_prediction_b200.training_duett import engine
    G = load("g4_teacher")

    class StubCXR(torch.nn.Module):
        d_out = 16
        def forward(self, pv):
            return pv[:, 0], pv[:, 1:]

    duett = DuettFeatureExtractor(pretrain=False, **KW)
    perceiver = PatchDualPathologyPerceiver(7, duett.d_representation, d_latent=32, n_heads=4, dropout=0.0, head_hidden=16,
                                            head_dropout=0.0)
    teacher = TeacherModel(duett, StubCXR(), perceiver, patch_dual_pathology_mode=True, d_img=16)
    teacher.load_state_dict(G["param"], strict=True)
    assert set(teacher.state_dict().keys()) == set(G["param"].keys())
    I = G["in"]
    loss_fn = DualPathologyLoss(I["label_weights"], I["pos_weight"], 0.5, 0.5, 1.0)
    opt = torch.optim.SGD(teacher.parameters(), lr=0.0)
    batch = {"x_ts": tuple(I["x_ts"]), "x_static": tuple(I["x_static"]), "bin_ends": tuple(I["bin_ends"]),
             "y": torch.zeros(6), "pixel_values": I["pixel_values"], "y_multi": I["y_multi"],
             "y_multi_mask": I["y_multi_mask"]}
    res = engine.train_teacher_dual_pathology_batch(batch, teacher, loss_fn, opt, torch.device("cpu"),
                                                    aux_residual_alpha=0.3)
    assert abs(res["loss"] - float(G["out"]["loss"])) < 1e-4 * abs(float(G["out"]["loss"]))
    assert abs(res["aux_residual"] - float(G["out"]["aux_residual"])) < 1e-4
    for k in ("main_logit", "img_logits", "ts_logits", "fusion_logits"):
        assert rel(res[k], G["out"][k]) < TOL, k
    for k in ("img_per", "ts_per", "fus_per"):
        assert rel(res[k], G["out"][k]) < TOL, k
    _grad_check(_ref_keyed_grads(teacher), G["grad"])


def test_teacher_return_attn_matches_reference(emu):
    """TeacherModel.forward(return_attn=True) in eval mode against the reference's own call (fixture g6,
    oracle/make_golden_attn.py): attention maps, latent tokens, logits; the extra keys appear only when asked for."""
    from multimodal_edema_prediction_b200.models.main_architecture_duett import (DuettFeatureExtractor,
                                                                                 PatchDualPathologyPerceiver, TeacherModel)
    G4, G = load("g4_teacher"), load("g6_teacher_attn")

    class StubCXR(torch.nn.Module):
        d_out = 16
        def forward(self, pv):
            return pv[:, 0], pv[:, 1:]

    duett = DuettFeatureExtractor(pretrain=False, **KW)
    perceiver = PatchDualPathologyPerceiver(7, duett.d_representation, d_latent=32, n_heads=4, dropout=0.1, head_hidden=16,
                                            head_dropout=0.0)
    teacher = TeacherModel(duett, StubCXR(), perceiver, patch_dual_pathology_mode=True, d_img=16)
    teacher.load_state_dict(G4["param"], strict=True)
    I = G4["in"]
    args = (tuple(I["x_ts"]), tuple(I["x_static"]), tuple(I["bin_ends"]), I["pixel_values"])
    teacher.eval()
    with torch.no_grad():
        out = teacher(*args, return_attn=True)
        plain = teacher(*args)
    assert set(out) - set(plain) == {"img_tokens", "ts_tokens", "fusion_tokens", "img_attn", "ts_attn"}
    for k, v in G["out"].items():
        assert out[k].shape == v.shape and rel(out[k], v) < TOL, k
    assert out["img_attn"].dtype == torch.float32
    for k in plain:
        assert torch.equal(plain[k], out[k]), k
    # ts_ablation variants change the key axis of ts_attn (T hourly tokens / T+1 with [REP] / the [REP] token alone)
    tok = torch.randn(6, 5, duett.d_representation)
    proj = torch.randn(6, 10, 32)
    for abl, n in (("hourly_only", 4), ("full", 5), ("rep_only", 1)):
        o = perceiver(tok, proj, return_attn=True, ts_ablation=abl)
        assert o["ts_attn"].shape == (6, 7, n) and o["img_attn"].shape == (6, 7, 10)
        assert torch.allclose(o["ts_attn"].sum(-1), torch.ones(6, 7), atol=1e-5)
    teacher.train()
    with pytest.raises(NotImplementedError):      # attention dropout active: torch would return the dropped maps
        teacher(*args, return_attn=True)


def test_state_dict_round_trip_and_tolerant_checkpoint_loading(emu, tmp_path):
    from multimodal_edema_prediction_b200.models.main_architecture_duett import DuettFeatureExtractor, load_duett_backbone
    G = load("g3_ssl")        # an SSL "checkpoint" in the reference's key layout
    ck = tmp_path / "ssl.ckpt"
    sd = dict(G["param"])
    sd["head.0.weight"] = torch.zeros(7, 3)          # reshaped head.* must be skipped (duett/duett.py:473-478)
    sd["stale.key"] = torch.zeros(1)                  # unknown keys are dropped
    del sd["tab_encoder.4.bias"]                      # missing keys keep their init
    torch.save({"state_dict": sd, "optimizer_states": [1]}, ck)
    m = load_duett_backbone(str(ck), 3, 5, 4, freeze=True, d_embedding=8, n_duett_layers=2, d_feedforward=96)
    assert isinstance(m, DuettFeatureExtractor) and not m.training
    assert not any(p.requires_grad for p in m.parameters())
    got = m.state_dict()
    assert torch.equal(got["embedding_layers.3.4.weight"], G["param"]["embedding_layers.3.4.weight"])
    assert torch.equal(got["time_transformers.1.layers.0.1.to_k.weight"], G["param"]["time_transformers.1.layers.0.1.to_k.weight"])
    assert got["head.0.weight"].shape == (64, 48)
    assert m.d_representation == 48


def test_evaluate_binary_surface(emu):
    """evaluate_binary(model, loader, device, forward_fn) -> {"auroc","auprc","n","pos_frac"} like the reference's
    evaluator (training_duett/evaluator.py:10-37); scoring itself is the kernel's job (emulated here by its contract)."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    from multimodal_edema_prediction_b200.training_duett import evaluator

    class Dummy(torch.nn.Module):
        def forward(self, x):
            return x * 2.0 - 0.3

    g = torch.Generator().manual_seed(0)
    batches = [{"x": torch.randn(17, generator=g), "y": (torch.rand(17, generator=g) < 0.4).float()} for _ in range(5)]
    res = evaluator.evaluate_binary(Dummy(), batches, torch.device("cpu"), lambda m, b, dev: {"logits": m(b["x"]), "y": b["y"]})
    z = torch.cat([b["x"] * 2.0 - 0.3 for b in batches])
    y = torch.cat([b["y"] for b in batches]).numpy()
    assert res["n"] == 85 and abs(res["pos_frac"] - y.mean()) < 1e-12
    assert abs(res["auroc"] - roc_auc_score(y, torch.sigmoid(z).numpy())) < 1e-12
    assert abs(res["auprc"] - average_precision_score(y, torch.sigmoid(z).numpy())) < 1e-12
    assert set(res) == {"auroc", "auprc", "n", "pos_frac"}
    empty = evaluator.evaluate_binary(Dummy(), [], torch.device("cpu"), lambda m, b, dev: b)       # empty loader: NaNs, no launch
    assert empty["n"] == 0 and empty["auroc"] != empty["auroc"] and set(empty) == set(res)
    for name in ("make_teacher_forward", "make_teacher_aux_forward", "make_student_forward"):
        assert callable(getattr(evaluator, name)())


def test_student_step_with_dropout_host_logic(emu):
    """Dropout orchestration of the product (sites and seeds, in-place drop of the saved FFN hidden, mask applied to the
    GELU' outputs, row-dot recomputed after it, attention-probability mask in forward and backward) against the oracle
    replaying the same step with masks from the same generator (seeds read back from the product's dropout log)."""
    from oracle import duett_oracle as O
    from multimodal_edema_prediction_b200 import ops
    from multimodal_edema_prediction_b200.loss.losses_duett import StudentKDLoss
    from multimodal_edema_prediction_b200.models.main_architecture_duett import DuettFeatureExtractor, StudentModel
    cfg = O.DuettConfig(d_static_num=3, d_time_series_num=5, n_timesteps=4, d_embedding=8, n_layers=2, d_feedforward=96)
    B = 6
    P, H = O.init_params(cfg, seed=21), O.init_student_head(cfg, seed=22, head_hidden=16)
    batch = O.synth_batch(cfg, B, seed=777)
    duett = DuettFeatureExtractor(cfg.d_static_num, cfg.V, 1, d_embedding=cfg.d_embedding, n_duett_layers=cfg.n_layers,
                                  masked_transform_timesteps=cfg.T, max_len=cfg.T, d_feedforward=cfg.d_feedforward,
                                  pretrain=False, precision="fp32", transformer_dropout=0.25)
    student = StudentModel(duett, pool="mean", head_hidden=16, head_dropout=0.1)
    sd = {"duett." + k: v for k, v in P.items()}
    sd.update(H)
    student.load_state_dict(sd, strict=True)
    student.train()
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
    assert sorted(drop) == sorted([f"{k}_transformers.{l}.{s}" for k in ("event", "time") for l in range(2) for s in ("attn", "ff")]
                                  + ["head"])
    StudentKDLoss(kd_T=4.0, kd_alpha=0.5)(z, z_t, batch["y"])["total"].backward()
    xs_static, xs_ts, xs_times, _ = O.feats_to_input(batch["x_ts"], batch["x_static"], batch["bin_ends"], cfg.T)
    Pl = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in P.items()}
    Hl = {k: v.clone().requires_grad_(True) for k, v in H.items()}
    z_ref = O.student_forward(Pl, Hl, cfg, xs_static, xs_ts, xs_times, pool="mean", drop=drop)
    O.student_kd_loss(z_ref, z_t, batch["y"], 4.0, 0.5, None)["total"].backward()
    assert rel(z, z_ref) < TOL
    want = {"duett." + k: v.grad for k, v in Pl.items() if torch.is_tensor(v) and v.requires_grad and v.grad is not None}
    want.update({k: v.grad for k, v in Hl.items()})
    _grad_check(_ref_keyed_grads(student), want)
    # a second training forward draws new seeds (fresh masks); eval mode draws none
    ops.DROP_LOG = []
    try:
        student(*x)
        assert len(ops.DROP_LOG) == 9 and {s for _, _, s, _ in ops.DROP_LOG}.isdisjoint({s for _, _, s, _ in log})
        student.eval()
        ops.DROP_LOG = []
        with torch.no_grad():
            student(*x)
        assert ops.DROP_LOG == []
    finally:
        ops.DROP_LOG = None


def test_evaluate_dual_pathology_surface(emu):
    """evaluate_dual_pathology: result keys and values against the reference's formulas restated with numpy / sklearn
    (training_duett/evaluator.py:197-335) on a stub teacher; the ranking kernel is emulated by its contract."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    from multimodal_edema_prediction_b200.training_duett import evaluator
    K, labels = 3, ("a", "b", "c")
    g = torch.Generator().manual_seed(3)

    class Perc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.beta = torch.nn.Parameter(torch.tensor([0.5, 1.0, 1.5]))

    class Teacher(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.perceiver = Perc()

        def forward(self, x_ts, x_static, bin_ends, pv):
            s = torch.stack([t.sum() for t in x_ts])[:, None]
            img = pv[:, :K] + 0.1 * s
            corr = 0.3 * pv[:, K:2 * K]
            return {"img_logits": img, "ts_logits": pv[:, 2 * K:3 * K], "fusion_logits": img + corr, "scaled_correction": corr,
                    "main_logit": img[:, 0]}

    batches = []
    for _ in range(4):
        n = 9
        batches.append({"x_ts": tuple(torch.randn(2, 3, generator=g) for _ in range(n)),
                        "x_static": tuple(torch.zeros(1) for _ in range(n)), "bin_ends": tuple(torch.zeros(1) for _ in range(n)),
                        "y": torch.zeros(n), "pixel_values": torch.randn(n, 3 * K, generator=g),
                        "y_multi": (torch.rand(n, K, generator=g) < 0.4).float(),
                        "y_multi_mask": (torch.rand(n, K, generator=g) < 0.8).float()})
    model = Teacher()
    res = evaluator.evaluate_dual_pathology(model, batches, torch.device("cpu"), labels)
    outs = [model(b["x_ts"], b["x_static"], b["bin_ends"], b["pixel_values"]) for b in batches]
    cat = lambda k: torch.cat([o[k] for o in outs]).detach().numpy()
    img, fus, corr = cat("img_logits"), cat("fusion_logits"), cat("scaled_correction")
    y = torch.cat([b["y_multi"] for b in batches]).numpy()
    mk = torch.cat([b["y_multi_mask"] for b in batches]).numpy().astype(bool)
    assert res["labels"] == list(labels) and res["n"] == 36 and len(res["per_label"]) == K
    fa = []
    for k in range(K):
        m = mk[:, k]
        yk, li, lf, ck = y[m, k], img[m, k], fus[m, k], corr[m, k]
        r = res["per_label"][k]
        sig = lambda z: 1.0 / (1.0 + np.exp(-z))
        assert r["n_valid"] == int(m.sum()) and abs(r["pos_frac"] - yk.mean()) < 1e-12
        assert abs(r["img_auroc"] - roc_auc_score(yk, sig(li))) < 1e-6 and abs(r["fus_auprc"] - average_precision_score(yk, sig(lf))) < 1e-6
        bce = lambda l: float((np.maximum(l, 0) - l * yk + np.log1p(np.exp(-np.abs(l)))).mean())
        assert abs(r["delta_bce"] - (bce(lf) - bce(li))) < 1e-6 and abs(r["mean_abs_corr"] - np.abs(ck).mean()) < 1e-6
        assert abs(r["corr_residual"] - np.corrcoef(ck, yk - sig(li))[0, 1]) < 1e-5
        assert abs(r["gap_i2f"] - (r["fus_auroc"] - r["img_auroc"])) < 1e-12 and abs(r["beta"] - [0.5, 1.0, 1.5][k]) < 1e-7
        fa.append(r["fus_auroc"])
    assert abs(res["main_auroc"] - sum(fa) / K) < 1e-12


def test_upload_staging_is_bounded_by_the_largest_batch():
    """Model._upload's pinned staging (duett._Staging, exercised here unpinned): two flat buffers per (slot, dtype) sized
    for the largest batch seen, viewed at the requested shape; alternating buffers wait for the event of their last copy."""
    from multimodal_edema_prediction_b200.duett.duett import _Staging

    class Ev:
        waited = 0

        def synchronize(self):
            Ev.waited += 1

    st = _Staging(pin=False)
    ptrs = []
    for k, shape in enumerate([(4, 3, 5), (4, 2, 5), (2, 3, 5), (4, 3, 5), (6, 3, 5), (1, 1, 5)]):
        buf, release = st.acquire("ts", shape, torch.float32)
        assert tuple(buf.shape) == shape and buf.is_contiguous()
        src = [torch.full(shape[1:], float(k + j)) for j in range(shape[0])]
        torch.stack(src, out=buf)
        assert torch.equal(buf, torch.stack(src))
        ptrs.append(buf.data_ptr())
        release(Ev())
    # double buffering; a smaller batch reuses the allocation (2, 5), a larger one replaces it (3: 40 -> 60 elements, 4: 60 -> 90)
    assert ptrs[0] != ptrs[1] and ptrs[2] == ptrs[0] and ptrs[3] != ptrs[1] and ptrs[5] == ptrs[3]
    assert Ev.waited == 4                                                        # every reuse waited for the previous copy
    assert st.pinned_bytes() == (6 * 3 * 5 + 4 * 3 * 5) * 4                      # the largest batch each buffer has held
    buf64, _ = st.acquire("ts", (2, 2), torch.float64)                           # another dtype: its own pair
    assert buf64.dtype == torch.float64 and len(st.slots) == 2
    import copy, pickle                                                          # copies of a Model start with empty staging
    assert copy.deepcopy(st).slots == {} and pickle.loads(pickle.dumps(st)).slots == {} and copy.deepcopy(st).pin is False

