"""Pins oracle/duett_oracle.py against fixtures produced by the REFERENCE'S OWN FILES (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from golden_util import golden_cfg, load, rel
from oracle import duett_oracle as O

TOL = 2e-5   # fp32 re-association only (batched einsum vs per-variable loop)


def _leaf(P):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
            for k, v in P.items()}


def _check_grads(P, grads, prefix="", tol=2e-4):
    bad = []
    gscale = max(float(g.abs().max()) for k, g in grads.items() if k.startswith(prefix))
    for k, g in grads.items():
        if not k.startswith(prefix):
            continue
        kk = k[len(prefix):]
        if kk not in P or not P[kk].requires_grad:
            continue
        got = P[kk].grad if P[kk].grad is not None else torch.zeros_like(P[kk])
        if g.abs().max() < 1e-12 and got.abs().max() < 1e-9:
            continue
        # ill-conditioned grads (a bias feeding tanh->BatchNorm, a ScaleNorm gain whose terms cancel: magnitudes
        # 1e-3 of their siblings) get an absolute floor of 1e-5 x the largest gradient entry — fp32 re-association noise
        if (got.double() - g.double()).norm() > tol * g.double().norm() + 1e-5 * gscale * g.numel() ** 0.5:
            bad.append((kk, rel(got, g)))
    assert not bad, bad[:10]


def test_student_kd_matches_reference():
    G, cfg = load("g1_student_kd"), golden_cfg()
    P = _leaf({k[len("duett."):]: v for k, v in G["param"].items() if k.startswith("duett.")})
    H = _leaf({k: v for k, v in G["param"].items() if k.startswith("head.")})
    x_static, xs_ts, xs_times, _ = O.feats_to_input(G["in"]["x_ts"], G["in"]["x_static"], G["in"]["bin_ends"], cfg.T)
    stats = {}
    tokens = O.encode(P, cfg, x_static, xs_ts, xs_times, training=True, stats_out=stats)
    assert rel(tokens, G["out"]["tokens"]) < TOL
    z_s = O.student_forward(P, H, cfg, x_static, xs_ts, xs_times, pool="mean")
    assert rel(z_s, G["out"]["z_s"]) < TOL
    L = O.student_kd_loss(z_s, G["in"]["z_t"], G["in"]["y"], 4.0, 0.5, 2.0)
    for k in ("total", "bce", "kd"):
        assert rel(L[k], G["out"][k]) < TOL, k
    L["total"].backward()
    _check_grads(P, G["grad"], "duett.")
    _check_grads(H, G["grad"], "")
    # BatchNorm running statistics after ONE training forward (momentum 0.1, unbiased running var)
    for prefix, (mean, var_unb) in stats.items():
        rm = 0.9 * G["param"]["duett." + prefix + ".batch_norm.running_mean"] + 0.1 * mean
        rv = 0.9 * G["param"]["duett." + prefix + ".batch_norm.running_var"] + 0.1 * var_unb
        assert rel(rm, G["after"]["duett." + prefix + ".batch_norm.running_mean"]) < 1e-5 or rm.abs().max() < 1e-6
        assert rel(rv, G["after"]["duett." + prefix + ".batch_norm.running_var"]) < 1e-5


def test_supervised_step_matches_reference_with_ragged_batch():
    G, cfg = load("g2_supervised"), golden_cfg()
    P = _leaf(G["param"])
    I = G["in"]
    y_hat = O.model_forward_supervised(P, cfg, I["xs_static"], I["xs_ts"], I["xs_times"], "rep_token")
    assert rel(y_hat, G["out"]["y_hat"]) < TOL
    loss = O.supervised_loss(y_hat, I["y"], pos_frac=0.3)
    assert loss.dtype == torch.float64          # reference quirk: float64 labels -> float64 loss
    assert rel(loss, G["out"]["loss"]) < TOL
    loss.backward()
    _check_grads(P, G["grad"])


def test_feats_to_input_pads_ragged():
    cfg = golden_cfg()
    b = O.synth_batch(cfg, 4, seed=5)
    lens = [4, 2, 3, 1]
    xs, st, tm, n = None, None, None, None
    st, xs, tm, n = O.feats_to_input([t[:k] for t, k in zip(b["x_ts"], lens)], b["x_static"],
                                     [t[:k] for t, k in zip(b["bin_ends"], lens)], cfg.T)
    assert xs.shape == (4, 4, 2 * cfg.V + 1) and n == lens
    assert xs[1, 2:].abs().sum() == 0 and tm[3, 1:].abs().sum() == 0
    # longer than max_len keeps the LAST max_len steps (duett/duett.py:165-167)
    long = torch.arange(6 * 2 * cfg.V).float().reshape(6, 2 * cfg.V)
    st, xs, tm, n = O.feats_to_input([long], [b["x_static"][0]], [torch.arange(6).float()], cfg.T)
    assert n == [4] and torch.equal(xs[0, :, :-1], long[-4:]) and torch.equal(tm[0], torch.arange(2, 6).float())


def test_ssl_step_matches_reference_including_host_rng():
    G, cfg = load("g3_ssl"), golden_cfg()
    P = _leaf(G["param"])
    I = G["in"]
    x_static, xs_ts, xs_times, n_t = O.feats_to_input(I["x_ts"], I["x_static"], I["bin_ends"], cfg.T)
    rng = np.random.default_rng(42)
    x, y, mask, y_ev, y_ev_mask = O.pretrain_prep_batch(rng, cfg, xs_ts, n_t, 0.5)
    # integer / index work: bit-exact
    assert torch.equal(x, G["out"]["xs_ts_clipped"])
    assert torch.equal(y, G["out"]["y"]) and torch.equal(mask, G["out"]["mask"])
    assert torch.equal(y_ev, G["out"]["y_events"]) and torch.equal(y_ev_mask, G["out"]["y_events_mask"])
    outs = O.model_forward_pretrain(P, cfg, x_static, x, xs_times)
    for got, key in zip(outs, ("y_hat_value", "y_hat_presence", "y_hat_events", "y_hat_events_presence")):
        assert rel(got, G["out"][key]) < TOL, key
    loss = O.ssl_loss(*outs, y, mask, y_ev, y_ev_mask, 0.2)
    assert rel(loss, G["out"]["loss"]) < TOL
    loss.backward()
    _check_grads(P, G["grad"])


def test_ssl_step_with_two_masked_steps_matches_reference():
    """pretrain_masked_steps = 2 (oracle/make_golden_masked_steps.py): draws with replacement, distinct masked rows zero-padded
    to two rows, loss averaged over the steps (duett/duett.py:199-203, 287-293, 337-349)."""
    G, cfg = load("g7_ssl_masked_steps"), golden_cfg()
    P = _leaf(G["param"])
    I = G["in"]
    x_static, xs_ts, xs_times, n_t = O.feats_to_input(I["x_ts"], I["x_static"], I["bin_ends"], cfg.T)
    rng = np.random.default_rng(42)
    x, y, mask, y_ev, y_ev_mask = O.pretrain_prep_batch(rng, cfg, xs_ts, n_t, 0.5, masked_steps=2)
    assert y.shape == (6, 2, cfg.V) and torch.equal(x, G["out"]["xs_ts_clipped"])
    assert torch.equal(y, G["out"]["y"]) and torch.equal(mask, G["out"]["mask"])
    assert torch.equal(y_ev, G["out"]["y_events"]) and torch.equal(y_ev_mask, G["out"]["y_events_mask"])
    assert torch.equal((x[:, :, -1] > 0).sum(1), G["out"]["n_masked"]) and int(G["out"]["n_masked"].min()) == 1
    outs = O.model_forward_pretrain(P, cfg, x_static, x, xs_times, masked_steps=2)
    for got, key in zip(outs, ("y_hat_value", "y_hat_presence", "y_hat_events", "y_hat_events_presence")):
        assert got.shape == G["out"][key].shape and rel(got, G["out"][key]) < TOL, key
    loss = O.ssl_loss(*outs, y, mask, y_ev, y_ev_mask, 0.2)
    assert rel(loss, G["out"]["loss"]) < TOL
    loss.backward()
    _check_grads(P, G["grad"])
    with pytest.raises(ValueError):
        O.pretrain_prep_batch(np.random.default_rng(0), cfg, xs_ts, [4, 4, 1, 4, 4, 4], 0.5, masked_steps=2)


def test_teacher_patch_dual_step_matches_reference():
    G, cfg = load("g4_teacher"), golden_cfg()
    P = _leaf({k[len("duett."):]: v for k, v in G["param"].items() if k.startswith("duett.")})
    Pt = _leaf({k: v for k, v in G["param"].items() if not k.startswith("duett.")})
    I = G["in"]
    x_static, xs_ts, xs_times, _ = O.feats_to_input(I["x_ts"], I["x_static"], I["bin_ends"], cfg.T)
    out = O.teacher_forward(P, Pt, cfg, x_static, xs_ts, xs_times, I["pixel_values"][:, 1:])
    for k in ("main_logit", "img_logits", "ts_logits", "fusion_logits", "ts_correction", "scaled_correction"):
        assert rel(out[k], G["out"][k]) < TOL, k
    L = O.dual_pathology_loss(out["img_logits"], out["ts_logits"], out["fusion_logits"], I["y_multi"],
                              I["y_multi_mask"], I["label_weights"], I["pos_weight"])
    for k in ("img_per", "ts_per", "fus_per", "img_total", "ts_total", "fus_total"):
        assert rel(L[k], G["out"][k]) < TOL, k
    aux = O.aux_residual_kl(out["img_logits"], out["scaled_correction"], I["y_multi"], I["y_multi_mask"])
    assert rel(aux, G["out"]["aux_residual"]) < TOL
    total = L["total"] + 0.3 * aux
    assert rel(total, G["out"]["loss"]) < TOL
    total.backward()
    _check_grads(P, G["grad"], "duett.")
    _check_grads(Pt, G["grad"], "")


def test_teacher_return_attn_matches_reference():
    """Eval-mode TeacherModel.forward(return_attn=True) of the reference (oracle/make_golden_attn.py): head-averaged
    attention maps of the two cross-attention blocks, latent tokens and logits."""
    G4, G, cfg = load("g4_teacher"), load("g6_teacher_attn"), golden_cfg()
    P = {k[len("duett."):]: v for k, v in G4["param"].items() if k.startswith("duett.")}
    Pt = {k: v for k, v in G4["param"].items() if not k.startswith("duett.")}
    I = G4["in"]
    x_static, xs_ts, xs_times, _ = O.feats_to_input(I["x_ts"], I["x_static"], I["bin_ends"], cfg.T)
    with torch.no_grad():
        out = O.teacher_forward(P, Pt, cfg, x_static, xs_ts, xs_times, I["pixel_values"][:, 1:], training=False)
    for k in ("main_logit", "img_logits", "ts_logits", "fusion_logits", "ts_correction", "scaled_correction", "img_tokens",
              "ts_tokens", "img_attn", "ts_attn"):
        assert out[k].shape == G["out"][k].shape and rel(out[k], G["out"][k]) < TOL, k


def test_ff_inner_expression():
    # x_transformers: inner = int(dim * (d_ff / dim)) — float rounding can give d_ff - 1 (SURVEY §8 note)
    c = O.DuettConfig(3, 34, 24)
    assert c.ff_inner(600) == int(600 * (512 / 600)) and c.ff_inner(840) == int(840 * (512 / 840))
    assert any(int(dim * (512 / dim)) == 511 for dim in range(8, 4000, 8))


def test_binning_oracle_matches_reference_golden():
    """oracle/binning_oracle.py against the output of the reference's own build_stay_tensor (duett/mimic_dataset.py:33-46)
    on 7 synthetic stays: bit-exact, NaN cells included."""
    import os
    import numpy as np
    from oracle import binning_oracle
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g5_binning.npz"))
    x = binning_oracle.bin_events(G["slot_raw"], G["vals"], G["cnts"], G["row_start"], G["means"], G["stds"], int(G["T"]))
    assert x.dtype == np.float32 and x.shape == G["x"].shape
    assert np.array_equal(x, G["x"], equal_nan=True)
