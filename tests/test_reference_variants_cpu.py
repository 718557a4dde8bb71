"""Option variants of the hot-path modules against the REFERENCE'S OWN modules, run side by side in this process: the
reference's duett/duett.py and models/main_architecture_duett.py are imported unmodified from /root/reference through
oracle/shims (as oracle/make_golden.py does), the product modules run on the kernel-contract emulator, both load the same
state dict and see the same batch.  Covers the constructor / call options the golden fixtures do not exercise: the three
fusion methods, representation / save_representation returns, eval mode (running statistics), SSL heads with a hidden
layer, SSL without event prediction or without the presence heads, the student's rep_token pooling, the perceiver's
ts_ablation modes.  Skipped where the reference tree does not exist (the GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch

import ops_emulator
from golden_util import rel
from oracle import duett_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DUETT_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "duett", "duett.py")), reason="reference tree not present")

CFG = O.DuettConfig(d_static_num=3, d_time_series_num=5, n_timesteps=4, d_embedding=8, n_layers=2, d_feedforward=96)
KW = dict(d_static_num=3, d_time_series_num=5, d_target=1, d_embedding=8, masked_transform_timesteps=4, max_len=4,
          n_duett_layers=2, d_feedforward=96)
TOL = 5e-5


@pytest.fixture
def emu(monkeypatch):
    ops_emulator.install(monkeypatch)


@pytest.fixture(scope="module")
def ref():
    """The reference's modules (duett.duett, models.main_architecture_duett), imported in place."""
    added = [os.path.join(ROOT, "oracle", "shims"), REF]
    sys.path[:0] = added
    try:
        import duett.duett as ref_duett
        import models.main_architecture_duett as ref_arch
        torch.set_float32_matmul_precision("highest")      # the reference sets 'high' at import (TF32; irrelevant on CPU)
        yield ref_duett, ref_arch
    finally:
        for p in added:
            sys.path.remove(p)


def _batch(B=6, seed=77, density=0.5):
    b = O.synth_batch(CFG, B=B, seed=seed, density=density)
    return (tuple(b["x_ts"]), tuple(b["x_static"]), list(b["bin_ends"])), b


def _pair(ref_cls, prod_cls, seed=0, **kw):
    torch.manual_seed(seed)
    r = ref_cls(**kw)
    p = prod_cls(**kw)
    p.load_state_dict(r.state_dict(), strict=True)
    return r, p


def _grads_match(r, p):
    """Every parameter gradient of the product module p (keyed like the reference through state_keys) against module r's."""
    from multimodal_edema_prediction_b200 import state_keys
    got = {n: (q.grad if q.grad is not None else torch.zeros_like(q)).detach() for n, q in p.named_parameters()}
    state_keys.to_reference(got, "")
    gscale = max(float(q.grad.abs().max()) for q in r.parameters() if q.grad is not None)
    for n, q in r.named_parameters():
        w = q.grad if q.grad is not None else torch.zeros_like(q)
        # absolute floor for cancellation-dominated gradients (a bias feeding ReLU -> BatchNorm, a ScaleNorm gain: fp32 noise of
        # ~1e-4 x the largest gradient entry in BOTH implementations); single-element gains are sums over every token, so their
        # floor is that of a 64-element tensor
        assert (got[n].double() - w.double()).norm() <= 3e-4 * w.double().norm() + 3e-5 * gscale * max(w.numel(), 64) ** 0.5, n


def _clone_x(x):
    return tuple(t.clone() if torch.is_tensor(t) else t for t in x)


@pytest.mark.parametrize("fusion", ["rep_token", "averaging", "masked_embed"])
def test_supervised_fusion_methods_train_and_eval(emu, ref, fusion):
    """Model.forward(pretrain=False) for every fusion_method (duett/duett.py:282-299), training mode (batch statistics) and
    eval mode (running statistics), plus representation=True and save_representation."""
    from multimodal_edema_prediction_b200.duett.duett import Model
    r, p = _pair(ref[0].Model, Model, seed=10, pretrain=False, fusion_method=fusion, **KW)
    x, b = _batch()
    xin_r, xin_p = r.feats_to_input(x, 6), p.feats_to_input(x, 6)
    for a, c in zip(xin_r[:3], xin_p[:3]):
        assert torch.equal(a, c)
    assert list(xin_r[3]) == list(xin_p[3])
    if fusion == "masked_embed":                       # exactly one flagged timestep per sample (what aug / SSL masking produce)
        steps = torch.tensor([0, 3, 1, 2, 2, 0])
        for xin in (xin_r, xin_p):
            xin[1][torch.arange(6), steps, :] = 0.0
            xin[1][torch.arange(6), steps, -1] = 1.0
    for mode in ("train", "eval"):
        getattr(r, mode)(); getattr(p, mode)()
        with torch.no_grad():
            zr, zp = r.forward(_clone_x(xin_r)), p.forward(_clone_x(xin_p))
            assert zr.shape == zp.shape and rel(zp, zr) < TOL, (fusion, mode)
            rr, rp = r.forward(_clone_x(xin_r), representation=True), p.forward(_clone_x(xin_p), representation=True)
            assert rr.shape == rp.shape and rel(rp, rr) < TOL, (fusion, mode, "representation")
    # after the same number of training-mode forwards the running statistics agree
    sr, sp = r.state_dict(), p.state_dict()
    for k in sr:
        if "running" in k:
            assert rel(sp[k], sr[k]) < 1e-5 or float(sr[k].abs().max()) < 1e-6, k
        if "num_batches_tracked" in k:
            assert int(sp[k].reshape(-1)[0]) == int(sr[k].reshape(-1)[0]), k
    r.save_representation = p.save_representation = True
    with torch.no_grad():
        (zr, hr), (zp, hp) = r.forward(_clone_x(xin_r)), p.forward(_clone_x(xin_p))
    assert rel(zp, zr) < TOL and rel(hp, hr) < TOL


@pytest.mark.parametrize("opts", [dict(pretrain_n_hidden=1, pretrain_d_hidden=16), dict(predict_events=False),
                                  dict(pretrain_dropout=0.0), dict(pretrain_presence_weight=0.7)])
def test_ssl_step_option_variants(emu, ref, opts):
    """Model.training_step(pretrain=True) with SSL heads that have a hidden layer + BatchNorm (pretrain_n_hidden=1), without
    event prediction, without variable dropout, with another presence weight: masking bit-exact, loss and gradients."""
    from multimodal_edema_prediction_b200.duett.duett import Model
    r, p = _pair(ref[0].Model, Model, seed=11, pretrain=True, **KW, **opts)
    r.train(); p.train()
    x, b = _batch(seed=78)
    r.rng, p.rng = np.random.default_rng(5), np.random.default_rng(5)
    pre_r, pre_p = r.pretrain_prep_batch(x, 6), p.pretrain_prep_batch(x, 6)
    assert torch.equal(pre_r[0][1], pre_p[0][1])
    for a, c in zip(pre_r[1:], pre_p[1:]):
        assert (torch.is_tensor(a) and torch.equal(a, c)) or (not torch.is_tensor(a) and len(a) == len(c) == 0)
    r.rng, p.rng = np.random.default_rng(5), np.random.default_rng(5)
    y = tuple([0.0] * 6)
    lr_, lp = r.training_step((x, y), 0), p.training_step((x, y), 0)
    assert rel(lp, lr_) < TOL
    lr_.backward(); lp.backward()
    _grads_match(r, p)


@pytest.mark.parametrize("pool", ["mean", "rep_token"])
def test_student_pooling_modes(emu, ref, pool):
    """StudentModel(pool=...) (models/main_architecture_duett.py:1216-1234), train and eval."""
    from multimodal_edema_prediction_b200.models.main_architecture_duett import DuettFeatureExtractor, StudentModel
    torch.manual_seed(12)
    rs = ref[1].StudentModel(ref[1].DuettFeatureExtractor(pretrain=False, **KW), pool=pool, head_hidden=16, head_dropout=0.0)
    ps = StudentModel(DuettFeatureExtractor(pretrain=False, **KW), pool=pool, head_hidden=16, head_dropout=0.0)
    ps.load_state_dict(rs.state_dict(), strict=True)
    x, b = _batch(seed=79)
    for mode in ("train", "eval"):
        getattr(rs, mode)(); getattr(ps, mode)()
        with torch.no_grad():
            zr, zp = rs(x[0], x[1], list(x[2])), ps(x[0], x[1], list(x[2]))
        assert zr.shape == zp.shape and rel(zp, zr) < TOL, (pool, mode)
    with pytest.raises(ValueError):
        StudentModel(DuettFeatureExtractor(pretrain=False, **KW), pool="max")(x[0], x[1], list(x[2]))


@pytest.mark.parametrize("ablation", ["hourly_only", "full", "rep_only"])
def test_perceiver_ts_ablation_modes(emu, ref, ablation):
    """PatchDualPathologyPerceiver.forward(ts_ablation=...) (models/main_architecture_duett.py:595-654) incl. the
    attention maps of return_attn=True; unknown modes and a 2-D ts_tokens raise ValueError like the reference."""
    from multimodal_edema_prediction_b200.models.main_architecture_duett import PatchDualPathologyPerceiver
    torch.manual_seed(13)
    kw = dict(d_latent=32, n_heads=4, dropout=0.0, head_hidden=16, head_dropout=0.0)
    rp = ref[1].PatchDualPathologyPerceiver(7, 48, **kw)
    pp = PatchDualPathologyPerceiver(7, 48, **kw)
    with torch.no_grad():
        rp.correction_head[-1].weight.normal_(0, 0.05)
        rp.beta.uniform_(0.5, 1.5)
    pp.load_state_dict(rp.state_dict(), strict=True)
    rp.eval(); pp.eval()
    g = torch.Generator().manual_seed(3)
    tok, patches = torch.randn(5, 5, 48, generator=g), torch.randn(5, 11, 32, generator=g)
    with torch.no_grad():
        ro, po = rp(tok, patches, return_attn=True, ts_ablation=ablation), pp(tok, patches, return_attn=True, ts_ablation=ablation)
    assert set(ro) == set(po)
    for k in ro:
        assert ro[k].shape == po[k].shape and rel(po[k], ro[k]) < TOL, (ablation, k)
    for bad in (dict(ts_ablation="daily"),):
        with pytest.raises(ValueError):
            rp(tok, patches, **bad)
        with pytest.raises(ValueError):
            pp(tok, patches, **bad)
    with pytest.raises(ValueError):
        pp(tok[0], patches)


@pytest.mark.parametrize("opts", [dict(pretrain_presence=False), dict(pretrain_value=False)])
def test_ssl_without_value_or_presence_heads(emu, ref, opts):
    """SSL with only the value heads or only the presence heads (duett/duett.py:112-122,336-357)."""
    from multimodal_edema_prediction_b200.duett.duett import Model
    r, p = _pair(ref[0].Model, Model, seed=14, pretrain=True, **KW, **opts)
    r.train(); p.train()
    x, b = _batch(seed=80)
    r.rng, p.rng = np.random.default_rng(6), np.random.default_rng(6)
    y = tuple([0.0] * 6)
    lr_, lp = r.training_step((x, y), 0), p.training_step((x, y), 0)
    assert rel(lp, lr_) < TOL
    lr_.backward(); lp.backward()
    _grads_match(r, p)


def test_feats_to_input_augmentation_ragged_and_truncation(emu, ref):
    """Model.feats_to_input (duett/duett.py:159-187) in training mode with aug_noise / aug_mask (same torch RNG calls in
    the same order => identical draws), ragged lengths and samples longer than max_len (the LAST max_len steps are kept)."""
    from multimodal_edema_prediction_b200.duett.duett import Model
    r, p = _pair(ref[0].Model, Model, seed=15, pretrain=False, aug_noise=0.1, aug_mask=0.3, **KW)
    r.train(); p.train()
    g = torch.Generator().manual_seed(4)
    lens = [4, 2, 6, 3, 5, 1]
    x_ts = tuple(torch.randn(n, 10, generator=g) for n in lens)
    x_static = tuple(torch.randn(3, generator=g) for _ in lens)
    times = [torch.arange(1, n + 1).float() / 24 for n in lens]
    outs = []
    for m in (r, p):
        torch.manual_seed(123)
        outs.append(m.feats_to_input((tuple(t.clone() for t in x_ts), x_static, [t.clone() for t in times]), 6))
    for a, c in zip(outs[0][:3], outs[1][:3]):
        assert a.shape == c.shape and torch.equal(a, c)
    assert list(outs[0][3]) == list(outs[1][3]) == [4, 2, 4, 3, 4, 1]
    assert float(outs[1][1][:, :, -1].sum()) > 0            # some steps were masked by aug_mask
    assert torch.equal(x_ts[2], x_ts[2]) and x_ts[0].shape == (4, 10)      # the caller's tensors keep their shape
    r.eval(); p.eval()                                       # no augmentation outside training
    a, c = r.feats_to_input((x_ts, x_static, list(times)), 6), p.feats_to_input((x_ts, x_static, list(times)), 6)
    for u, v in zip(a[:3], c[:3]):
        assert torch.equal(u, v)


def test_validation_and_test_steps(emu, ref):
    """validation_step / test_step losses (duett/duett.py:373-441) for the supervised model with class weights."""
    from multimodal_edema_prediction_b200.duett.duett import Model
    r, p = _pair(ref[0].Model, Model, seed=16, pretrain=False, fusion_method="rep_token", pos_frac=0.25, **KW)
    r.eval(); p.eval()
    x, b = _batch(seed=81)
    y = tuple(b["y"].tolist())
    with torch.no_grad():
        r.validation_step((x, y), 0)
        lp = p.validation_step((x, y), 0)
        tr, tp = r.test_step((x, y), 0), p.test_step((x, y), 0)
    assert rel(tp[0], tr[0]) < TOL and rel(lp, tr[0]) < TOL and tp[0].dtype == tr[0].dtype == torch.float64


@pytest.fixture(scope="module")
def ref_losses(ref):
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_losses_duett", os.path.join(REF, "loss", "losses_duett.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _loss_grads(fn, *logits):
    zs = [z.clone().requires_grad_(True) for z in logits]
    out = fn(*zs)
    out["total"].backward()
    return out, [z.grad for z in zs]


@pytest.mark.parametrize("kw", [dict(), dict(kd_T=2.0, kd_alpha=0.0), dict(kd_alpha=1.0, pos_weight=3.0),
                                dict(kd_T=8.0, kd_alpha=0.3, pos_weight=0.5)])
def test_student_kd_loss_options(emu, ref_losses, kw):
    """StudentKDLoss / VanillaKLKD (loss/losses_duett.py:8-57): temperatures, the pure-BCE and pure-KD ends of alpha, class
    weight on / off; logits large enough to reach the eps clamp."""
    from multimodal_edema_prediction_b200.loss import losses_duett as L
    g = torch.Generator().manual_seed(21)
    z_s, z_t = torch.randn(33, generator=g) * 6, torch.randn(33, generator=g) * 6
    z_s[0], z_t[1] = 40.0, -40.0                                # saturated probabilities: the clamp at eps / 1 - eps
    y = (torch.rand(33, generator=g) < 0.3).float()
    ro, rg = _loss_grads(lambda z: ref_losses.StudentKDLoss(**kw)(z, z_t, y), z_s)
    po, pg = _loss_grads(lambda z: L.StudentKDLoss(**kw)(z, z_t, y), z_s)
    assert set(ro) == set(po)
    for k in ro:
        assert rel(po[k], ro[k]) < TOL or abs(float(ro[k])) < 1e-12 and abs(float(po[k])) < 1e-9, (k, float(po[k]), float(ro[k]))
    assert rel(pg[0], rg[0]) < 1e-4
    assert set(L.KD_LOSSES) == set(ref_losses.KD_LOSSES)
    with pytest.raises(Exception):
        L.build_kd_loss("no_such_kd")


@pytest.mark.parametrize("with_pw", [True, False])
def test_pathology_losses_options(emu, ref_losses, with_pw):
    """DualPathologyLoss and PathologyMultiLabelLoss (loss/losses_duett.py:63-194) with and without pos_weight, a label that is
    masked out for every sample (denominator eps) and non-default branch weights."""
    from multimodal_edema_prediction_b200.loss import losses_duett as L
    g = torch.Generator().manual_seed(22)
    K, B = 7, 19
    img, ts, fus = (torch.randn(B, K, generator=g) * 2 for _ in range(3))
    y = (torch.rand(B, K, generator=g) < 0.3).float()
    m = (torch.rand(B, K, generator=g) < 0.8).float()
    m[:, 4] = 0.0
    lw = torch.rand(K, generator=g) + 0.1
    pw = (torch.rand(K, generator=g) * 3 + 0.2) if with_pw else None
    ro, rg = _loss_grads(lambda a, b, c: ref_losses.DualPathologyLoss(lw, pw, 0.3, 0.7, 1.1)(a, b, c, y, m), img, ts, fus)
    po, pg = _loss_grads(lambda a, b, c: L.DualPathologyLoss(lw, pw, 0.3, 0.7, 1.1)(a, b, c, y, m), img, ts, fus)
    assert set(ro) == set(po)
    for k in ro:
        assert ro[k].shape == po[k].shape and (rel(po[k], ro[k]) < TOL or float(ro[k].abs().max()) == 0.0), k
    for a, b in zip(pg, rg):
        assert rel(a, b) < 1e-4
    ro, rg = _loss_grads(lambda a, b: ref_losses.PathologyMultiLabelLoss(lw, pw, 0.4, 0.9)(a, b, y, m), img, fus)
    po, pg = _loss_grads(lambda a, b: L.PathologyMultiLabelLoss(lw, pw, 0.4, 0.9)(a, b, y, m), img, fus)
    assert set(ro) == set(po)
    for k in ro:
        assert ro[k].shape == po[k].shape and (rel(po[k], ro[k]) < TOL or float(ro[k].abs().max()) == 0.0), k
    for a, b in zip(pg, rg):
        assert rel(a, b) < 1e-4


@pytest.mark.parametrize("dims", [dict(d_embedding=16, n_transformer_head=4, d_feedforward=None, n_duett_layers=1),
                                  dict(d_embedding=8, n_hidden_head=2, d_hidden_head=12, n_hidden_tab_encoder=2,
                                       d_hidden_tab_encoder=10, n_duett_layers=3, d_feedforward=40)])
def test_constructor_dimension_options(emu, ref, dims):
    """Head count / head width, d_feedforward=None (= 4 d), deeper head and static-encoder MLPs, 1 and 3 DuETT layers: same
    state-dict keys and shapes as the reference and the same supervised training step."""
    from multimodal_edema_prediction_b200.duett.duett import Model
    kw = dict(KW)
    kw.update(dims)
    r, p = _pair(ref[0].Model, Model, seed=17, pretrain=False, fusion_method="rep_token", pos_frac=0.4, **kw)
    sr, sp = r.state_dict(), p.state_dict()
    assert set(sr) == set(sp) and all(sr[k].shape == sp[k].shape for k in sr)
    r.train(); p.train()
    x, b = _batch(seed=82)
    y = tuple(b["y"].tolist())
    lr_, lp = r.training_step((x, y), 0), p.training_step((x, y), 0)
    assert rel(lp, lr_) < TOL
    lr_.backward(); lp.backward()
    _grads_match(r, p)
