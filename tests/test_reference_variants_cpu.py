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
                                  dict(pretrain_dropout=0.0), dict(pretrain_presence_weight=0.7),
                                  dict(pretrain_masked_steps=3), dict(pretrain_masked_steps=4, pretrain_dropout=0.0),
                                  dict(pretrain_masked_steps=2, pretrain_n_hidden=1, pretrain_d_hidden=16),
                                  dict(pretrain_masked_steps=2, predict_events=False, pretrain_presence=False)])
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


@pytest.fixture(scope="module")
def ref_engine(ref):
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_engine_variants", os.path.join(REF, "training_duett", "engine.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _StubCXR(torch.nn.Module):      # CXR embeddings ride in the pixel_values slot
    d_out = 16

    def forward(self, pv):
        return pv[:, 0], pv[:, 1:]


def _teacher_pair(ref, seed=18, dropout=0.0):
    from multimodal_edema_prediction_b200.models import main_architecture_duett as A
    torch.manual_seed(seed)
    out = []
    for arch in (ref[1], A):
        duett = arch.DuettFeatureExtractor(pretrain=False, **KW)
        perc = arch.PatchDualPathologyPerceiver(7, duett.d_representation, d_latent=32, n_heads=4, dropout=dropout,
                                                head_hidden=16, head_dropout=dropout)
        out.append(arch.TeacherModel(duett, _StubCXR(), perc, patch_dual_pathology_mode=True, d_img=16))
    with torch.no_grad():
        out[0].perceiver.correction_head[-1].weight.normal_(0, 0.05)
        out[0].perceiver.beta.uniform_(0.5, 1.5)
    out[1].load_state_dict(out[0].state_dict(), strict=True)
    return out


def _engine_batch(seed=83):
    b = O.synth_batch(CFG, B=6, seed=seed)
    g = torch.Generator().manual_seed(seed)
    return {"x_ts": tuple(b["x_ts"]), "x_static": tuple(b["x_static"]), "bin_ends": tuple(b["bin_ends"]), "y": b["y"],
            "pixel_values": torch.randn(6, 11, 16, generator=g), "y_multi": (torch.rand(6, 7, generator=g) < 0.3).float(),
            "y_multi_mask": (torch.rand(6, 7, generator=g) < 0.9).float()}


def _same_result(rr, pr, skip=()):
    assert set(rr) == set(pr), set(rr) ^ set(pr)
    for k in rr:
        if k in skip:
            continue
        a, c = rr[k], pr[k]
        if torch.is_tensor(a):
            assert a.shape == c.shape and (rel(c.float(), a.float()) < TOL or float(a.abs().max()) < 1e-9), k
        else:
            assert abs(c - a) <= TOL * max(abs(a), 1e-3), (k, a, c)


def test_engine_lp_stage_and_eval_steps(emu, ref, ref_losses, ref_engine):
    """training_duett/engine.py: the linear-probe stage (everything eval() except the correction head; beta / correction L2
    regularisers, aux residual KL), eval_teacher_batch, train_student_batch (frozen teacher), eval_student_batch — the
    reference's engine on the reference's modules against the product's engine on the product's modules, same weights."""
    from multimodal_edema_prediction_b200.loss import losses_duett as L
    from multimodal_edema_prediction_b200.models import main_architecture_duett as A
    from multimodal_edema_prediction_b200.training_duett import engine
    rt, pt = _teacher_pair(ref)
    lw, pw = torch.tensor([1.0, 0.2, 0.2, 0.2, 0.2, 0.2, 0.2]), torch.tensor([2.0, 1.5, 1.0, 3.0, 1.0, 2.5, 1.2])
    batch, cpu = _engine_batch(), torch.device("cpu")
    kw = dict(beta_l2=0.05, corr_l2=0.1, aux_residual_alpha=0.3)
    rr = ref_engine.train_teacher_dual_pathology_lp_batch(batch, rt, ref_losses.DualPathologyLoss(lw, pw), torch.optim.SGD(rt.parameters(), lr=0.0), cpu, **kw)
    pr = engine.train_teacher_dual_pathology_lp_batch(batch, pt, L.DualPathologyLoss(lw, pw), torch.optim.SGD(pt.parameters(), lr=0.0), cpu, **kw)
    _same_result(rr, pr)
    assert not pt.duett.training and pt.perceiver.correction_head.training and not pt.perceiver.img_cross.training
    _grads_match_named(rt, pt)
    bce = torch.nn.BCEWithLogitsLoss()
    _same_result(ref_engine.eval_teacher_batch(batch, rt, bce, cpu), engine.eval_teacher_batch(batch, pt, bce, cpu))
    # student distilled from the frozen teacher
    torch.manual_seed(19)
    rs = ref[1].StudentModel(ref[1].DuettFeatureExtractor(pretrain=False, **KW), pool="mean", head_hidden=16, head_dropout=0.0)
    ps = A.StudentModel(A.DuettFeatureExtractor(pretrain=False, **KW), pool="mean", head_hidden=16, head_dropout=0.0)
    ps.load_state_dict(rs.state_dict(), strict=True)
    sb = _engine_batch(seed=84)
    rr = ref_engine.train_student_batch(sb, sb, rs, rt, ref_losses.StudentKDLoss(kd_T=3.0, kd_alpha=0.4, pos_weight=2.0),
                                        torch.optim.SGD(rs.parameters(), lr=0.0), cpu)
    pr = engine.train_student_batch(sb, sb, ps, pt, L.StudentKDLoss(kd_T=3.0, kd_alpha=0.4, pos_weight=2.0),
                                    torch.optim.SGD(ps.parameters(), lr=0.0), cpu)
    _same_result(rr, pr)
    _grads_match_named(rs, ps)
    _same_result(ref_engine.eval_student_batch(sb, rs, cpu), engine.eval_student_batch(sb, ps, cpu))


def _grads_match_named(r, p):
    """_grads_match for composite modules (a Model inside a Teacher / Student: keys mapped per sub-module prefix)."""
    from multimodal_edema_prediction_b200 import state_keys
    from multimodal_edema_prediction_b200.duett.duett import Model
    got = {n: (q.grad if q.grad is not None else torch.zeros_like(q)).detach() for n, q in p.named_parameters()}
    for name, m in p.named_modules():
        if isinstance(m, Model):
            state_keys.to_reference(got, name + "." if name else "")
    gs = [float(q.grad.abs().max()) for q in r.parameters() if q.grad is not None]
    gscale = max(gs) if gs else 0.0
    for n, q in r.named_parameters():
        w = q.grad if q.grad is not None else torch.zeros_like(q)
        assert (got[n].double() - w.double()).norm() <= 3e-4 * w.double().norm() + 3e-5 * gscale * max(w.numel(), 64) ** 0.5, n


def test_checkpoint_factories_and_freezing(emu, ref, tmp_path):
    """duett.pretrain_model / fine_tune_model (duett/duett.py:41-46), load_duett_backbone(freeze=True)
    (models/main_architecture_duett.py:98-123), Model.freeze via freeze_encoder (duett/duett.py:485-495): a checkpoint written
    by the REFERENCE's module is read by the product's factories; which parameters stay trainable matches the reference."""
    from multimodal_edema_prediction_b200.duett import duett as D
    from multimodal_edema_prediction_b200.models.main_architecture_duett import load_duett_backbone
    torch.manual_seed(20)
    r = ref[0].pretrain_model(**KW)
    ck = str(tmp_path / "ssl.ckpt")
    torch.save({"state_dict": r.state_dict()}, ck)
    kw = {k: v for k, v in KW.items()}
    ft_r, ft_p = ref[0].fine_tune_model(ck, freeze_encoder=True, **kw), D.fine_tune_model(ck, freeze_encoder=True, **kw)
    assert (ft_p.pretrain, ft_p.aug_mask, ft_p.aug_noise, ft_p.lr, ft_p.weight_decay, ft_p.fusion_method) == \
        (ft_r.pretrain, ft_r.aug_mask, ft_r.aug_noise, ft_r.lr, ft_r.weight_decay, ft_r.fusion_method)
    assert ft_p.transformer_dropout == 0.5 and ft_p.event_transformers[0].dropout == 0.5       # duett/duett.py:44-46
    tr_r = {n for n, q in ft_r.named_parameters() if q.requires_grad}
    tr_p = {n for n, q in ft_p.named_parameters() if q.requires_grad}
    assert tr_r == tr_p and tr_r and all("head" in n for n in tr_r)
    sr, sp = ft_r.state_dict(), ft_p.state_dict()
    assert set(sr) == set(sp) and all(torch.equal(sr[k], sp[k]) for k in sr if not k.startswith("head"))
    # the reference's loader takes no model hyper-parameters (Lightning restores them from the checkpoint; the shim and the
    # product fall back to the constructor defaults d_embedding=24, 2 layers, d_feedforward=512): a default-shaped checkpoint
    torch.manual_seed(21)
    r24 = ref[0].Model(3, 5, 1, pretrain=True, masked_transform_timesteps=4, max_len=4)
    ck = str(tmp_path / "ssl24.ckpt")
    torch.save({"state_dict": r24.state_dict()}, ck)
    bb_r = ref[1].load_duett_backbone(ck, 3, 5, 4, freeze=True)
    bb_p = load_duett_backbone(ck, 3, 5, 4, freeze=True)
    assert not bb_p.training and not any(q.requires_grad for q in bb_p.parameters()) and not any(q.requires_grad for q in bb_r.parameters())
    x, _ = _batch(seed=85)
    with torch.no_grad():
        assert rel(bb_p.encode(bb_p.feats_to_input(x, 6)), bb_r.encode(bb_r.feats_to_input(x, 6))) < TOL


def test_evaluator_console_tables_match_the_reference(ref):
    """format_dual_pathology_gap_table / format_pathology_gap_table (training_duett/evaluator.py:163-178,338-395; imported by
    trainer.py:24-25): the same text as the reference's formatters on a result dict with NaNs; evaluate_pathology (legacy
    pathology_mode teacher) raises instead of scoring something else."""
    import importlib.util
    import random
    spec = importlib.util.spec_from_file_location("reference_evaluator", os.path.join(REF, "training_duett", "evaluator.py"))
    R = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(R)
    from multimodal_edema_prediction_b200.training_duett import evaluator as P
    rnd = random.Random(1)
    keys = ["img_auroc", "ts_auroc", "fus_auroc", "gap_i2f", "gap_t2f", "img_auprc", "ts_auprc", "fus_auprc", "gap_i2f_pr",
            "gap_t2f_pr", "img_bce", "ts_bce", "fus_bce", "delta_bce", "mean_abs_corr", "corr_residual", "beta"]
    per = []
    for i, name in enumerate(["label_edema", "label_cardiomegaly", "x", "label_pleural_effusion_long_name"]):
        r = {"name": name, "n_valid": 10 + i, "pos_frac": 0.1 * i}
        r.update({k: rnd.uniform(-1, 1) for k in keys})
        per.append(r)
    per[2]["img_auroc"] = per[2]["gap_i2f"] = per[1]["beta"] = per[3]["img_auprc"] = float("nan")
    res = {"labels": [r["name"] for r in per], "n": 40, "main_auroc": 0.5, "main_auprc": 0.4, "per_label": per}
    assert P.format_dual_pathology_gap_table(res) == R.format_dual_pathology_gap_table(res)
    for r in per:
        r["img_auprc"] = r["ts_auprc"] = r["fus_auprc"] = float("nan")          # all-NaN macro row
    assert P.format_dual_pathology_gap_table(res) == R.format_dual_pathology_gap_table(res)
    per2 = [{"name": "label_a", "n_valid": 5, "pos_frac": 0.25, "stage2_auroc": 0.7, "stage4_auroc": 0.8, "gap_auroc": 0.1,
             "stage2_auprc": 0.3, "stage4_auprc": 0.25, "gap_auprc": -0.05}]
    assert P.format_pathology_gap_table({"per_label": per2}) == R.format_pathology_gap_table({"per_label": per2})
    with pytest.raises(NotImplementedError):
        P.evaluate_pathology(None, [], torch.device("cpu"), ("a",))
    assert {n for n in dir(R) if not n.startswith("_") and callable(getattr(R, n)) and getattr(R, n).__module__ == R.__name__} \
        <= set(dir(P)), "every public function of the reference's evaluator module exists in the product's"


class _StubBinaryTeacher(torch.nn.Module):
    """Any module with the teacher call signature: returns (main, aux) logits (use_aux_cxr teachers) or the main logit."""

    def __init__(self, tuple_out):
        super().__init__()
        self.tuple_out = tuple_out
        self.a, self.b = torch.nn.Linear(3, 1), torch.nn.Linear(16, 1)

    def forward(self, x_ts, x_static, bin_ends, pixel_values):
        main = self.a(torch.stack(list(x_static))).squeeze(-1) + torch.stack(list(x_ts)).mean((1, 2))
        aux = self.b(pixel_values[:, 0]).squeeze(-1)
        return (main + aux.detach(), aux) if self.tuple_out else main + aux


class _StubPathologyTeacher(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.s2, self.s4 = torch.nn.Linear(16, 7), torch.nn.Linear(16 + 3, 7)

    def forward(self, x_ts, x_static, bin_ends, pixel_values):
        s2 = self.s2(pixel_values[:, 0])
        s4 = self.s4(torch.cat([pixel_values[:, 1:].mean(1), torch.stack(list(x_static))], 1))
        return {"main_logit": s4[:, 0], "stage2_logits": s2, "stage4_logits": s4}


def test_engine_legacy_teacher_steps(emu, ref_losses, ref_engine):
    """train_teacher_batch (tensor or (main, aux) tuple teachers, aux_alpha) and train_teacher_pathology_batch
    (PathologyMultiLabelLoss) — training_duett/engine.py:41-74,93-130: model-agnostic steps, compared with the reference's
    engine on stub teachers (the legacy TeacherModel modes that produce such outputs are not built)."""
    import copy
    from multimodal_edema_prediction_b200.loss import losses_duett as L
    from multimodal_edema_prediction_b200.training_duett import engine
    batch, cpu = _engine_batch(seed=86), torch.device("cpu")
    bce = torch.nn.BCEWithLogitsLoss()
    for tuple_out in (True, False):
        torch.manual_seed(23)
        t_r = _StubBinaryTeacher(tuple_out)
        t_p = copy.deepcopy(t_r)
        rr = ref_engine.train_teacher_batch(batch, t_r, bce, torch.optim.SGD(t_r.parameters(), lr=0.1), cpu, aux_alpha=0.3)
        pr = engine.train_teacher_batch(batch, t_p, bce, torch.optim.SGD(t_p.parameters(), lr=0.1), cpu, aux_alpha=0.3)
        _same_result(rr, pr)
        for a, c in zip(t_r.parameters(), t_p.parameters()):
            assert torch.allclose(a, c, atol=1e-7)                       # the optimizer step happened, identically
    torch.manual_seed(24)
    t_r = _StubPathologyTeacher()
    t_p = copy.deepcopy(t_r)
    lw = torch.rand(7) + 0.1
    rr = ref_engine.train_teacher_pathology_batch(batch, t_r, ref_losses.PathologyMultiLabelLoss(lw, None, 0.5, 1.0),
                                                  torch.optim.SGD(t_r.parameters(), lr=0.1), cpu)
    pr = engine.train_teacher_pathology_batch(batch, t_p, L.PathologyMultiLabelLoss(lw, None, 0.5, 1.0),
                                              torch.optim.SGD(t_p.parameters(), lr=0.1), cpu)
    _same_result(rr, pr)
    for a, c in zip(t_r.parameters(), t_p.parameters()):
        assert torch.allclose(a, c, atol=1e-6)
    with pytest.raises(RuntimeError):
        engine.train_teacher_pathology_batch(batch, _StubBinaryTeacher(False), L.PathologyMultiLabelLoss(lw), None, cpu)
    # every public function of the reference's engine exists with the same parameter names
    import inspect
    for name, fn in vars(ref_engine).items():
        if inspect.isfunction(fn) and not name.startswith("_") and fn.__module__ == ref_engine.__name__:
            assert list(inspect.signature(getattr(engine, name)).parameters) == list(inspect.signature(fn).parameters), name


def test_mimic_dataset_matches_the_reference_dataset(emu, ref):
    """duett/mimic_dataset.py: MIMICDataset (host-only StayRows items binned per batch) + collate_into_seqs + encode_static +
    the d_* / pos_frac accessors against the reference's Dataset on synthetic frames — several rows per slot, slots beyond
    T, zero / NaN counts, NaN values and ages, a variable pair sharing one count column."""
    import pandas as pd
    import duett.mimic_dataset as R
    from multimodal_edema_prediction_b200.duett import mimic_dataset as P
    from multimodal_edema_prediction_b200.duett.duett import Model
    rng = np.random.default_rng(3)
    T, V, B = 6, 4, 5
    all_vars = [f"v{j}" for j in range(V)]
    all_counts = ["c0", "c1", "c1", "c3"]                       # v1 and v2 share a count column
    rows = []
    for b in range(B):
        for _ in range(int(rng.integers(1, 9))):
            r = {"stay_id": 100 + b, "slot_idx": int(rng.integers(0, T + 2))}
            r.update({v: float(rng.normal()) for v in all_vars})
            r.update({c: float(rng.integers(0, 4)) for c in set(all_counts)})
            rows.append(r)
    icu = pd.DataFrame(rows)
    icu.loc[1, "v0"], icu.loc[2, "c1"] = np.nan, np.nan
    static = pd.DataFrame({"stay_id": [100 + b for b in range(B)], "age_at_intime": [30.0, np.nan, 55.0, 71.5, 90.0],
                           "s0": [1, 0, 1, 0, 1], "s1": [0, 1, 0, 1, 0], "label": [0.0, 1.0, 1.0, 0.0, 1.0]})
    meta = {"N_TIMESTEPS": T, "LABEL_COL": "label", "age_mean": 60.0, "age_std": 12.0, "ONEHOT_STATIC": ["s0", "s1"],
            "means": {v: float(rng.normal()) for v in all_vars}, "stds": {v: float(abs(rng.normal()) + 0.2) for v in all_vars},
            "ALL_VARS": all_vars, "ALL_COUNTS": all_counts, "D_STATIC": 3}
    ids = [100 + b for b in range(B)]
    rd, pd_ = R.MIMICDataset(ids, icu, static, meta), P.MIMICDataset(ids, icu, static, meta, device="cpu")   # (default: the CUDA device)
    assert len(rd) == len(pd_) == B and rd.pos_frac() == pd_.pos_frac()
    assert (rd.d_static_num(), rd.d_time_series_num(), rd.d_target()) == (pd_.d_static_num(), pd_.d_time_series_num(), pd_.d_target())
    r_batch = R.collate_into_seqs([rd[i] for i in range(B)])
    p_batch = P.collate_into_seqs([pd_[i] for i in range(B)])
    assert r_batch[1] == p_batch[1]                                                   # labels
    assert all(isinstance(s, P.StayRows) and s.shape == (T, 2 * V) for s in p_batch[0][0])
    for a, c in zip(r_batch[0][1], p_batch[0][1]):
        assert torch.equal(a, c)                                                      # encode_static (NaN age -> 0)
    for a, c in zip(r_batch[0][2], p_batch[0][2]):
        assert torch.equal(a, c)                                                      # bin_ends
    model = Model(3, V, 1, d_embedding=8, masked_transform_timesteps=T, max_len=T, n_duett_layers=1, pretrain=False).eval()
    xs_static, xs_ts, xs_times, n_ts = model.feats_to_input(p_batch[0], B)
    want = torch.stack(list(r_batch[0][0]))
    assert np.array_equal(xs_ts[:, :, :-1].numpy(), want.numpy(), equal_nan=True) and float(xs_ts[:, :, -1].abs().sum()) == 0
    assert n_ts == [T] * B and torch.equal(xs_static, torch.stack(list(r_batch[0][1])))
    (bx_ts, bx_static, bx_times), by = pd_.batch([3, 0, 4])                            # the one-launch batch builder
    assert by == tuple(r_batch[1][i] for i in (3, 0, 4))
    assert np.array_equal(torch.stack(bx_ts).numpy(), want[[3, 0, 4]].numpy(), equal_nan=True)


def test_evaluators_match_the_reference_evaluators(emu, ref):
    """evaluate_dual_pathology / evaluate_binary against the reference's own functions (training_duett/evaluator.py:10-37,
    197-335; sklearn on the host there, the ranking kernel's contract here) on a stub teacher: a label with a single class
    (AUROC undefined), a label with one valid sample (Pearson undefined), tied logits.  (A label that is NEVER valid makes the
    reference's sklearn call raise IndexError past its `except ValueError`; the product reports NaNs for it.)"""
    import training_duett.evaluator as R
    from multimodal_edema_prediction_b200.training_duett import evaluator as P
    K, labels = 5, ("label_a", "label_b", "label_c", "label_d", "label_e")
    g = torch.Generator().manual_seed(9)

    class Perc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.beta = torch.nn.Parameter(torch.tensor([0.5, 1.0, 1.5, 2.0, 0.1]))

    class Teacher(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.perceiver = Perc()

        def forward(self, x_ts, x_static, bin_ends, pv):
            img = (pv[:, :K] * 4).round() / 4                         # quantised: ties between samples
            corr = 0.3 * pv[:, K:2 * K]
            return {"img_logits": img, "ts_logits": pv[:, 2 * K:3 * K], "fusion_logits": img + corr, "scaled_correction": corr,
                    "main_logit": img[:, 0] + corr[:, 0]}

    batches = []
    for i in range(4):
        n = 11
        ym = (torch.rand(n, K, generator=g) < 0.4).float()
        mk = (torch.rand(n, K, generator=g) < 0.8).float()
        ym[:, 2] = 1.0                                                # one class only
        mk[:, 3] = 0.0
        if i == 2:
            mk[4, 3] = 1.0                                            # exactly one valid sample
        batches.append({"x_ts": tuple(torch.randn(2, 3, generator=g) for _ in range(n)),
                        "x_static": tuple(torch.zeros(1) for _ in range(n)), "bin_ends": tuple(torch.zeros(1) for _ in range(n)),
                        "y": (torch.rand(n, generator=g) < 0.5).float(), "pixel_values": torch.randn(n, 3 * K, generator=g),
                        "y_multi": ym, "y_multi_mask": mk})
    model, cpu = Teacher(), torch.device("cpu")
    rr, pr = R.evaluate_dual_pathology(model, batches, cpu, labels), P.evaluate_dual_pathology(model, batches, cpu, labels)

    def close(a, c, what):
        if isinstance(a, float) and a != a:
            assert c != c, (what, a, c)
        elif isinstance(a, (int, float)):
            assert abs(a - c) <= 2e-6 * max(1.0, abs(a)), (what, a, c)
        else:
            assert a == c, (what, a, c)

    assert set(rr) == set(pr)
    for k in ("labels", "n", "main_auroc", "main_auprc"):
        close(rr[k], pr[k], k)
    for a, c in zip(rr["per_label"], pr["per_label"]):
        assert set(a) == set(c)
        for k in a:
            close(a[k], c[k], (a["name"], k))
    assert P.format_dual_pathology_gap_table(pr).count("--") == R.format_dual_pathology_gap_table(rr).count("--") > 0
    for b in batches:
        b["y_multi_mask"][:, 1] = 0.0                                 # never valid: NaNs, n_valid 0 (the reference raises here)
    never = P.evaluate_dual_pathology(model, batches, cpu, labels)["per_label"][1]
    assert never["n_valid"] == 0 and all(never[k] != never[k] for k in ("img_auroc", "fus_auprc", "img_bce", "pos_frac", "corr_residual"))
    fwd_r, fwd_p = R.make_teacher_forward(), P.make_teacher_forward()
    rb, pb = R.evaluate_binary(model, batches, cpu, fwd_r), P.evaluate_binary(model, batches, cpu, fwd_p)
    assert set(rb) == set(pb)
    for k in rb:
        close(rb[k], pb[k], k)


def test_engine_steps_with_frozen_parts(emu, ref, ref_losses, ref_engine):
    """The trainer's freezing patterns (training_duett/trainer.py:170-207, engine.py:7-20): (1) a fully frozen DuETT backbone
    goes to eval() inside the teacher step (running BatchNorm statistics) while the fusion head trains; (2) the LP stage
    with everything frozen except perceiver.correction_head and beta, and the correction head's Dropout probability
    overridden through its nn.Dropout modules (p = 0 here so both sides are deterministic)."""
    from multimodal_edema_prediction_b200.loss import losses_duett as L
    from multimodal_edema_prediction_b200.training_duett import engine
    lw, pw = torch.tensor([1.0, 0.2, 0.2, 0.2, 0.2, 0.2, 0.2]), torch.tensor([2.0, 1.5, 1.0, 3.0, 1.0, 2.5, 1.2])
    batch, cpu = _engine_batch(seed=87), torch.device("cpu")
    # (1) frozen backbone
    rt, pt = _teacher_pair(ref, seed=25)
    for t in (rt, pt):
        for q in t.duett.parameters():
            q.requires_grad = False
    rr = ref_engine.train_teacher_dual_pathology_batch(batch, rt, ref_losses.DualPathologyLoss(lw, pw),
                                                       torch.optim.SGD([q for q in rt.parameters() if q.requires_grad], lr=0.0),
                                                       cpu, aux_residual_alpha=0.2)
    pr = engine.train_teacher_dual_pathology_batch(batch, pt, L.DualPathologyLoss(lw, pw),
                                                   torch.optim.SGD([q for q in pt.parameters() if q.requires_grad], lr=0.0),
                                                   cpu, aux_residual_alpha=0.2)
    _same_result(rr, pr)
    assert not pt.duett.training and pt.perceiver.training and not rt.duett.training
    assert all(q.grad is None for q in pt.duett.parameters())
    for (n, a), (_, c) in zip(rt.perceiver.named_parameters(), pt.perceiver.named_parameters()):
        w = a.grad if a.grad is not None else torch.zeros_like(a)
        g = c.grad if c.grad is not None else torch.zeros_like(c)
        assert (g - w).norm() <= 3e-4 * w.norm() + 1e-6, n
    sr, sp = rt.state_dict(), pt.state_dict()
    assert all(torch.equal(sr[k], sp[k]) for k in sr if "running" in k or "num_batches" in k)      # frozen BN did not move
    # (2) LP stage: only the correction head + beta train
    rt, pt = _teacher_pair(ref, seed=26, dropout=0.1)
    for t in (rt, pt):
        for q in t.parameters():
            q.requires_grad = False
        for q in t.perceiver.correction_head.parameters():
            q.requires_grad = True
        t.perceiver.beta.requires_grad = True
        n_drop = 0
        for m in t.perceiver.correction_head.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
                n_drop += 1
        assert n_drop == 1
    kw = dict(beta_l2=0.02, corr_l2=0.3, aux_residual_alpha=0.1)
    rr = ref_engine.train_teacher_dual_pathology_lp_batch(batch, rt, ref_losses.DualPathologyLoss(lw, pw),
                                                          torch.optim.SGD([q for q in rt.parameters() if q.requires_grad], lr=0.0), cpu, **kw)
    pr = engine.train_teacher_dual_pathology_lp_batch(batch, pt, L.DualPathologyLoss(lw, pw),
                                                      torch.optim.SGD([q for q in pt.parameters() if q.requires_grad], lr=0.0), cpu, **kw)
    _same_result(rr, pr)
    assert all(q.grad is None for q in pt.duett.parameters()) and all(q.grad is None for q in rt.duett.parameters())
    pn = dict(pt.named_parameters())
    for n, a in rt.named_parameters():
        if n.startswith("duett."):
            continue
        c = pn[n]                                                    # perceiver / img_proj names are the reference's
        assert (a.grad is None) == (c.grad is None), n
        if a.grad is not None:
            assert n.startswith("perceiver.correction_head") or n == "perceiver.beta", n
            assert (c.grad - a.grad).norm() <= 3e-4 * a.grad.norm() + 1e-7, n


def test_multi_step_training_and_gradient_accumulation(emu, ref, ref_losses):
    """Three AdamW steps of the KD student (BatchNorm running statistics evolve, gradients are re-created every step) and a
    two-micro-batch gradient accumulation (no zero_grad in between): parameters and buffers stay in step with the reference's
    modules under the same torch optimizer — the product's kernels accumulate straight into .grad, so stale or doubly counted
    gradients would show here."""
    from multimodal_edema_prediction_b200 import state_keys
    from multimodal_edema_prediction_b200.loss import losses_duett as L
    from multimodal_edema_prediction_b200.models import main_architecture_duett as A
    torch.manual_seed(27)
    rs = ref[1].StudentModel(ref[1].DuettFeatureExtractor(pretrain=False, **KW), pool="mean", head_hidden=16, head_dropout=0.0)
    ps = A.StudentModel(A.DuettFeatureExtractor(pretrain=False, **KW), pool="mean", head_hidden=16, head_dropout=0.0)
    ps.load_state_dict(rs.state_dict(), strict=True)
    rs.train(); ps.train()
    ro, po = torch.optim.AdamW(rs.parameters(), lr=3e-3, weight_decay=0.05), torch.optim.AdamW(ps.parameters(), lr=3e-3, weight_decay=0.05)
    rl, pl = ref_losses.StudentKDLoss(kd_T=2.0, kd_alpha=0.5), L.StudentKDLoss(kd_T=2.0, kd_alpha=0.5)
    g = torch.Generator().manual_seed(28)
    for step in range(3):
        ro.zero_grad(); po.zero_grad()
        for micro in range(2 if step == 1 else 1):                  # step 1 accumulates two micro-batches
            x, b = _batch(seed=90 + 10 * step + micro)
            z_t = torch.randn(6, generator=g)
            lr_ = rl(rs(x[0], x[1], list(x[2])), z_t, b["y"])["total"]
            lp = pl(ps(x[0], x[1], list(x[2])), z_t, b["y"])["total"]
            assert rel(lp, lr_) < 2e-4, (step, micro)
            lr_.backward(); lp.backward()
        ro.step(); po.step()
    sr, sp = rs.state_dict(), ps.state_dict()
    assert set(sr) == set(sp)
    for k in sr:
        if sr[k].dtype.is_floating_point:
            assert (sp[k] - sr[k]).norm() <= 2e-3 * sr[k].norm() + 1e-5, k      # AdamW normalises: early steps amplify fp32 noise
        else:
            assert torch.equal(sp[k], sr[k]), k


def test_real_mimic_working_point(emu, ref, ref_losses, ref_engine):
    """The reference's real configuration (d_embedding 24, 2+2 layers, T = 24 hourly bins, V = 34 variables, 24
    static features, heads 2, d_ff 512 -> E = 600, E' = 840 whose x_transformers FFN width is int(dim * (512 / dim))): one
    teacher step through the engines and one SSL step, product next to the reference."""
    from multimodal_edema_prediction_b200.duett.duett import Model
    from multimodal_edema_prediction_b200.loss import losses_duett as L
    from multimodal_edema_prediction_b200.models import main_architecture_duett as A
    from multimodal_edema_prediction_b200.training_duett import engine
    cfg = O.DuettConfig(d_static_num=24, d_time_series_num=34, n_timesteps=24, d_embedding=24, n_layers=2)
    kw = dict(d_static_num=24, d_time_series_num=34, d_target=1, masked_transform_timesteps=24, max_len=24)
    B = 4
    b = O.synth_batch(cfg, B=B, seed=91)
    g = torch.Generator().manual_seed(92)
    batch = {"x_ts": tuple(b["x_ts"]), "x_static": tuple(b["x_static"]), "bin_ends": tuple(b["bin_ends"]), "y": b["y"],
             "pixel_values": torch.randn(B, 1 + 49, 768, generator=g), "y_multi": (torch.rand(B, 7, generator=g) < 0.3).float(),
             "y_multi_mask": (torch.rand(B, 7, generator=g) < 0.9).float()}

    class CXR(torch.nn.Module):
        d_out = 768

        def forward(self, pv):
            return pv[:, 0], pv[:, 1:]

    torch.manual_seed(29)
    teachers = []
    for arch in (ref[1], A):
        duett = arch.DuettFeatureExtractor(pretrain=False, **kw)
        perc = arch.PatchDualPathologyPerceiver(7, duett.d_representation, d_latent=256, n_heads=4, dropout=0.0, head_hidden=64,
                                                head_dropout=0.0)
        teachers.append(arch.TeacherModel(duett, CXR(), perc, patch_dual_pathology_mode=True, d_img=768))
    rt, pt = teachers
    assert pt.duett.event_transformers[0].ff_inner == int(600 * (512 / 600)) and pt.duett.time_transformers[0].ff_inner == int(840 * (512 / 840))
    with torch.no_grad():
        rt.perceiver.correction_head[-1].weight.normal_(0, 0.05)
    pt.load_state_dict(rt.state_dict(), strict=True)
    lw, cpu = torch.tensor([1.0, 0.2, 0.2, 0.2, 0.2, 0.2, 0.2]), torch.device("cpu")
    rr = ref_engine.train_teacher_dual_pathology_batch(batch, rt, ref_losses.DualPathologyLoss(lw), torch.optim.SGD(rt.parameters(), lr=0.0), cpu,
                                                       aux_residual_alpha=0.3)
    pr = engine.train_teacher_dual_pathology_batch(batch, pt, L.DualPathologyLoss(lw), torch.optim.SGD(pt.parameters(), lr=0.0), cpu,
                                                   aux_residual_alpha=0.3)
    _same_result(rr, pr)
    _grads_match_named(rt, pt)
    r, p = _pair(ref[0].Model, Model, seed=30, pretrain=True, **kw)
    r.train(); p.train()
    x = (tuple(b["x_ts"]), tuple(b["x_static"]), list(b["bin_ends"]))
    r.rng, p.rng = np.random.default_rng(8), np.random.default_rng(8)
    lr_, lp = r.training_step((x, tuple([0.0] * B)), 0), p.training_step((x, tuple([0.0] * B)), 0)
    assert rel(lp, lr_) < TOL
    lr_.backward(); lp.backward()
    _grads_match(r, p)
