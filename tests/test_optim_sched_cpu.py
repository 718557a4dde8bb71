"""Host logic of the optimiser recipes: FusedAdamW.from_trainer_args reproduces the reference's LR groups
(training_duett/trainer.py:77-116) and its LinearLR -> CosineAnnealingLR schedule (trainer.py:119-125, checked against
torch's own SequentialLR); FusedAdamW.for_ssl reproduces WarmUpCallback (duett/train_duett_ssl.py:27-50); checkpoint
loading refuses a foreign x_transformers key layout instead of silently re-initialising the encoders."""
import math
import os
import sys
import types

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
KW = dict(d_static_num=3, d_time_series_num=5, d_target=1, d_embedding=8, masked_transform_timesteps=4, max_len=4,
          n_duett_layers=2, d_feedforward=96)


@pytest.fixture()
def emu(monkeypatch):
    import ops_emulator
    ops_emulator.install(monkeypatch)


def _teacher():
    from multimodal_edema_prediction_b200.models.main_architecture_duett import (DuettFeatureExtractor,
                                                                                 PatchDualPathologyPerceiver, TeacherModel)

    class StubCXR(torch.nn.Module):
        d_out = 16

        def forward(self, pv):
            return pv[:, 0], pv[:, 1:]

    duett = DuettFeatureExtractor(pretrain=False, **KW)
    perceiver = PatchDualPathologyPerceiver(7, duett.d_representation, d_latent=32, n_heads=4, dropout=0.0, head_hidden=16,
                                            head_dropout=0.0)
    return TeacherModel(duett, StubCXR(), perceiver, patch_dual_pathology_mode=True, d_img=16)


def test_trainer_param_groups_and_warmup_cosine_match_torch(emu):
    from multimodal_edema_prediction_b200.ddp import FlatParams, FusedAdamW
    teacher = _teacher()
    args = types.SimpleNamespace(lr=2e-4, backbone_lr_mult=0.2, query_lr_mult=0.2, correction_lr_mult=1.0, weight_decay=0.05,
                                 warmup_steps=7, min_lr_ratio=0.01)
    total = 40
    # the reference's grouping rule (training_duett/trainer.py:88-103), restated on names
    want = {}
    for n, p in teacher.named_parameters():
        if n.startswith(("duett.", "cxr.")):
            want[n] = "backbone"
        elif "correction_head" in n or n.endswith(".beta") or n == "beta":
            want[n] = "correction_head"
        elif n.endswith("_queries"):
            want[n] = "pathology_queries"
        else:
            want[n] = "rest"
    flat = FlatParams(teacher)
    opt = FusedAdamW.from_trainer_args(flat, args, total_steps=total)
    got = {}
    for lo, hi, s, wd, gi in opt.runs:
        for n, off in zip(flat.names, flat.offsets):
            if lo <= off < hi:
                got[n] = opt.group_names[gi]
                assert wd == 0.05
    used = {n for n in want if n not in flat.unused}
    assert {n: want[n] for n in used} == got and {"backbone", "correction_head", "pathology_queries", "rest"} == set(got.values())
    assert all(n.startswith("duett.") for n in flat.unused) and flat.unused          # backbone's own heads: skipped like grad=None
    # torch's schedule on a dummy optimiser with the same four base LRs
    mult = {"backbone": 0.2, "correction_head": 1.0, "pathology_queries": 0.2, "rest": 1.0}
    dummy = torch.optim.AdamW([{"params": [torch.nn.Parameter(torch.zeros(1))], "lr": args.lr * mult[g]} for g in opt.group_names])
    from torch.optim.lr_scheduler import CosineAnnealingLR, LinearLR, SequentialLR
    warm = LinearLR(dummy, start_factor=1e-4, end_factor=1.0, total_iters=args.warmup_steps)
    cos = CosineAnnealingLR(dummy, T_max=total - args.warmup_steps, eta_min=args.lr * args.min_lr_ratio)
    sched = SequentialLR(dummy, schedulers=[warm, cos], milestones=[args.warmup_steps])
    for step in range(total):
        lrs = opt.current_lrs()
        for g, pg in zip(opt.group_names, dummy.param_groups):
            assert math.isclose(lrs[g], pg["lr"], rel_tol=2e-5, abs_tol=1e-12), (step, g, lrs[g], pg["lr"])
        dummy.step()
        sched.step()
        opt.sched_step()


def test_ssl_recipe_matches_warmup_callback(emu):
    from multimodal_edema_prediction_b200.ddp import FlatParams, FusedAdamW
    from multimodal_edema_prediction_b200.duett.duett import Model
    model = Model(pretrain=True, **KW)
    flat = FlatParams(model)
    opt = FusedAdamW.for_ssl(flat, lr=model.lr, weight_decay=model.weight_decay, warmup_steps=5)
    assert opt.max_grad_norm == 1.0 and opt.wd == 0.1 and opt.lr == 3e-4
    assert {n for n in flat.unused} == {n for n in flat.names if n.startswith("head.")}
    base, steps, decay = 3e-4, 5, 5
    for s in range(20):      # WarmUpCallback.on_train_batch_start sets the LR for step s before incrementing its counter
        want = s / steps * base if s < steps else base * (decay / (s - steps + decay)) ** 0.5
        assert math.isclose(opt.current_lrs()["rest"], want, rel_tol=1e-6, abs_tol=1e-12), (s, want)
        opt.sched_step()


def test_checkpoint_with_foreign_encoder_key_layout_is_refused(emu, tmp_path):
    from multimodal_edema_prediction_b200.duett.duett import Model
    src = Model(pretrain=False, **KW)
    sd = src.state_dict()
    assert "event_transformers.0.layers.1.1.ff.2.weight" in sd            # reference (x_transformers 1.x/2.x) layout
    good = tmp_path / "good.ckpt"
    torch.save({"state_dict": dict(sd)}, good)
    m = Model.load_from_checkpoint(str(good), pretrain=False, **KW)
    assert torch.equal(m.event_transformers[0].w2, src.event_transformers[0].w2)
    # another library vintage names the second FFN Linear ff.3: the tolerant loader must not silently keep random weights
    bad_sd = {k.replace("ff.2.", "ff.3."): v for k, v in sd.items()}
    bad = tmp_path / "bad.ckpt"
    torch.save({"state_dict": bad_sd}, bad)
    with pytest.raises(RuntimeError, match="x_transformers key layout"):
        Model.load_from_checkpoint(str(bad), pretrain=False, **KW)
    # a missing *head* key is still tolerated (duett/duett.py:459-487)
    part = {k: v for k, v in sd.items() if not k.startswith("head.4")}
    pth = tmp_path / "part.ckpt"
    torch.save({"state_dict": part}, pth)
    Model.load_from_checkpoint(str(pth), pretrain=False, **KW)


def test_bf16_shadow_validity_follows_the_version_counter(emu):
    from multimodal_edema_prediction_b200.ddp import FlatParams, FusedAdamW, shadow_of
    from multimodal_edema_prediction_b200.duett.duett import Model
    model = Model(pretrain=False, **KW)
    flat = FlatParams(model).enable_shadow()
    w = model.event_transformers[0].w1
    assert shadow_of(w) is not None and torch.equal(shadow_of(w).float(), w.detach().to(torch.bfloat16).float())
    opt = FusedAdamW(flat, lr=1e-2, weight_decay=0.0)
    flat.grad.fill_(1.0)
    opt.step()                                        # the optimiser kernel refreshes the shadow: still valid, new values
    assert shadow_of(w) is not None and torch.equal(shadow_of(w).float(), w.detach().to(torch.bfloat16).float())
    with torch.no_grad():
        w.mul_(2.0)                                   # any torch in-place write (load_state_dict, manual edit) invalidates it
    assert shadow_of(w) is None
    flat.sync_shadow()
    assert shadow_of(w) is not None and torch.equal(shadow_of(w).float(), w.detach().to(torch.bfloat16).float())
