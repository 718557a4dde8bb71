"""Drop-in check at the reference's own call site: the reference's UNMODIFIED training_duett/engine.py
(imported from /root/reference — it needs no third-party package) drives the product's TeacherModel / StudentModel /
StudentKDLoss / DualPathologyLoss through its one-step functions, and the returned numbers are the golden ones produced by
the reference's own modules (tests/golden, oracle/make_golden.py).  The kernels are replaced by the torch contract emulator
(CPU container; the same modules run on the CUDA kernels in tests/test_parity_gpu.py).  Skipped where the reference tree
does not exist (the GPU box)."""
import importlib.util
import os

import pytest
import torch

import ops_emulator
from golden_util import load, rel

REF = os.environ.get("DUETT_REFERENCE", "/root/reference")
ENGINE = os.path.join(REF, "training_duett", "engine.py")
pytestmark = pytest.mark.skipif(not os.path.exists(ENGINE), reason="reference tree not present")

KW = dict(d_static_num=3, d_time_series_num=5, d_target=1, d_embedding=8, masked_transform_timesteps=4, max_len=4,
          n_duett_layers=2, d_feedforward=96)
TOL = 5e-5


@pytest.fixture
def emu(monkeypatch):
    ops_emulator.install(monkeypatch)


@pytest.fixture(scope="module")
def ref_engine():
    spec = importlib.util.spec_from_file_location("reference_training_duett_engine", ENGINE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class StubCXR(torch.nn.Module):
    d_out = 16

    def forward(self, pv):
        return pv[:, 0], pv[:, 1:]


def _teacher(G):
    from multimodal_edema_prediction_b200.models.main_architecture_duett import (DuettFeatureExtractor,
                                                                                 PatchDualPathologyPerceiver, TeacherModel)
    duett = DuettFeatureExtractor(pretrain=False, **KW)
    perceiver = PatchDualPathologyPerceiver(7, duett.d_representation, d_latent=32, n_heads=4, dropout=0.0, head_hidden=16,
                                            head_dropout=0.0)
    teacher = TeacherModel(duett, StubCXR(), perceiver, patch_dual_pathology_mode=True, d_img=16)
    teacher.load_state_dict(G["param"], strict=True)
    return teacher


def test_reference_engine_trains_the_product_teacher_to_the_golden_numbers(emu, ref_engine):
    from multimodal_edema_prediction_b200.loss.losses_duett import DualPathologyLoss
    G = load("g4_teacher")
    teacher = _teacher(G)
    I = G["in"]
    loss_fn = DualPathologyLoss(I["label_weights"], I["pos_weight"], 0.5, 0.5, 1.0)
    opt = torch.optim.SGD(teacher.parameters(), lr=0.0)
    batch = {"x_ts": tuple(I["x_ts"]), "x_static": tuple(I["x_static"]), "bin_ends": tuple(I["bin_ends"]),
             "y": torch.zeros(6), "pixel_values": I["pixel_values"], "y_multi": I["y_multi"], "y_multi_mask": I["y_multi_mask"]}
    res = ref_engine.train_teacher_dual_pathology_batch(batch, teacher, loss_fn, opt, torch.device("cpu"),
                                                        aux_residual_alpha=0.3)
    assert abs(res["loss"] - float(G["out"]["loss"])) < 1e-4 * abs(float(G["out"]["loss"]))
    assert abs(res["aux_residual"] - float(G["out"]["aux_residual"])) < 1e-4
    for k in ("main_logit", "img_logits", "ts_logits", "fusion_logits"):
        assert rel(res[k], G["out"][k]) < TOL, k
    for k in ("img_per", "ts_per", "fus_per"):
        assert rel(res[k], G["out"][k]) < TOL, k
    # and the product's engine returns the same dict (same keys, same values) as the reference's
    from multimodal_edema_prediction_b200.training_duett import engine
    teacher2 = _teacher(G)
    res2 = engine.train_teacher_dual_pathology_batch(batch, teacher2, loss_fn, torch.optim.SGD(teacher2.parameters(), lr=0.0),
                                                     torch.device("cpu"), aux_residual_alpha=0.3)
    assert set(res2) == set(res)
    for k in res:
        a, b = res[k], res2[k]
        if torch.is_tensor(a):
            assert torch.allclose(a.float(), b.float(), rtol=1e-5, atol=1e-6), k
        else:
            assert abs(a - b) <= 1e-5 * max(1.0, abs(a)), k


def test_reference_engine_distils_the_product_student_from_the_product_teacher(emu, ref_engine):
    from multimodal_edema_prediction_b200.loss.losses_duett import StudentKDLoss
    from multimodal_edema_prediction_b200.models.main_architecture_duett import DuettFeatureExtractor, StudentModel
    from oracle import duett_oracle as O
    Gs, Gt = load("g1_student_kd"), load("g4_teacher")
    student = StudentModel(DuettFeatureExtractor(pretrain=False, **KW), pool="mean", head_hidden=16, head_dropout=0.0)
    student.load_state_dict(Gs["param"], strict=True)
    teacher = _teacher(Gt)
    for p in teacher.parameters():
        p.requires_grad = False
    I, It = Gs["in"], Gt["in"]
    batch_stu = {"x_ts": tuple(I["x_ts"]), "x_static": tuple(I["x_static"]), "bin_ends": tuple(I["bin_ends"]), "y": I["y"]}
    batch_tea = dict(batch_stu, pixel_values=It["pixel_values"])
    kd = StudentKDLoss(kd_T=4.0, kd_alpha=0.5, pos_weight=2.0)
    opt = torch.optim.SGD(student.parameters(), lr=0.0)
    res = ref_engine.train_student_batch(batch_stu, batch_tea, student, teacher, kd, opt, torch.device("cpu"))
    # student logits are the golden ones (the reference's StudentModel on the same weights and batch)
    assert rel(res["logits"], Gs["out"]["z_s"]) < TOL
    # teacher logits / KD terms: the oracle (pinned against the reference's TeacherModel by g4) on this batch, eval mode
    cfg = O.DuettConfig(d_static_num=3, d_time_series_num=5, n_timesteps=4, d_embedding=8, n_layers=2, d_feedforward=96)
    P = {k[len("duett."):]: v for k, v in Gt["param"].items() if k.startswith("duett.")}
    Pt = {k: v for k, v in Gt["param"].items() if not k.startswith("duett.")}
    xs, xt, tm, _ = O.feats_to_input(I["x_ts"], I["x_static"], I["bin_ends"], cfg.T)
    with torch.no_grad():
        z_t = O.teacher_forward(P, Pt, cfg, xs, xt, tm, It["pixel_values"][:, 1:], training=False)["main_logit"]
    want = O.student_kd_loss(Gs["out"]["z_s"], z_t, I["y"], 4.0, 0.5, 2.0)
    assert abs(res["loss"] - float(want["total"])) < 1e-4 * abs(float(want["total"]))
    assert abs(res["bce"] - float(want["bce"])) < 1e-4 and abs(res["kd"] - float(want["kd"])) < 1e-4
    assert not teacher.training and student.training          # engine.py:279-280 semantics
