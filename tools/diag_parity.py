"""Prints per-tensor relative errors of the golden replays on the GPU (diagnostic; not a test)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
from golden_util import load, rel
import test_parity_gpu as TP

def grads_report(module, G, top=12):
    got = TP._ref_keyed_grads(module)
    rows = []
    gscale = max(float(v.abs().max()) for v in G["grad"].values())
    for k, w in G["grad"].items():
        g = got[k]
        rows.append((rel(g, w), float((g.double()-w.double()).norm()), float(w.norm()), k))
    rows.sort(reverse=True)
    print("  gscale", gscale)
    for r in rows[:top]:
        print("  grad rel %.3e abs %.3e norm %.3e %s" % r)

def ssl(mode):
    from multimodal_edema_prediction_b200.duett.duett import Model
    G = load("g3_ssl")
    model = Model(pretrain=True, seed=42, precision=mode, **TP.KW)
    model.load_state_dict(G["param"], strict=True); model.cuda().train()
    I = G["in"]
    x = (tuple(I["x_ts"]), tuple(I["x_static"]), list(I["bin_ends"]))
    x_pre, y, mask, y_ev, y_ev_mask = model.pretrain_prep_batch(x, 6)
    print("ssl", mode, "prep equal:", torch.equal(x_pre[1].cpu(), G["out"]["xs_ts_clipped"]), torch.equal(y.cpu(), G["out"]["y"]),
          torch.equal(y_ev.cpu(), G["out"]["y_events"]), torch.equal(y_ev_mask.cpu(), G["out"]["y_events_mask"]))
    outs = model.forward(x_pre, pretrain=True)
    for got, key in zip(outs, ("y_hat_value", "y_hat_presence", "y_hat_events", "y_hat_events_presence")):
        print("  ", key, rel(got.cpu(), G["out"][key]), tuple(got.shape), tuple(G["out"][key].shape))
    model.rng = np.random.default_rng(42)
    model.load_state_dict(G["param"]); model.cuda()
    loss = model.training_step((x, tuple([0.0] * 6)), 0)
    print("  loss", float(loss), float(G["out"]["loss"]))
    loss.backward()
    grads_report(model, G)

def student(mode):
    from multimodal_edema_prediction_b200.loss.losses_duett import StudentKDLoss
    from multimodal_edema_prediction_b200.models.main_architecture_duett import DuettFeatureExtractor, StudentModel
    G = load("g1_student_kd")
    duett = DuettFeatureExtractor(pretrain=False, precision=mode, **TP.KW)
    st = StudentModel(duett, pool="mean", head_hidden=16, head_dropout=0.0)
    st.load_state_dict(G["param"], strict=True); st.cuda().train()
    I = G["in"]
    x_ts, x_static, bin_ends = tuple(I["x_ts"]), tuple(I["x_static"]), list(I["bin_ends"])
    tokens = duett.encode(duett.feats_to_input((x_ts, x_static, bin_ends), 6))
    print("student", mode, "tokens", rel(tokens.float().cpu(), G["out"]["tokens"]))
    st.load_state_dict(G["param"]); st.cuda()
    z_s = st(x_ts, x_static, bin_ends)
    print("  z_s", rel(z_s.cpu(), G["out"]["z_s"]), z_s.cpu().tolist(), G["out"]["z_s"].tolist())
    losses = StudentKDLoss(kd_T=4.0, kd_alpha=0.5, pos_weight=2.0)(z_s, I["z_t"].cuda(), I["y"].cuda())
    for k in ("total", "bce", "kd"):
        print("  ", k, float(losses[k]), float(G["out"][k]))
    losses["total"].backward()
    grads_report(st, G)

def supervised(mode):
    from multimodal_edema_prediction_b200.duett.duett import Model
    G = load("g2_supervised")
    model = Model(pretrain=False, fusion_method="rep_token", pos_frac=0.3, precision=mode, **TP.KW)
    model.load_state_dict(G["param"], strict=True); model.cuda().train()
    I = G["in"]
    lens = I["n_timesteps"].tolist()
    x_ts = tuple(I["xs_ts"][i, :n, :-1] for i, n in enumerate(lens))
    times = [I["xs_times"][i, :n] for i, n in enumerate(lens)]
    loss = model.training_step(((x_ts, tuple(I["xs_static"]), times), tuple(I["y"].tolist())), 0)
    print("supervised", mode, "loss", float(loss), float(G["out"]["loss"]))
    loss.backward()
    grads_report(model, G)

if __name__ == "__main__":
    ssl("fp32"); student("bf16"); supervised("bf16"); ssl("bf16")
