"""Run-to-run determinism probe: the same student step (fwd + KD loss + bwd) N times on identical inputs; prints the largest
relative deviation of tokens / logits / every gradient from run 0.  Atomic accumulation order gives ~1e-6 (fp32) or a few
1e-3 (bf16 rounding flips); anything larger points at a race.
usage: python tests/flake_check.py [N] [mode] [d] [L] [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import duett_oracle as O      # synthetic inputs + initial weights only
from multimodal_edema_prediction_b200.loss.losses_duett import StudentKDLoss
from multimodal_edema_prediction_b200.models.main_architecture_duett import DuettFeatureExtractor, StudentModel

N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
d = int(sys.argv[3]) if len(sys.argv) > 3 else 64
L = int(sys.argv[4]) if len(sys.argv) > 4 else 2
B = int(sys.argv[5]) if len(sys.argv) > 5 else 4
cfg = O.DuettConfig(d_static_num=24, d_time_series_num=128, n_timesteps=32, d_embedding=d, n_layers=L)
P, H = O.init_params(cfg, seed=3), O.init_student_head(cfg, seed=4)
batch = O.synth_batch(cfg, B, seed=1237)
duett = DuettFeatureExtractor(cfg.d_static_num, cfg.V, 1, d_embedding=d, n_duett_layers=L, masked_transform_timesteps=cfg.T,
                              max_len=cfg.T, d_feedforward=cfg.d_feedforward, pretrain=False, precision=mode)
student = StudentModel(duett, pool="rep_token", head_hidden=128, head_dropout=0.0)
sd = {"duett." + k: v for k, v in P.items()}
sd.update(H)
x = (batch["x_ts"], batch["x_static"], list(batch["bin_ends"]))
z_t = (torch.randn(B, generator=torch.Generator().manual_seed(5)) * 1.5).cuda()
y = batch["y"].cuda()
crit = StudentKDLoss(kd_T=4.0, kd_alpha=0.5)


def run():
    student.load_state_dict(sd, strict=True)
    student.cuda().train()
    for p in student.parameters():
        p.grad = None
    tokens = duett.encode(duett.feats_to_input(x, B)).float()
    student.load_state_dict(sd)
    z = student(*x)
    crit(z, z_t, y)["total"].backward()
    out = {"tokens": tokens.detach().clone(), "z": z.detach().float().clone()}
    for n, p in student.named_parameters():
        if p.grad is not None:
            out["g:" + n] = p.grad.detach().float().clone()
    return out


ref = run()
if os.environ.get("FLAKE_ORACLE"):      # distribution of the parity test's logit / token error against the CPU oracle
    xs, xt, tm, _ = O.feats_to_input(batch["x_ts"], batch["x_static"], batch["bin_ends"], cfg.T)
    with torch.no_grad():
        z_ref = O.student_forward(P, H, cfg, xs, xt, tm, pool="rep_token")
        tok_ref = O.encode(P, cfg, xs, xt, tm, training=True)
    rz, rt = [], []
    for i in range(N):
        cur = run()
        rz.append(float((cur["z"].cpu() - z_ref).norm() / z_ref.norm()))
        rt.append(float((cur["tokens"].cpu() - tok_ref).norm() / tok_ref.norm()))
    print("rel(z)      min/median/max", min(rz), sorted(rz)[len(rz) // 2], max(rz))
    print("rel(tokens) min/median/max", min(rt), sorted(rt)[len(rt) // 2], max(rt))
    sys.exit(0)
gscale = max(float(v.abs().max()) for k, v in ref.items() if k.startswith("g:"))
TOL, FLOOR = (0.05, 5e-2) if mode == "bf16" else (1e-3, 5e-2)    # the parity test's criterion (tests/test_parity_gpu.py)
worst_all = 0.0
for i in range(1, N):
    cur = run()
    worst, wk = 0.0, ""
    for k, v in ref.items():
        bound = TOL * float(v.norm()) + FLOOR * TOL * gscale * v.numel() ** 0.5 + 1e-7
        e = float((cur[k] - v).norm()) / bound
        if not (e <= worst):      # also catches NaN
            worst, wk = e, k
    worst_all = max(worst_all, worst) if worst == worst else float("nan")
    print(f"run {i:3d}: worst deviation / test bound = {worst:.3e} at {wk}", flush=True)
print("WORST", worst_all)
