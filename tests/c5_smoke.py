"""BASELINE config 5 shape (T=128, V=512, d=256, L=2, heads 2 -> dh=128: SIMT attention path, 131k-wide time tokens) at a
small batch: forward + backward through the nn.Module API in fp32 and bf16 modes on the GPU; checks finiteness and that
the two precisions agree (bf16 bound).  usage: python tests/c5_smoke.py [B]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import duett_oracle as O          # synthetic inputs + initial weights only
from multimodal_edema_prediction_b200.models.main_architecture_duett import DuettFeatureExtractor, StudentModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = O.DuettConfig(d_static_num=24, d_time_series_num=512, n_timesteps=128, d_embedding=256, n_layers=2)
P, H = O.init_params(cfg, seed=1), O.init_student_head(cfg, seed=2)
batch = O.synth_batch(cfg, B, seed=99)
x = (batch["x_ts"], batch["x_static"], list(batch["bin_ends"]))
res = {}
for mode in ("fp32", "bf16"):
    duett = DuettFeatureExtractor(cfg.d_static_num, cfg.V, 1, d_embedding=cfg.d_embedding, n_duett_layers=cfg.n_layers,
                                  masked_transform_timesteps=cfg.T, max_len=cfg.T, d_feedforward=cfg.d_feedforward,
                                  pretrain=False, precision=mode)
    student = StudentModel(duett, pool="mean", head_hidden=128, head_dropout=0.0)
    sd = {"duett." + k: v for k, v in P.items()}
    sd.update(H)
    student.load_state_dict(sd, strict=True)
    student.cuda().train()
    torch.cuda.synchronize(); t0 = time.time()
    z = student(*x)
    z.sum().backward()
    torch.cuda.synchronize()
    g = torch.cat([p.grad.flatten().float() for p in student.parameters() if p.grad is not None])
    assert torch.isfinite(z).all() and torch.isfinite(g).all(), mode
    res[mode] = (z.detach().float().cpu(), g.cpu())
    print(f"{mode}: {time.time() - t0:.2f} s  |z| {float(z.abs().mean()):.4f}  |grad| {float(g.norm()):.4e}  "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del student, duett, z, g
    torch.cuda.empty_cache()
rz = float((res["bf16"][0] - res["fp32"][0]).norm() / res["fp32"][0].norm())
rg = float((res["bf16"][1] - res["fp32"][1]).norm() / res["fp32"][1].norm())
print(f"bf16 vs fp32: logits rel {rz:.3e}  gradients rel {rg:.3e}")
assert rg < 5e-2, rg
print("C5 SHAPE OK")
