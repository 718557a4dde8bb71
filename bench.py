#!/usr/bin/env python
"""Benchmark of the DuETT hot path (BASELINE.json metric: DuETT train samples/sec at 1/2/4/8 B200).

  python bench.py [--gpus N --steps K --warmup W] [--config c2]     this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference [--config c2] [...]               the reference's algorithm on the host CPU cores

Workloads = BASELINE.json `configs` (default: configs[1], the one the metric is quoted on):
  c1  configs[0]  DuETT small (d=64, 2+2 layers, T=32, V=128) supervised step, B=32/GPU            (the CPU-runnable case)
  c2  configs[1]  DuETT base  (d=128, 4+4 layers, T=32, V=128) supervised edema head, bf16, B=256/GPU      [weak scaling]
  c3  configs[2]  SSL pre-training step of the base model (masked value / event reconstruction, host numpy-RNG masking,
                  AdamW 3e-4 / wd 0.1 / clip-norm 1.0), GLOBAL batch 1024 split over the ranks           [strong scaling]
  c4  configs[3]  KD step: frozen patch_dual teacher (DuETT + CXR-embedding perceiver fusion on synthetic [B,1+1369,768]
                  RAD-DINO embeddings) forward under no_grad + student forward/backward + StudentKDLoss + AdamW with the
                  trainer's LR groups, GLOBAL batch 512 (64/GPU on 8)                                      [strong scaling]
  c5  configs[4]  stress shape T=128, V=512, d=256 (2+2 layers) supervised step, GLOBAL batch 128          [strong scaling]
One step = forward + loss + backward (+ bucketed gradient all-reduce over NCCL for N>1, launched per weight group from
inside backward) + fused AdamW.  Data are synthetic MIMIC-shaped tensors, weights are random-init.

value  : samples/s with the step's inputs already resident in HBM (whole step replayed from a CUDA graph).
e2e    : samples/s through the host-facing API: every step stacks the per-sample host tensors of the collate format into
         pinned staging memory, copies them to the device (Model.feats_to_input; c3 additionally runs the host-RNG masking
         of Model.pretrain_prep_batch, c4 uploads the [B,1370,768] CXR embeddings) and reads the loss back.
roofline: all tcgen05 GEMM launches of the timed region, timed with CUDA events on the launching stream in an eager pass
         of the same step (achieved = their algorithmic FLOPs / their summed duration) against the measured sustained
         bf16 peak.
parity_check: after the timed region the graph-replayed bf16 step is compared, on one batch and identical weights, with
         the same step run eagerly in fp32 mode (exact-parity FFMA kernels): loss and global gradient norm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "duett_train_samples_per_sec"
BASE = dict(d_static_num=24, V=128, T=32, d=128, L=4, heads=2, d_ff=512)
CONFIGS = {
    "c1": dict(dims=dict(BASE, d=64, L=2), task="supervised", B=32, scaling="weak", cfg_index=0,
               workload="DuETT small (d=64, 2+2 layers, T=32, V=128) supervised step, bf16, B=32/GPU (BASELINE.json configs[0])"),
    "c2": dict(dims=dict(BASE), task="supervised", B=256, scaling="weak", cfg_index=1,
               workload="DuETT base (d=128, 4+4 layers, T=32, V=128) supervised edema head, bf16, B=256/GPU "
                        "(BASELINE.json configs[1]); fwd+loss+bwd+allreduce+AdamW"),
    "c3": dict(dims=dict(BASE), task="ssl", B=1024, scaling="strong", cfg_index=2,
               workload="DuETT base SSL pre-training step (masked value/event reconstruction, train_duett_ssl recipe), bf16, "
                        "global B=1024 (BASELINE.json configs[2]); fwd+loss+bwd+allreduce+clip+AdamW"),
    "c4": dict(dims=dict(BASE), task="kd", B=512, scaling="strong", cfg_index=3,
               workload="KD student step with frozen patch_dual teacher + CXR-embedding fusion (DuETT base both), bf16, "
                        "global B=512 (BASELINE.json configs[3]); teacher fwd + student fwd+loss+bwd+allreduce+AdamW"),
    "c5": dict(dims=dict(BASE, T=128, V=512, d=256, L=2), task="supervised", B=128, scaling="strong", cfg_index=4,
               workload="DuETT stress shape (T=128, V=512, d=256, 2+2 layers) supervised step, bf16, global B=128 "
                        "(BASELINE.json configs[4]); fwd+loss+bwd+allreduce+AdamW"),
}
N_PATCH, D_IMG, D_LATENT, K_PATH = 1369, 768, 256, 7


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"bf16_tflops_sustained": j.get("bf16_tflops_sustained", 1400.0), "bf16_tflops": j.get("bf16_tflops", 1590.0),
                "hbm_gbs": j.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_tflops_sustained": 1400.0, "bf16_tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


def fwd_flops_per_sample(w):
    """fwd = L*[4*T1*V1*d*(4d+2F) + 4*d*(V1^2+T1^2)] + 2*T*V*(2*64+64*d)."""
    T1, V1, d, F, L = w["T"] + 1, w["V"] + 1, w["d"], w["d_ff"], w["L"]
    return L * (4 * T1 * V1 * d * (4 * d + 2 * F) + 4 * d * (V1 * V1 + T1 * T1)) + 2 * w["T"] * w["V"] * (2 * 64 + 64 * d)


def train_flops_per_sample(w, task="supervised"):
    f = 3.0 * fwd_flops_per_sample(w)
    if task == "kd":      # + frozen teacher forward: backbone + img_proj over the patches + K/V projection of the image cross-attention
        f += fwd_flops_per_sample(w) + 2.0 * N_PATCH * D_IMG * D_LATENT + 2.0 * N_PATCH * D_LATENT * 2 * D_LATENT
    return f


def config_dict(cfg, world, B):
    """The `config` object of the JSON line — identical for the b200 arm and the reference arm."""
    return {"workload": cfg["workload"], "config": cfg["key"], "task": cfg["task"], "per_gpu_batch": B,
            "global_batch": world * B, "parallelism": f"dp{world}", **cfg["dims"]}


def synth_host_batch(B, seed, w, pin=True, with_cxr=False):
    """Collate-format batch on the host: tuples of per-sample tensors (duett/mimic_dataset.py:83,93-95)."""
    from multimodal_edema_prediction_b200.synth import synth_batch
    b = synth_batch(w["d_static_num"], w["V"], w["T"], B, seed)
    if with_cxr:          # RAD-DINO embeddings ride in the pixel_values slot: [B, 1 + 1369, 768]
        b["pixel_values"] = torch.randn(B, 1 + N_PATCH, D_IMG, generator=torch.Generator().manual_seed(seed + 1))
    if pin:
        b = {k: (tuple(t.pin_memory() for t in v) if isinstance(v, tuple) else v.pin_memory()) for k, v in b.items()}
    return b


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation on the host cores
# ------------------------------------------------------------------------------------------------------------------
class _StubCXR(torch.nn.Module):
    """CXREncoder stand-in (frozen RAD-DINO is out of scope): the embeddings arrive in the pixel_values slot."""
    d_out = D_IMG

    def forward(self, pv):
        return pv[:, 0], pv[:, 1:]


def _reference_in_place(cfg, Bs):
    """The reference's OWN files (duett/duett.py, models/main_architecture_duett.py, loss/losses_duett.py) imported
    unmodified from /root/reference through oracle/shims, exactly as oracle/make_golden.py does.  Only possible where the
    reference tree exists (the authoring container); returns a step() closure or None."""
    ref = os.environ.get("DUETT_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "duett")):
        return None
    sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), ref]
    try:
        from duett.duett import Model
        from loss.losses_duett import StudentKDLoss
        from models.main_architecture_duett import (DuettFeatureExtractor, PatchDualPathologyPerceiver, StudentModel,
                                                    TeacherModel)
    except Exception as ex:          # noqa: BLE001
        print("reference import failed:", repr(ex)[:200], file=sys.stderr)
        return None
    w, task = cfg["dims"], cfg["task"]
    kw = dict(d_static_num=w["d_static_num"], d_time_series_num=w["V"], d_target=1, d_embedding=w["d"], n_duett_layers=w["L"],
              masked_transform_timesteps=w["T"], max_len=w["T"], d_feedforward=w["d_ff"], n_transformer_head=w["heads"])
    torch.manual_seed(0)
    b = synth_host_batch(Bs, 1234, w, pin=False, with_cxr=task == "kd")
    x = (b["x_ts"], b["x_static"], list(b["bin_ends"]))
    if task == "supervised":
        model = Model(pretrain=False, fusion_method="rep_token", pos_frac=0.3, lr=1e-4, weight_decay=1e-5, **kw).train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
        y = tuple(b["y"].tolist())

        def step():
            opt.zero_grad(set_to_none=True)
            loss = model.training_step((x, y), 0)
            loss.backward()
            opt.step()
            return float(loss)
    elif task == "ssl":
        model = Model(pretrain=True, seed=42, **kw).train()
        opt = torch.optim.AdamW(model.parameters(), lr=3e-4, weight_decay=0.1)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = model.training_step((x, tuple([0.0] * Bs)), 0)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            return float(loss)
    else:
        student = StudentModel(DuettFeatureExtractor(pretrain=False, **kw), pool="mean").train()
        td = DuettFeatureExtractor(pretrain=False, **kw)
        teacher = TeacherModel(td, _StubCXR(), PatchDualPathologyPerceiver(K_PATH, td.d_representation, D_LATENT, 4),
                               patch_dual_pathology_mode=True, d_img=D_IMG).eval()
        for p in teacher.parameters():
            p.requires_grad = False
        kd = StudentKDLoss()
        opt = torch.optim.AdamW(student.parameters(), lr=1e-4, weight_decay=0.05)

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.no_grad():
                z_t = teacher(*x, b["pixel_values"])["main_logit"]
            loss = kd(student(*x), z_t, b["y"])["total"]
            loss.backward()
            opt.step()
            return float(loss)
    return step


def _oracle_port(cfg, Bs):
    """oracle/duett_oracle.py (the CPU restatement, pinned against the reference's files by tests/golden)."""
    from oracle import duett_oracle as O
    w, task = cfg["dims"], cfg["task"]
    oc = O.DuettConfig(d_static_num=w["d_static_num"], d_time_series_num=w["V"], n_timesteps=w["T"], d_embedding=w["d"],
                       n_layers=w["L"], d_feedforward=w["d_ff"])
    b = synth_host_batch(Bs, 1234, w, pin=False, with_cxr=task == "kd")
    leaf = lambda P: {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in P.items()}
    Pl = leaf(O.init_params(oc, seed=0))
    xs, xt, tm, n_ts = O.feats_to_input(b["x_ts"], b["x_static"], b["bin_ends"], oc.T)
    train = [v for v in Pl.values() if torch.is_tensor(v) and v.requires_grad]
    if task == "supervised":
        opt = torch.optim.AdamW(train, lr=1e-4, weight_decay=1e-5)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = O.supervised_loss(O.model_forward_supervised(Pl, oc, xs, xt, tm, "rep_token"), b["y"], 0.3)
            loss.backward()
            opt.step()
            return float(loss)
    elif task == "ssl":
        import numpy as np
        opt = torch.optim.AdamW(train, lr=3e-4, weight_decay=0.1)
        rng = np.random.default_rng(42)

        def step():
            opt.zero_grad(set_to_none=True)
            x_m, y, mask, y_ev, y_ev_mask = O.pretrain_prep_batch(rng, oc, xt, n_ts, pretrain_dropout=0.5)
            loss = O.ssl_loss(*O.model_forward_pretrain(Pl, oc, xs, x_m, tm, training=True), y, mask, y_ev, y_ev_mask)
            loss.backward()
            torch.nn.utils.clip_grad_norm_([p for p in train if p.grad is not None], 1.0)
            opt.step()
            return float(loss)
    else:
        Hl = leaf(O.init_student_head(oc, seed=1))
        Pt, Ht = O.init_params(oc, seed=2), O.init_teacher_head(oc, seed=3, K=K_PATH, d_latent=D_LATENT, d_img=D_IMG)
        opt = torch.optim.AdamW(train + list(Hl.values()), lr=1e-4, weight_decay=0.05)

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.no_grad():
                z_t = O.teacher_forward(Pt, Ht, oc, xs, xt, tm, b["pixel_values"][:, 1:], training=False)["main_logit"]
            loss = O.student_kd_loss(O.student_forward(Pl, Hl, oc, xs, xt, tm, pool="mean"), z_t, b["y"])["total"]
            loss.backward()
            opt.step()
            return float(loss)
    return step


def run_reference(args, cfg):
    """The reference's CPU path on the box's host cores, same workload / metric / config object as the b200 arm; each step
    is a bounded sample of the per-GPU batch.  kind = "reference-in-place" when /root/reference is present (its own
    files, unmodified, through oracle/shims), else "port" (oracle/duett_oracle.py — /root/reference does not exist on
    the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = args.batch or (cfg["B"] if cfg["scaling"] == "weak" else max(1, cfg["B"] // max(args.gpus, 1)))
    Bs = min(args.cpu_sample or {"c1": 32, "c5": 2}.get(cfg["key"], 16), B)
    step, kind = None, "port"
    if not args.force_port:
        step = _reference_in_place(cfg, Bs)
        kind = "reference-in-place" if step is not None else "port"
    if step is None:
        step = _oracle_port(cfg, Bs)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step()
    dt = (time.perf_counter() - t0) / args.steps
    val = Bs / dt
    sample = (f"{Bs} of the {B} samples of one per-GPU step ({cfg['task']}: fwd+loss+bwd+AdamW), fp32, torch CPU, "
              f"{'reference files via oracle/shims' if kind != 'port' else 'oracle port'}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_dict(cfg, max(args.gpus, 1), B),
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "last_loss": loss,
    }))


# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm = sorted(float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        if sm:
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = float(rows[0][2])
            out["samples"] = len(sm)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for j, n in enumerate(names):
                if any(len(r) >= 9 and r[5 + j].strip().lower().startswith("active") for r in rows):
                    out["reasons"].append(n)
        return out


# ------------------------------------------------------------------------------------------------------------------
# the three step kinds behind one interface
# ------------------------------------------------------------------------------------------------------------------
class Workload:
    """model(s) + optimiser + reducer for one task; `static_from(host_batch)` -> dict of device tensors (the step's inputs),
    `loss_fn(static)` -> scalar loss (forward only); `train_step` adds backward / all-reduce / optimiser."""

    def __init__(self, cfg, device, precision="bf16", shadow=True, attach=True):
        from multimodal_edema_prediction_b200.ddp import FlatParams, FusedAdamW, GradReducer
        from multimodal_edema_prediction_b200.duett.duett import Model
        self.cfg, self.device, self.task = cfg, device, cfg["task"]
        w = self.w = cfg["dims"]
        kw = dict(d_embedding=w["d"], n_duett_layers=w["L"], masked_transform_timesteps=w["T"], max_len=w["T"],
                  d_feedforward=w["d_ff"], n_transformer_head=w["heads"], precision=precision)
        torch.manual_seed(0)
        self.teacher, self.n_steps = None, None
        if self.task == "supervised":
            self.model = Model(w["d_static_num"], w["V"], 1, pretrain=False, fusion_method="rep_token", pos_frac=0.3, lr=1e-4,
                               weight_decay=1e-5, **kw).to(device).train()
        elif self.task == "ssl":
            self.model = Model(w["d_static_num"], w["V"], 1, pretrain=True, seed=42, **kw).to(device).train()
        else:
            from multimodal_edema_prediction_b200.loss.losses_duett import StudentKDLoss
            from multimodal_edema_prediction_b200.models.main_architecture_duett import (
                DuettFeatureExtractor, PatchDualPathologyPerceiver, StudentModel, TeacherModel)
            sd = DuettFeatureExtractor(w["d_static_num"], w["V"], 1, pretrain=False, **kw)
            self.model = StudentModel(sd, pool="mean").to(device).train()
            td = DuettFeatureExtractor(w["d_static_num"], w["V"], 1, pretrain=False, **kw)
            self.teacher = TeacherModel(td, _StubCXR(), PatchDualPathologyPerceiver(K_PATH, td.d_representation, D_LATENT, 4),
                                        patch_dual_pathology_mode=True, d_img=D_IMG).to(device).eval()
            for p in self.teacher.parameters():
                p.requires_grad = False
            self.kd = StudentKDLoss().to(device)
        self.trainable = self.model
        self.flat = FlatParams(self.trainable, broadcast=attach)
        if shadow and precision == "bf16":
            self.flat.enable_shadow()          # bf16 weight shadows refreshed by the optimizer kernel (no cast launches)
        if self.task == "ssl":                 # duett/duett.py:325-327 + train_duett_ssl.py:27-50,191
            self.opt = FusedAdamW.for_ssl(self.flat, lr=3e-4, weight_decay=0.1, warmup_steps=2000)
            self.opt.set_lr_scale(1.0)         # measure at the full learning rate (step 0 of the warm-up is lr = 0)
        elif self.task == "kd":                # training_duett/trainer.py:77-125,902
            import types
            a = types.SimpleNamespace(lr=1e-4, backbone_lr_mult=0.2, query_lr_mult=0.2, correction_lr_mult=1.0,
                                      weight_decay=0.05, warmup_steps=300, min_lr_ratio=0.01)
            self.opt = FusedAdamW.from_trainer_args(self.flat, a)
        else:
            self.opt = FusedAdamW(self.flat, lr=1e-4, weight_decay=1e-5)
        self.red = GradReducer(self.flat)
        if attach:
            self.red.attach()

    # ---- host batch -> device tensors ---------------------------------------------------------------------------
    def static_from(self, hb, B):
        m = self.model if self.task != "kd" else self.model.duett
        x = (hb["x_ts"], hb["x_static"], list(hb["bin_ends"]))
        if self.task == "ssl":
            (xs_static, x_c, xs_times, n_ts), y, mask, y_ev, y_ev_mask = m.pretrain_prep_batch(x, B)
            self.n_steps = n_ts
            return {"xs_static": xs_static, "xs_ts": x_c, "xs_times": xs_times, "y": y, "mask": mask, "y_ev": y_ev,
                    "y_ev_mask": y_ev_mask}
        xs_static, xs_ts, xs_times, n_ts = m.feats_to_input(x, B)
        self.n_steps = n_ts
        out = {"xs_static": xs_static, "xs_ts": xs_ts, "xs_times": xs_times, "y": hb["y"].to(self.device, non_blocking=True)}
        if self.task == "kd":
            out["pix"] = hb["pixel_values"].to(self.device, non_blocking=True)
        return out

    @staticmethod
    def h2d_bytes(st):
        return sum(t.numel() * t.element_size() for t in st.values())

    # ---- forward + loss on device-resident inputs ---------------------------------------------------------------
    def loss_fn(self, st, z_t=None):
        if self.task == "supervised":
            y_hat = self.model.forward((st["xs_static"], st["xs_ts"], st["xs_times"], self.n_steps))
            return self.model._supervised_loss(y_hat, st["y"])
        if self.task == "ssl":
            outs = self.model.forward((st["xs_static"], st["xs_ts"], st["xs_times"], self.n_steps), pretrain=True)
            return self.model._ssl_loss(outs, st["y"], st["mask"], st["y_ev"], st["y_ev_mask"])
        # kd: the body of training_duett/engine.py:270-301 on tensors that are already on the device; the per-sample tuples the
        # module API takes are views of the static tensors (feats_to_input re-appends the mask column on the device)
        x_ts = tuple(st["xs_ts"][:, :, :-1].unbind(0))
        x_static, bin_ends = tuple(st["xs_static"].unbind(0)), tuple(st["xs_times"].unbind(0))
        if z_t is None:
            with torch.no_grad():
                z_t = self.teacher(x_ts, x_static, bin_ends, st["pix"])["main_logit"]
        self.last_z_t = z_t
        z_s = self.model(x_ts, x_static, bin_ends)
        return self.kd(z_s, z_t, st["y"])["total"]

    def train_step(self, st, optimizer=True):
        self.opt.zero_grad()
        self.red.start_step()
        loss = self.loss_fn(st)
        loss.backward()
        scale = self.red.finish()
        if optimizer:
            self.opt.step(grad_scale=scale)
        return loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32", "fp32"],
                    help="bf16 (default: tcgen05 kind::f16), tf32 (fp32 storage, tcgen05 kind::tf32: the reference's SSL precision), "
                         "fp32 (exact FFMA parity mode)")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (experiments; the line then says so)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="samples per CPU-baseline step (default 16; c1: 32; c5: 2)")
    ap.add_argument("--force-port", action="store_true", help="reference arm: use the oracle port even if /root/reference exists")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--profile-out", default=None, help="write the per-shape GEMM timing table to this JSON file")
    ap.add_argument("--comm-trace", default=None, help="N>1: also run traced eager steps and write the per-bucket all-reduce "
                                                       "timeline (ready / start / end vs the backward) to this JSON file")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config], key=args.config)
    if args.impl == "reference":
        return run_reference(args, cfg)
    args.warmup = max(args.warmup, 3)

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: native libraries that print to fd 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from multimodal_edema_prediction_b200 import ops
    if cfg["scaling"] == "strong" and cfg["B"] % world:
        raise SystemExit(f"global batch {cfg['B']} does not split over {world} ranks")
    B = args.batch or (cfg["B"] if cfg["scaling"] == "weak" else cfg["B"] // world)
    wl = Workload(cfg, device, precision=args.precision)
    w = cfg["dims"]
    nb = 4 if cfg["key"] != "c5" else 2                # distinct synthetic batches, cycled
    host = [synth_host_batch(B, 1234 + 17 * rank + i, w, with_cxr=wl.task == "kd") for i in range(nb)]
    dev_batches = [wl.static_from(hb, B) for hb in host]
    torch.cuda.synchronize()
    static = {k: v.clone() for k, v in dev_batches[0].items()}

    def step_eager(i):
        return wl.train_step(dev_batches[i % nb])

    # ---- whole-step CUDA graph (forward + loss + backward + all-reduce + AdamW), inputs copied into static tensors -----
    gstep, gstep_noopt, graph_err, launches_per_graph = None, None, None, 0
    if not args.no_graph:
        try:
            from multimodal_edema_prediction_b200.graph import CudaGraphStep
            l0 = ops.launches()
            gstep = CudaGraphStep(lambda: wl.train_step(static), static, warmup=max(args.warmup, 3))
            launches_per_graph = (ops.launches() - l0) // (max(args.warmup, 3) + 1)
            # SURVEY §8(d) also asks for the step WITHOUT the optimizer (fwd + loss + bwd + all-reduce): a second graph
            # (same memory pool: the two graphs are never replayed concurrently, and at the stress shape a second private
            # pool of saved activations would not fit next to the eager profiling pass)
            torch.cuda.empty_cache()
            gstep_noopt = CudaGraphStep(lambda: wl.train_step(static, optimizer=False), static, warmup=0, pool=gstep.graph.pool())
        except Exception as ex:          # capture unsupported in this configuration: run eagerly and say so
            import traceback
            traceback.print_exc(file=sys.stderr)
            graph_err = repr(ex)[:200]
            torch.cuda.synchronize()

    def step_resident(i):
        return gstep(**dev_batches[i % nb]) if gstep is not None else step_eager(i)

    def step_noopt(i):
        return gstep_noopt(**dev_batches[i % nb])

    # e2e: every step copies ITS inputs host -> device (collate format -> pinned staging -> async H2D on a copy stream, so the
    # transfer overlaps the previous step still running on the compute stream) and ITS loss device -> host.  The loss of
    # step i is read while step i+1 is already enqueued (pinned double buffer + event), the way a training loop logs.
    loss_pin = [torch.empty(1, dtype=torch.float64).pin_memory() for _ in range(2)]
    loss_evt = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"pending": None, "losses": []}
    copy_stream = torch.cuda.Stream(device=device)

    def e2e_collect():
        j = e2e_state["pending"]
        if j is not None:
            loss_evt[j].synchronize()
            e2e_state["losses"].append(float(loss_pin[j][0]))
            e2e_state["pending"] = None

    def step_e2e(i):
        hb = host[i % nb]
        main = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            st = wl.static_from(hb, B)
        main.wait_stream(copy_stream)
        for t_ in st.values():
            t_.record_stream(main)
        loss = gstep(**st) if gstep is not None else wl.train_step(st)
        j = i & 1
        loss_pin[j].copy_(loss.detach().reshape(1).double(), non_blocking=True)      # device -> host read of the step's result
        loss_evt[j].record()
        e2e_collect()                                                       # previous step's loss
        e2e_state["pending"] = j

    step_e2e.drain = e2e_collect

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        barrier()
        l0 = ops.launches()
        if profile:
            ops.PROFILE = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h0 = time.perf_counter()
        for i in range(steps):
            fn(i)
        if hasattr(fn, "drain"):
            fn.drain()                                                 # last outstanding read-back, inside the timed region
        timed.host_ms = (time.perf_counter() - h0) * 1e3 / steps     # host enqueue time per step (no sync inside)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        prof, ops.PROFILE = ops.PROFILE, None
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms, ops.launches() - l0, prof

    for i in range(args.warmup):
        step_resident(i)
    clocks = ClockSampler(local) if rank == 0 else None
    ms, launches, _ = timed(step_resident, args.steps)
    host_ms = timed.host_ms
    clk = clocks.stop() if clocks else None
    if gstep is not None:
        launches = launches_per_graph * args.steps       # kernels replayed from the graph in the timed region
    for i in range(3):
        step_e2e(i)
    e2e_collect()
    e2e_state["losses"].clear()
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    ms_noopt = None
    if gstep_noopt is not None:
        for i in range(3):
            step_noopt(i)
        ms_noopt, _, _ = timed(step_noopt, args.steps)
    assert len(e2e_state["losses"]) == args.steps and all(l == l for l in e2e_state["losses"]), "e2e: a loss was not read back"

    # ---- all-reduce overlap evidence: per-bucket events of a few traced eager steps ------------------------------------------
    comm = None
    if args.comm_trace and world > 1:
        wl.red.trace = True
        steps_tl = []
        for i in range(4):
            ev_b0, ev_b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            st_ = dev_batches[i % nb]
            wl.opt.zero_grad()
            wl.red.start_step()
            loss_ = wl.loss_fn(st_)
            ev_b0.record()
            loss_.backward()
            ev_b1.record()
            scale_ = wl.red.finish()
            ev_f = torch.cuda.Event(enable_timing=True)
            ev_f.record()
            wl.opt.step(grad_scale=scale_)
            torch.cuda.synchronize()
            t0 = wl.red._t0
            steps_tl.append({"backward_start_ms": t0.elapsed_time(ev_b0), "backward_end_ms": t0.elapsed_time(ev_b1),
                             "allreduce_all_done_ms": t0.elapsed_time(ev_f),
                             "buckets": [{"MB": mb, "ready_ms": r, "start_ms": s_, "end_ms": e_} for mb, r, s_, e_ in wl.red.timeline()]})
        wl.red.trace = False
        last = steps_tl[-1]
        exposed = last["allreduce_all_done_ms"] - last["backward_end_ms"]
        comm = {"steps": steps_tl, "exposed_after_backward_ms": exposed, "n_buckets": len(last["buckets"]),
                "total_MB": sum(b["MB"] for b in last["buckets"])}
        if rank == 0:
            os.makedirs(os.path.dirname(os.path.abspath(args.comm_trace)), exist_ok=True)
            json.dump(comm, open(args.comm_trace, "w"), indent=1)
        comm = {k: v for k, v in comm.items() if k != "steps"}

    # ---- parity of the measured configuration: graph-replayed bf16 step vs the eager fp32-mode step, same weights/batch ----
    parity = None
    if not args.no_parity_check and world == 1 and args.precision == "bf16":      # single-process check (it builds a replica and steps it without collectives)
        parity = parity_check(cfg, wl, gstep, dev_batches[0], device)

    # the captured graphs hold their own pool of saved activations (86 GB at the stress shape): release it before the eager
    # passes below, which need the same amount again
    gstep_was = gstep is not None
    gstep = gstep_noopt = None
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    # per-GEMM CUDA-event timing needs eager launches (events cannot bracket nodes of a replayed graph)
    esteps = max(2, min(args.steps, 10))
    for i in range(2):
        step_eager(i)
    ms_eager, _, prof = timed(step_eager, esteps, profile=True)

    if rank == 0:
        pk = peaks()
        if args.precision == "tf32":      # kind::tf32 runs at half the bf16 tensor rate (nominal 1.1 vs 2.25 PFLOP/s dense)
            pk = dict(pk, bf16_tflops_sustained=pk["bf16_tflops_sustained"] / 2, bf16_tflops=pk["bf16_tflops"] / 2,
                      source=pk["source"] + " bf16 peak / 2 (tf32 tensor rate)")
        value = world * B * args.steps / (ms / 1e3)
        e2e = world * B * args.steps / (ms_e2e / 1e3)
        # ---- roofline of the tcgen05 GEMM kernel (dominant kernel family) ---------------------------------------------
        tc = [(t, s, f, b, a.elapsed_time(z)) for (t, s, f, b, a, z) in prof if t == "tc"]
        tot_ms = sum(r[4] for r in tc)
        tot_fl = sum(r[2] for r in tc)
        ach = tot_fl / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0
        by_shape = {}
        for t, s, f, b, m_ in tc:
            d = by_shape.setdefault(s, [0, 0.0, 0.0, 0.0])
            d[0] += 1; d[1] += f; d[2] += b; d[3] += m_
        sol = lambda fl_, by_: max(fl_ / (pk["bf16_tflops_sustained"] * 1e12), by_ / (pk["hbm_gbs"] * 1e9))
        table = sorted(({"shape": s, "launches_per_step": v[0] / esteps, "us_per_launch": v[3] / v[0] * 1e3, "ms_per_step": v[3] / esteps,
                         "tflops": v[1] / (v[3] * 1e-3) / 1e12, "gbs": v[2] / (v[3] * 1e-3) / 1e9,
                         "frac_of_speed_of_light": sol(v[1], v[2]) / (v[3] * 1e-3)}
                        for s, v in by_shape.items()), key=lambda r: -r["ms_per_step"])
        # ---- the other half of the metric: the dual-axis attention kernels -----------------------------------------------
        attn = {}
        for (t, s_, f, b, a, z) in prof:
            if t.startswith("attn"):
                d = attn.setdefault((t, s_), [0, 0.0, 0.0, 0.0])
                d[0] += 1; d[1] += f; d[2] += b; d[3] += a.elapsed_time(z)
        attn_rec = {"kernels": [{"kind": k[0], "shape": k[1], "launches_per_step": v[0] / esteps, "us_per_launch": v[3] / v[0] * 1e3,
                                 "tflops": v[1] / (v[3] * 1e-3) / 1e12, "gbs": v[2] / (v[3] * 1e-3) / 1e9} for k, v in sorted(attn.items())],
                    "ms_per_step": sum(v[3] for v in attn.values()) / esteps,
                    "share_of_step": sum(v[3] for v in attn.values()) / esteps / (ms_eager / esteps),
                    "tensor_pipe_util_ncu": _attn_ncu()}
        if args.profile_out:
            os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
            ffma = {}
            for (t, s, f, b, a, z) in prof:
                if t == "ffma":
                    d = ffma.setdefault(s, [0, 0.0])
                    d[0] += 1; d[1] += a.elapsed_time(z)
            json.dump({"config": cfg["key"], "per_gpu_batch": B, "ms_per_step": ms / args.steps, "eager_ms_per_step": ms_eager / esteps,
                       "gemm_ms_per_step": tot_ms / esteps, "gemm_share_of_step": (tot_ms / esteps) / (ms_eager / esteps),
                       "attn": attn_rec, "by_shape": table, "ffma_by_shape": {k: {"launches_per_step": v[0] / esteps, "ms_per_step": v[1] / esteps} for k, v in ffma.items()}},
                      open(args.profile_out, "w"), indent=1)
        # the family mixes tensor-bound (deep K) and HBM-bound (K <= 512, N = dim) launches: per-launch speed of light
        # max(flops / bf16 peak, algorithmic bytes / copy bandwidth), summed, against the measured time
        ideal_ms = sum(sol(r[2], r[3]) for r in tc) * 1e3
        traffic, traffic_src, step_dram = None, None, None
        try:      # DRAM bytes per launch of the same kernel family from the committed ncu capture of this command (c2 only)
            if cfg["key"] == "c2":
                tj = json.load(open(os.path.join(ROOT, "profiles", "r02_gemm_dram_bytes.json")))
                traffic = tj["dram_bytes_per_launch"]
                # whole-step DRAM traffic of the same capture next to SURVEY 8(d)'s ideal-fusion figure (8.9 GB activations +
                # 0.26 GB weights forward, x3 for training = 26.8 GB at B=256)
                step_dram = {"measured_bytes_per_step": tj["step_dram_bytes"], "read": tj["step_dram_read_bytes"],
                             "write": tj["step_dram_write_bytes"], "survey_ideal_bytes_per_step": 26.8e9,
                             "ratio": tj["step_dram_bytes"] / 26.8e9, "source": "profiles/r02_launches_dram.csv (ncu, 2 eager steps)"}
                traffic_src = "profiles/r02_gemm_dram_bytes.json (ncu dram__bytes_read+write, mean over the step's GEMM launches)"
        except Exception:
            pass
        roof = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                "frac_of_per_launch_speed_of_light": (ideal_ms / tot_ms) if tot_ms > 0 else None,
                "kernel": "dx_gemm_tc_kernel (tcgen05, all launches)",
                "launches_per_step": len(tc) / esteps, "flops_per_launch": tot_fl / max(len(tc), 1),
                "ms_per_launch": tot_ms / max(len(tc), 1), "share_of_step": (tot_ms / esteps) / (ms_eager / esteps), "peak_source": pk["source"],
                "timed_in": f"eager pass of the same step, {esteps} steps, CUDA events around every launch",
                "top_shapes": table[:6]}
        # ---- CPU baseline (reference's CPU path on this box's host cores, bounded sample) -------------------------------
        cpu, cpu_c1 = None, None
        if not args.no_cpu_baseline and world == 1:
            def ref_run(extra):
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "1"] + extra,
                                   capture_output=True, text=True, timeout=900)
                for line in r.stdout.splitlines():
                    if line.startswith("{"):
                        return json.loads(line)["cpu_baseline"]
                return None
            cpu = ref_run(["--config", cfg["key"]] + (["--cpu-sample", str(args.cpu_sample)] if args.cpu_sample else [])
                          + (["--batch", str(args.batch)] if args.batch else []))
            if cfg["key"] == "c2":      # the configuration BASELINE.json designates for the CPU: configs[0], full batch of 32
                cpu_c1 = ref_run(["--config", "c1"])
        fl = train_flops_per_sample(w, wl.task)
        out = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": args.precision,
            "data": "synthetic",
            "config": dict(config_dict(cfg, world, B),
                           l2="working set >> 126 MB L2 (every residual-stream tensor alone exceeds it), distinct input batches cycled"),
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": wl.h2d_bytes(dev_batches[0]), "d2h_bytes_per_step": 8,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "roofline": roof, "step_dram": step_dram, "attn": attn_rec, "cpu_baseline": cpu, "cpu_baseline_c1": cpu_c1, "clocks": clk,
            "parity_check": parity,
            "model_tflops": value * fl / 1e12,
            "model_frac_of_bf16_peak": value * fl / 1e12 / (world * pk["bf16_tflops_sustained"]),
            "fwd_bwd_allreduce_only": None if ms_noopt is None else {
                "value": world * B * args.steps / (ms_noopt / 1e3), "unit": "samples/s", "ms_per_step": ms_noopt / args.steps},
            "allreduce_buckets_per_step": wl.red.launched, "allreduce_trace": comm, "host_enqueue_ms_per_step": host_ms,
            "cuda_graph": gstep_was, "cuda_graph_error": graph_err, "eager_ms_per_step": ms_eager / esteps,
        }
        if args.batch:
            out["config"]["batch_override"] = args.batch
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        # The captured step graph holds NCCL kernels of this communicator; tearing the communicator down underneath it
        # (destroy_process_group / interpreter shutdown order) was seen to block forever on 2 GPUs.  Everything is
        # measured and printed: drain the device, meet the other ranks, and leave without the collective teardown.
        barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def _attn_ncu():
    """sm__pipe_tensor_cycles_active of the attention kernels from the committed ncu capture (profiles/), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_attn_tensor_pipe.json")))
    except Exception:
        return None


def parity_check(cfg, wl, gstep, st, device):
    """VERDICT r1 item 1c.  The measured thing (bf16 tcgen05 kernels, whole step replayed from the CUDA graph) against an
    independent arithmetic path on the SAME batch and weights: the same modules in fp32 mode (FFMA GEMMs, fp32 attention —
    the kernels the 1e-3 oracle parity tests run on), executed eagerly.  Hard bounds (2e-2 = the north-star bf16 bound):
      * encoder tokens of the backbone, bf16 vs fp32 (relative L2) — the part that is > 99.9 % of the FLOPs;
      * loss of the graph-replayed step vs the same bf16 step run eagerly (the replay computes what the eager code computes);
      * loss bf16 vs fp32.  For the supervised task this one is held to 1e-1 instead: its head normalises with a BatchNorm
        over the batch whose input (the [REP] token) differs between samples by ~0.5 % of its norm on a randomly initialised
        model, which amplifies the bf16 rounding of the tokens ~100x (tests/test_parity_bench_shapes_gpu.py measures this
        against the reference's own bf16 self-deviation and checks the head given identical tokens at the plain bound).
    The global gradient norms are reported.  Raises if a bound is broken, so a corrupted step cannot print a number."""
    try:
        torch.cuda.empty_cache()
        free, _ = torch.cuda.mem_get_info(device)
        need = wl.flat.numel * 4 * 4 + int(torch.cuda.max_memory_allocated(device) * 1.7)
        if free < need:
            return {"skipped": f"fp32 replica needs ~{need >> 30} GiB, {free >> 30} GiB free"}
        sd = {k: v.detach().clone() for k, v in wl.trainable.state_dict().items()}
        wl32 = Workload(cfg, device, precision="fp32", shadow=False, attach=False)
        wl32.trainable.load_state_dict(sd)
        if wl.teacher is not None:
            wl32.teacher.load_state_dict(wl.teacher.state_dict())
        wl32.n_steps = wl.n_steps
        rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
        enc = lambda W: (W.model if W.task != "kd" else W.model.duett).encode((st["xs_static"], st["xs_ts"], st["xs_times"], W.n_steps))
        with torch.no_grad():
            tok_err = rel(enc(wl).float(), enc(wl32).float())
        wl32.trainable.load_state_dict(sd)                       # the encode above advanced the BatchNorm running statistics
        wl.trainable.load_state_dict(sd)
        wl.flat.sync_shadow()
        loss16e = wl.train_step(st, optimizer=False)             # bf16, eager, no weight update
        torch.cuda.synchronize()
        l16e, g16e = float(loss16e.detach()), float(wl.flat.grad.double().norm())
        wl32.opt.zero_grad()
        # kd: the fp32 student is distilled from the SAME teacher logits as the bf16 one (the KD gradient is proportional to
        # sigmoid(z_s/T) - sigmoid(z_t/T), a difference of nearly equal numbers at initialisation, so the bf16 rounding of the
        # frozen teacher's logits would otherwise dominate the comparison of the student's gradients); the teacher's own
        # bf16-vs-fp32 deviation is reported separately
        zt16 = wl.last_z_t.detach().float() if wl.task == "kd" else None
        loss32 = wl32.loss_fn(st, z_t=zt16)
        loss32.backward()
        torch.cuda.synchronize()
        l32, g32 = float(loss32.detach()), float(wl32.flat.grad.double().norm())
        teacher_dev = None
        if wl.task == "kd":
            with torch.no_grad():
                wl32.loss_fn(st)
            teacher_dev = rel(zt16, wl32.last_z_t.float())
        # the graph replay on the same batch: its optimizer step runs after the loss and the gradients are complete
        wl.trainable.load_state_dict(sd)
        wl.flat.sync_shadow()
        l16g, g16g = l16e, g16e
        if gstep is not None:
            loss16g = gstep(**st)
            torch.cuda.synchronize()
            l16g, g16g = float(loss16g.detach()), float(wl.flat.grad.double().norm())
        r = lambda a, b: abs(a - b) / max(abs(b), 1e-30)
        loss_bound = 1e-1 if cfg["task"] == "supervised" else 2e-2
        res = {"tokens_rel_err_bf16_vs_fp32": tok_err, "loss_bf16_graph": l16g, "loss_bf16_eager": l16e, "loss_fp32_eager": l32,
               "loss_rel_err_graph_vs_eager": r(l16g, l16e), "loss_rel_err_bf16_vs_fp32": r(l16g, l32),
               "grad_norm_bf16_graph": g16g, "grad_norm_fp32_eager": g32, "grad_norm_rel_err": r(g16g, g32),
               "teacher_logits_rel_err_bf16_vs_fp32": teacher_dev,
               "bounds": {"tokens": 2e-2, "graph_vs_eager": 2e-2, "loss_bf16_vs_fp32": loss_bound}}
        res["ok"] = bool(tok_err < 2e-2 and res["loss_rel_err_graph_vs_eager"] < 2e-2 and res["loss_rel_err_bf16_vs_fp32"] < loss_bound)
        del wl32
        torch.cuda.empty_cache()
    except torch.cuda.OutOfMemoryError as ex:
        torch.cuda.empty_cache()
        return {"skipped": "out of memory for the fp32 replica: " + repr(ex)[:80]}
    if not res["ok"]:
        raise RuntimeError(f"parity check failed: {res}")
    return res


if __name__ == "__main__":
    main()
