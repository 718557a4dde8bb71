"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU (plain PyTorch, fp32) restatement of the reference's DuETT hot path.

Nothing in the product package may import this module; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do.  It restates, as pure functions over a dict of tensors keyed by the
reference's state-dict names, the algorithms of

  * duett/duett.py:24-39,84-125,151-157,239-323   (embedding MLPs, special tokens, time embedding, event/time
                                                   transformer stack, pooling, supervised + SSL heads)
  * duett/duett.py:189-237                        (pretrain_prep_batch: host numpy RNG masking)
  * duett/duett.py:337-365                        (SSL and supervised losses)
  * models/main_architecture_duett.py:31-94       (DuettFeatureExtractor.encode)
  * models/main_architecture_duett.py:536-654,745-774,1075-1129,1202-1235
                                                   (PatchDualPathologyPerceiver, _PerceiverBlock, TeacherModel
                                                   patch_dual branch on given CXR embeddings, StudentModel)
  * loss/losses_duett.py:8-25,39-57,135-194       (VanillaKLKD, StudentKDLoss, DualPathologyLoss)
  * training_duett/engine.py:149-165              (aux residual KL)
  * third party: x_transformers.Encoder (PyPI x-transformers, NOT pinned by the reference, not installed): published
    algorithm restated per SURVEY.md Appendix A — see oracle/shims/x_transformers/__init__.py.

Pinning: the reference has no tests / golden vectors for this path.  This restatement is pinned
against the reference's OWN files executed in place (oracle/make_golden.py imports /root/reference through
oracle/shims and writes tests/golden/*.npz; tests/test_oracle_golden.py replays them).  The x_transformers part is
PARITY UNPINNED (restated library, no reference vector exists).

The formulation is deliberately different from the reference's (batched einsum over variables instead of a Python loop,
functional instead of nn.Module) — it is a restatement of the algorithm, not a copy of the code.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


@dataclass
class DuettConfig:
    d_static_num: int
    d_time_series_num: int          # V
    n_timesteps: int                # T  (= masked_transform_timesteps = max_len)
    d_target: int = 1
    d_embedding: int = 24
    d_feedforward: int = 512
    n_heads: int = 2
    n_layers: int = 2
    d_hidden_mlp_embedding: int = 64
    d_hidden_tab_encoder: int = 128
    d_hidden_head: int = 64
    pretrain_presence_weight: float = 0.2
    final_norm: bool = True
    scalenorm_eps_mode: str = "normalize"

    @property
    def V(self): return self.d_time_series_num
    @property
    def T(self): return self.n_timesteps
    @property
    def et_dim(self): return self.d_embedding * (self.n_timesteps + 1)
    @property
    def tt_dim(self): return self.d_embedding * (self.d_time_series_num + 1)
    @property
    def d_time_hidden(self): return int(np.sqrt(self.tt_dim))   # duett/duett.py:154
    def ff_inner(self, dim):                                       # x_transformers: int(dim * (d_ff / dim))
        return int(dim * (self.d_feedforward / dim))


# ------------------------------------------------------------------------------------------------------------------
# parameter construction (shapes per SURVEY.md Appendix B); values are arbitrary — parity tests feed the SAME dict
# to the CUDA path and to this oracle.
# ------------------------------------------------------------------------------------------------------------------
def init_params(cfg: DuettConfig, seed: int = 0, dtype=torch.float32) -> dict:
    g = torch.Generator().manual_seed(seed)
    P = {}

    def lin(name, out_f, in_f, bias=True):
        bound = 1.0 / math.sqrt(in_f)
        P[name + ".weight"] = (torch.rand(out_f, in_f, generator=g, dtype=dtype) * 2 - 1) * bound
        if bias:
            P[name + ".bias"] = (torch.rand(out_f, generator=g, dtype=dtype) * 2 - 1) * bound

    def bnp(name, c):
        P[name + ".batch_norm.weight"] = 1.0 + 0.1 * torch.randn(c, generator=g, dtype=dtype)
        P[name + ".batch_norm.bias"] = 0.1 * torch.randn(c, generator=g, dtype=dtype)
        P[name + ".batch_norm.running_mean"] = torch.zeros(c, dtype=dtype)
        P[name + ".batch_norm.running_var"] = torch.ones(c, dtype=dtype)
        P[name + ".batch_norm.num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    d, V, T, S = cfg.d_embedding, cfg.V, cfg.T, cfg.d_static_num
    H = cfg.d_hidden_mlp_embedding
    P["MASKED_EMBEDDING_KEY"] = torch.tensor(0)
    P["REPRESENTATION_EMBEDDING_KEY"] = torch.tensor(1)
    P["special_embeddings.weight"] = torch.randn(8, d, generator=g, dtype=dtype)
    for i in range(V):
        lin(f"embedding_layers.{i}.0", H, 2)
        bnp(f"embedding_layers.{i}.3", H)
        lin(f"embedding_layers.{i}.4", d, H)
    P["n_obs_embedding.weight"] = torch.randn(16, 1, generator=g, dtype=dtype)
    for kind, dim in (("event", cfg.et_dim), ("time", cfg.tt_dim)):
        Fi = cfg.ff_inner(dim)
        for l in range(cfg.n_layers):
            p = f"{kind}_transformers.{l}"
            P[f"{p}.layers.0.0.0.g"] = 1.0 + 0.1 * torch.randn(1, generator=g, dtype=dtype)
            for nm in ("to_q", "to_k", "to_v"):
                lin(f"{p}.layers.0.1.{nm}", d, dim, bias=False)
            lin(f"{p}.layers.0.1.to_out", dim, d, bias=False)
            P[f"{p}.layers.1.0.0.g"] = 1.0 + 0.1 * torch.randn(1, generator=g, dtype=dtype)
            lin(f"{p}.layers.1.1.ff.0.0", Fi, dim)
            lin(f"{p}.layers.1.1.ff.2", dim, Fi)
            if cfg.final_norm:
                P[f"{p}.final_norm.g"] = 1.0 + 0.1 * torch.randn(1, generator=g, dtype=dtype)
    P["full_event_embedding.weight"] = torch.randn(V + 1, cfg.et_dim, generator=g, dtype=dtype)
    ht = cfg.d_time_hidden
    lin("full_time_embedding.0", ht, 1)
    bnp("full_time_embedding.2", ht)
    lin("full_time_embedding.3", cfg.tt_dim, ht)
    P["full_rep_embedding.weight"] = torch.randn(cfg.tt_dim, 1, generator=g, dtype=dtype)
    lin("head.0", cfg.d_hidden_head, cfg.tt_dim)
    bnp("head.3", cfg.d_hidden_head)
    lin("head.4", cfg.d_target, cfg.d_hidden_head)
    lin("pretrain_value_proj.0", V, cfg.tt_dim)
    lin("pretrain_presence_proj.0", V, cfg.tt_dim)
    lin("predict_events_proj.0", T, cfg.et_dim)
    lin("predict_events_presence_proj.0", T, cfg.et_dim)
    lin("tab_encoder.0", cfg.d_hidden_tab_encoder, S)
    bnp("tab_encoder.3", cfg.d_hidden_tab_encoder)
    lin("tab_encoder.4", d, cfg.d_hidden_tab_encoder)
    return P


def trainable_keys(P: dict) -> list:
    return [k for k, v in P.items() if v.is_floating_point() and "running_" not in k]


# ------------------------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------------------------
def batch_norm_lastdim(x, P, prefix, training, stats_out=None):
    """BatchNormLastDim (duett/duett.py:11-22): nn.BatchNorm1d over the last dim, statistics over all leading dims."""
    w, b = P[prefix + ".batch_norm.weight"], P[prefix + ".batch_norm.bias"]
    flat = x.reshape(-1, x.shape[-1])
    if training:
        xf = flat.float()
        mean = xf.mean(0)
        var = xf.var(0, unbiased=False)
        if stats_out is not None:
            n = flat.shape[0]
            stats_out[prefix] = (mean.detach(), (var * n / max(n - 1, 1)).detach())
    else:
        mean, var = P[prefix + ".batch_norm.running_mean"].float(), P[prefix + ".batch_norm.running_var"].float()
    y = (flat.float() - mean) * torch.rsqrt(var + BN_EPS) * w.float() + b.float()
    return y.to(x.dtype).reshape(x.shape)


def scale_norm(x, g, mode="normalize"):
    """x_transformers ScaleNorm: x / ||x|| * sqrt(dim) * g (computed in fp32 like F.normalize under autocast)."""
    dim = x.shape[-1]
    xf = x.float()
    n = xf.norm(dim=-1, keepdim=True)
    if mode == "normalize":
        return xf / n.clamp_min(1e-12) * (dim ** 0.5) * g.float()
    return xf / (n * dim ** -0.5).clamp_min(1e-5) * g.float()


def keep_factor(seed, n, p, step=0):
    """The kernels' dropout generator restated (csrc/dx_common.cuh dx_rng32): keep/(1-p) over flat indices 0..n-1, keep iff
    splitmix64(seed + idx * 0x9E3779B97F4A7C15) >> 32 >= p * 2^32.  torch's Philox masks are not reproducible outside
    torch, so dropout parity is checked with both sides on THIS generator (the seeds are read back from the product)."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64)
        z = (np.uint64((int(seed) + int(step)) & 0xFFFFFFFFFFFFFFFF) + idx * np.uint64(0x9E3779B97F4A7C15)) & M
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M
        z = z ^ (z >> np.uint64(31))
        r = z >> np.uint64(32)
    thresh = min(int(float(np.float32(p)) * 4294967296.0), 4294967295)
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return torch.from_numpy(np.where(r >= thresh, scale, np.float32(0)).astype(np.float32))


def _drop(x, drop, key):
    """x * mask/(1-p) when `drop` (dict site -> (p, seed)) names this site; identity otherwise."""
    if not drop or key not in drop:
        return x
    p, seed = drop[key]
    return x * keep_factor(seed, x.numel(), p).reshape(x.shape).to(x.dtype)


def encoder(P, prefix, x, cfg: DuettConfig, drop=None):
    """One x_transformers.Encoder(depth=1) exactly as constructed at duett/duett.py:95-105 (attn_dropout on the softmax
    output, ff_dropout after the GELU: sites `<prefix>.attn`, `<prefix>.ff` of `drop`)."""
    B, N, dim = x.shape
    h, d = cfg.n_heads, cfg.d_embedding
    dh = d // h
    a = scale_norm(x, P[f"{prefix}.layers.0.0.0.g"], cfg.scalenorm_eps_mode)
    q = F.linear(a, P[f"{prefix}.layers.0.1.to_q.weight"])
    k = F.linear(a, P[f"{prefix}.layers.0.1.to_k.weight"])
    v = F.linear(a, P[f"{prefix}.layers.0.1.to_v.weight"])
    q, k, v = (t.reshape(B, N, h, dh).permute(0, 2, 1, 3) for t in (q, k, v))
    sim = torch.matmul(q, k.transpose(-1, -2)) * dh ** -0.5
    attn = torch.softmax(sim, dim=-1, dtype=torch.float32).to(sim.dtype)
    attn = _drop(attn, drop, prefix + ".attn")
    o = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(B, N, h * dh)
    x = x + F.linear(o, P[f"{prefix}.layers.0.1.to_out.weight"])
    f = scale_norm(x, P[f"{prefix}.layers.1.0.0.g"], cfg.scalenorm_eps_mode)
    hdn = F.gelu(F.linear(f, P[f"{prefix}.layers.1.1.ff.0.0.weight"], P[f"{prefix}.layers.1.1.ff.0.0.bias"]))
    hdn = _drop(hdn, drop, prefix + ".ff")
    x = x + F.linear(hdn, P[f"{prefix}.layers.1.1.ff.2.weight"], P[f"{prefix}.layers.1.1.ff.2.bias"])
    if cfg.final_norm:
        x = scale_norm(x, P[f"{prefix}.final_norm.g"], cfg.scalenorm_eps_mode)
    return x


def embed_psi(P, cfg: DuettConfig, xs_static, xs_feats, training, stats_out=None):
    """psi[B,T+1,V+1,d] (duett/duett.py:245-266 == models/main_architecture_duett.py:31-65), all variables at once."""
    B, T, _ = xs_feats.shape
    V, d = cfg.V, cfg.d_embedding
    values, counts, step_mask = xs_feats[:, :, :V], xs_feats[:, :, V:2 * V], xs_feats[:, :, -1]
    event_masked = counts == -1                                           # [B,T,V]
    idx = counts.to(torch.long).clamp(0, 15)
    cnt_emb = P["n_obs_embedding.weight"][:, 0][idx]                      # [B,T,V]
    inp = torch.stack((values, cnt_emb.to(values.dtype)), dim=-1)         # [B,T,V,2]
    W0 = torch.stack([P[f"embedding_layers.{i}.0.weight"] for i in range(V)])   # [V,H,2]
    b0 = torch.stack([P[f"embedding_layers.{i}.0.bias"] for i in range(V)])     # [V,H]
    W4 = torch.stack([P[f"embedding_layers.{i}.4.weight"] for i in range(V)])   # [V,d,H]
    b4 = torch.stack([P[f"embedding_layers.{i}.4.bias"] for i in range(V)])     # [V,d]
    hid = torch.relu(torch.einsum("btvi,vhi->btvh", inp, W0.to(inp.dtype)) + b0)        # [B,T,V,H]
    cols = []
    for i in range(V):   # per-variable BatchNorm statistics over B*T rows
        cols.append(batch_norm_lastdim(hid[:, :, i, :], P, f"embedding_layers.{i}.3", training, stats_out))
    hid = torch.stack(cols, dim=2)
    emb = torch.einsum("btvh,vdh->btvd", hid, W4.to(hid.dtype)) + b4      # [B,T,V,d]
    psi = torch.zeros(B, T + 1, V + 1, d, dtype=xs_feats.dtype, device=xs_feats.device)
    psi[:, :T, :V] = emb.to(psi.dtype)
    tab = torch.relu(F.linear(xs_static, P["tab_encoder.0.weight"], P["tab_encoder.0.bias"]))
    tab = batch_norm_lastdim(tab, P, "tab_encoder.3", training, stats_out)
    tab = F.linear(tab, P["tab_encoder.4.weight"], P["tab_encoder.4.bias"])           # [B,d]
    psi[:, :T, V] = tab.to(psi.dtype)[:, None, :]
    sp = P["special_embeddings.weight"]
    psi[:, T] = sp[1].to(psi.dtype)                                       # [REP] row (written after the static column)
    m_row = torch.cat((step_mask == 1, torch.zeros(B, 1, dtype=torch.bool, device=psi.device)), dim=1)   # [B,T+1]
    m_cell = torch.zeros(B, T + 1, V + 1, dtype=torch.bool, device=psi.device)
    m_cell[:, :T, :V] = event_masked
    m_cell[:, T, :V] = event_masked[:, 0]                                 # [REP] row copies row 0's mask
    masked = m_row[:, :, None] | m_cell
    psi = torch.where(masked[..., None], sp[0].to(psi.dtype), psi)
    return psi, event_masked


def time_embeddings(P, cfg: DuettConfig, xs_times, training, stats_out=None):
    """cve(...) on bin-end times + [REP] row (duett/duett.py:151-157,269-272)."""
    B = xs_times.shape[0]
    t = torch.tanh(F.linear(xs_times[..., None], P["full_time_embedding.0.weight"], P["full_time_embedding.0.bias"]))
    t = batch_norm_lastdim(t, P, "full_time_embedding.2", training, stats_out)
    t = F.linear(t, P["full_time_embedding.3.weight"], P["full_time_embedding.3.bias"])   # [B,T,E']
    rep = P["full_rep_embedding.weight"][:, 0].to(t.dtype)[None, None, :].expand(B, 1, -1)
    return torch.cat((t, rep), dim=1)


def encode(P, cfg: DuettConfig, xs_static, xs_feats, xs_times, training=True, stats_out=None, return_psi=False, drop=None):
    """DuettFeatureExtractor.encode -> transformed [B,T+1,E'] (models/main_architecture_duett.py:31-94)."""
    psi, event_masked = embed_psi(P, cfg, xs_static, xs_feats, training, stats_out)
    te = time_embeddings(P, cfg, xs_times, training, stats_out)
    B, T1, V1, d = psi.shape
    pos_e = P["full_event_embedding.weight"]
    for l in range(cfg.n_layers):
        ev = psi.permute(0, 2, 1, 3).reshape(B, V1, T1 * d) + pos_e           # event view: token = variable
        ev = encoder(P, f"event_transformers.{l}", ev, cfg, drop)
        tv = ev.reshape(B, V1, T1, d).permute(0, 2, 1, 3).reshape(B, T1, V1 * d) + te   # time view: token = time bin
        psi = encoder(P, f"time_transformers.{l}", tv, cfg, drop).reshape(B, T1, V1, d)
    out = psi.reshape(B, T1, V1 * d)
    if return_psi:
        return out, psi, event_masked
    return out


def simple_head(P, prefix, z, training, stats_out=None):
    """simple_mlp(d_in, d_out, 1, d_hidden, hidden_batch_norm=True): Linear-ReLU-Dropout(0)-BN-Linear (duett.py:24-39)."""
    h = torch.relu(F.linear(z, P[prefix + ".0.weight"], P[prefix + ".0.bias"]))
    h = batch_norm_lastdim(h, P, prefix + ".3", training, stats_out)
    return F.linear(h, P[prefix + ".4.weight"], P[prefix + ".4.bias"])


def model_forward_supervised(P, cfg, xs_static, xs_feats, xs_times, fusion_method="rep_token", training=True,
                             stats_out=None):
    """Model.forward(pretrain=False) (duett/duett.py:280-323) -> logits [B] (d_target == 1) or [B,d_target]."""
    tr = encode(P, cfg, xs_static, xs_feats, xs_times, training, stats_out)
    if fusion_method == "rep_token":
        z = tr[:, -1]
    elif fusion_method == "averaging":
        z = tr[:, :-1].mean(1)
    elif fusion_method == "masked_embed":
        step = (xs_feats[:, :, -1] == 1).float().argmax(1)
        z = tr[torch.arange(tr.shape[0]), step]
    else:
        raise ValueError(fusion_method)
    return simple_head(P, "head", z, training, stats_out).squeeze(1)


def model_forward_pretrain(P, cfg, xs_static, xs_feats, xs_times, training=True, stats_out=None, masked_steps=1):
    """Model.forward(pretrain=True) (duett/duett.py:284-316).  masked_steps > 1 (:287-293): the rows of the DISTINCT masked
    timesteps in time order, zero-padded to masked_steps rows -> every head output gains a [masked_steps] axis."""
    tr, psi, event_masked = encode(P, cfg, xs_static, xs_feats, xs_times, training, stats_out, return_psi=True)
    B = tr.shape[0]
    ar = torch.arange(B, device=tr.device)
    if masked_steps > 1:
        flag = F.pad(xs_feats[:, :, -1] > 0, (0, 1), value=False)         # [B,T+1]
        z = torch.stack([F.pad(tr[b, flag[b]], (0, 0, 0, masked_steps - int(flag[b].sum()))) for b in range(B)])   # [B,k,E']
    else:
        step = (xs_feats[:, :, -1] == 1).float().argmax(1)                # the single masked timestep per sample
        z = tr[ar, step]                                                  # [B,E']
    y_val = F.linear(z, P["pretrain_value_proj.0.weight"], P["pretrain_value_proj.0.bias"])
    y_pres = F.linear(z, P["pretrain_presence_proj.0.weight"], P["pretrain_presence_proj.0.bias"])
    var = event_masked[:, 0].float().argmax(1)                            # the single masked variable per sample
    z_ev = psi[ar, :, var].reshape(B, -1)                                 # [B,(T+1)*d]: that variable's column
    y_ev = F.linear(z_ev, P["predict_events_proj.0.weight"], P["predict_events_proj.0.bias"])
    y_ev_pres = F.linear(z_ev, P["predict_events_presence_proj.0.weight"], P["predict_events_presence_proj.0.bias"])
    return y_val, y_pres, y_ev, y_ev_pres


# ------------------------------------------------------------------------------------------------------------------
# host-side batch preparation (kept host-side and byte-identical: same numpy RNG calls in the same order)
# ------------------------------------------------------------------------------------------------------------------
def feats_to_input(x_ts, x_static, times, max_len):
    """Model.feats_to_input without augmentation (duett/duett.py:159-187): append mask column, pad, stack."""
    x_ts, times = list(x_ts), list(times)
    for i, f in enumerate(x_ts):
        if f.shape[0] > max_len:
            f, times[i] = f[-max_len:], times[i][-max_len:]
        x_ts[i] = torch.cat((f, torch.zeros_like(f[:, :1])), dim=1)
    n_timesteps = [len(t) for t in times]
    pad_to = max(n_timesteps)
    xs_ts = torch.stack([F.pad(t, (0, 0, 0, pad_to - t.shape[0])) for t in x_ts])
    xs_times = torch.stack([F.pad(t, (0, pad_to - t.shape[0])) for t in times])
    return torch.stack(list(x_static)), xs_ts, xs_times, n_timesteps


def pretrain_prep_batch(rng: np.random.Generator, cfg: DuettConfig, xs_ts, n_timesteps, pretrain_dropout=0.5, masked_steps=1):
    """Model.pretrain_prep_batch, predict_events=True (duett/duett.py:189-237).  masked_steps > 1 (:199-203): that many
    timesteps per sample drawn WITH replacement (targets keep the draw order and the duplicates); every sample needs at
    least max(2, masked_steps) timesteps (shorter ones make the reference's torch.stack fail on ragged targets)."""
    B, T, _ = xs_ts.shape
    V = cfg.V
    steps, evs = [], []
    for n in n_timesteps:                       # RNG call order: (timestep(s), variable) per sample
        if masked_steps > 1:
            if n < max(2, masked_steps):
                raise ValueError("pretrain_masked_steps > 1 needs at least that many timesteps in every sample")
            steps.append([int(i) for i in rng.choice(np.arange(n), size=masked_steps)])
        else:
            steps.append(n if n < 2 else int(rng.choice(np.arange(0, n))))
        evs.append(int(rng.choice(np.arange(0, V))))
    steps_t, evs_t = torch.tensor(steps), torch.tensor(evs)
    ar = torch.arange(B)
    if masked_steps > 1:
        y_ts = xs_ts[ar[:, None], steps_t, :V].clone()                    # [B,k,V]
        y_mask = xs_ts[ar[:, None], steps_t, V:2 * V].clip(0, 1)
    else:
        y_ts = xs_ts[ar, steps_t, :V].clone()
        y_mask = xs_ts[ar, steps_t, V:2 * V].clip(0, 1)
    y_events = xs_ts[ar, :, evs_t].clone()
    y_events_mask = xs_ts[ar, :, evs_t + V].clip(0, 1)
    x = xs_ts.clone()
    for b in range(B):                          # order matters where the masked step and masked variable intersect
        x[b, steps[b], :] = 0.0
        x[b, steps[b], -1] = 1.0
        x[b, :, evs[b]] = 0.0
        x[b, :, evs[b] + V] = -1.0
    if pretrain_dropout > 0:
        keep = torch.tensor(rng.random((B, V)) > pretrain_dropout)
        keep = torch.logical_or(1 - (y_mask.sum(dim=1).clip(0, 1) if masked_steps > 1 else y_mask), keep)
        keep = torch.cat((keep.tile(1, 2), torch.ones(B, 1)), dim=1)
        x = x * torch.logical_or(keep.unsqueeze(1), x == -1)
    return x, y_ts, y_mask, y_events, y_events_mask


# ------------------------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------------------------
def ssl_loss(y_val, y_pres, y_ev, y_ev_pres, y, mask, y_events, y_events_mask, presence_weight=0.2):
    """duett/duett.py:337-358.  With pretrain_masked_steps > 1 the targets are [B,k,V] and the reference averages the k
    per-step losses (:338-349) — every step has B*V terms, so that is the mean over all [B,k,V] terms."""
    loss = F.mse_loss(y_val * mask, y * mask)
    loss = loss + F.binary_cross_entropy_with_logits(y_pres, mask) * presence_weight
    loss = loss + F.mse_loss(y_ev * y_events_mask, y_events * y_events_mask)
    loss = loss + F.binary_cross_entropy_with_logits(y_ev_pres, y_events_mask) * presence_weight
    return loss


def supervised_loss(logits, y, pos_frac=None):
    """duett/duett.py:360-365: BCE-with-logits, optional class-balance weights; y is float64 in the reference."""
    y = y.double()
    if pos_frac is None:
        return F.binary_cross_entropy_with_logits(logits.double(), y)
    w = torch.where(y > 0, 1.0 / (2 * pos_frac), 1.0 / (2 * (1 - pos_frac)))
    return F.binary_cross_entropy_with_logits(logits.double(), y, w)


def vanilla_kl_kd(z_s, z_t, T=4.0, eps=1e-7):
    """loss/losses_duett.py:8-25."""
    p_t = torch.sigmoid(z_t.detach() / T).clamp(eps, 1 - eps)
    p_s = torch.sigmoid(z_s / T).clamp(eps, 1 - eps)
    kl = p_t * (p_t.log() - p_s.log()) + (1 - p_t) * ((1 - p_t).log() - (1 - p_s).log())
    return (T ** 2) * kl.mean()


def student_kd_loss(z_s, z_t, y, kd_T=4.0, kd_alpha=0.5, pos_weight=None):
    """loss/losses_duett.py:39-57 -> dict(total, bce, kd)."""
    kd = vanilla_kl_kd(z_s, z_t, kd_T)
    pw = None if pos_weight is None else torch.tensor([pos_weight], dtype=torch.float32, device=z_s.device)
    bce = F.binary_cross_entropy_with_logits(z_s, y.float(), pos_weight=pw)
    return {"total": kd_alpha * bce + (1 - kd_alpha) * kd, "bce": bce.detach(), "kd": kd.detach()}


def masked_multilabel_bce(logits, y, mask, pos_weight=None, eps=1e-6):
    """per-pathology sum(bce*m)/(sum(m)+eps) -> [K]  (loss/losses_duett.py:152-165)."""
    pw = None if pos_weight is None else pos_weight.float()
    l = F.binary_cross_entropy_with_logits(logits, y, reduction="none", pos_weight=pw)
    return (l * mask).sum(0) / (mask.sum(0) + eps)


def dual_pathology_loss(img, ts, fus, y, mask, label_weights, pos_weight=None, alpha_img=0.5, alpha_ts=0.5,
                        alpha_fus=1.0, eps=1e-6):
    """loss/losses_duett.py:135-194."""
    per = [masked_multilabel_bce(l, y, mask, pos_weight, eps) for l in (img, ts, fus)]
    tot = [(label_weights.float() * p).sum() for p in per]
    total = alpha_img * tot[0] + alpha_ts * tot[1] + alpha_fus * tot[2]
    return {"total": total, "img_total": tot[0].detach(), "ts_total": tot[1].detach(), "fus_total": tot[2].detach(),
            "img_per": per[0].detach(), "ts_per": per[1].detach(), "fus_per": per[2].detach()}


def aux_residual_kl(img_logits, scaled_correction, y_multi, mask, eps=0.05):
    """training_duett/engine.py:149-165."""
    y = y_multi.float()
    ys = y * (1 - eps) + (1 - y) * eps
    p = torch.sigmoid(img_logits.detach() + scaled_correction).clamp(1e-6, 1 - 1e-6)
    kl = ys * (ys.log() - p.log()) + (1 - ys) * ((1 - ys).log() - (1 - p).log())
    m = mask.float()
    return (kl * m).sum() / m.sum().clamp(min=1.0)


# ------------------------------------------------------------------------------------------------------------------
# student / teacher heads
# ------------------------------------------------------------------------------------------------------------------
def init_student_head(cfg: DuettConfig, seed=1, head_hidden=128):
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, o, i in (("head.0", head_hidden, cfg.tt_dim), ("head.3", 1, head_hidden)):
        b = 1 / math.sqrt(i)
        P[name + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * b
        P[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * b
    return P


def student_forward(P, H, cfg, xs_static, xs_feats, xs_times, pool="mean", training=True, stats_out=None, drop=None):
    """StudentModel.forward (models/main_architecture_duett.py:1221-1235); `drop`: dropout sites (encoders, "head")."""
    tok = encode(P, cfg, xs_static, xs_feats, xs_times, training, stats_out, drop=drop)
    feat = tok[:, -1] if pool == "rep_token" else tok[:, :-1].mean(1)
    h = F.gelu(F.linear(feat, H["head.0.weight"], H["head.0.bias"]))
    h = _drop(h, drop, "head")
    return F.linear(h, H["head.3.weight"], H["head.3.bias"]).squeeze(-1)


def layer_norm(x, w, b):
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), 1e-5).to(x.dtype)


def mha(Pp, prefix, q_in, kv_in, n_heads):
    """nn.MultiheadAttention(batch_first=True), key == value, no masks, dropout off."""
    E = q_in.shape[-1]
    W, bias = Pp[prefix + ".in_proj_weight"], Pp[prefix + ".in_proj_bias"]
    q = F.linear(q_in, W[:E], bias[:E])
    k = F.linear(kv_in, W[E:2 * E], bias[E:2 * E])
    v = F.linear(kv_in, W[2 * E:], bias[2 * E:])
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    dh = E // n_heads
    q = q.reshape(B, Lq, n_heads, dh).transpose(1, 2)
    k = k.reshape(B, Lk, n_heads, dh).transpose(1, 2)
    v = v.reshape(B, Lk, n_heads, dh).transpose(1, 2)
    att = torch.softmax(torch.matmul(q, k.transpose(-1, -2)).float() / math.sqrt(dh), dim=-1).to(v.dtype)
    o = torch.matmul(att, v).transpose(1, 2).reshape(B, Lq, E)
    return F.linear(o, Pp[prefix + ".out_proj.weight"], Pp[prefix + ".out_proj.bias"]), att.mean(1)


def perceiver_block(Pp, prefix, latents, kv, n_heads):
    """_PerceiverBlock.forward (models/main_architecture_duett.py:745-774), dropout off."""
    q = layer_norm(latents, Pp[prefix + ".norm_q.weight"], Pp[prefix + ".norm_q.bias"])
    k = layer_norm(kv, Pp[prefix + ".norm_kv.weight"], Pp[prefix + ".norm_kv.bias"])
    a, w = mha(Pp, prefix + ".attn", q, k, n_heads)
    latents = latents + a
    f = layer_norm(latents, Pp[prefix + ".norm_ff.weight"], Pp[prefix + ".norm_ff.bias"])
    f = F.linear(F.gelu(F.linear(f, Pp[prefix + ".ff.0.weight"], Pp[prefix + ".ff.0.bias"])),
                 Pp[prefix + ".ff.3.weight"], Pp[prefix + ".ff.3.bias"])
    return latents + f, w


def init_teacher_head(cfg: DuettConfig, seed=2, K=7, d_latent=256, d_img=768, head_hidden=64):
    """Parameters of TeacherModel.img_proj + PatchDualPathologyPerceiver, keyed as in TeacherModel.state_dict()."""
    g = torch.Generator().manual_seed(seed)
    P = {}

    def lin(name, o, i, bias=True, zero=False):
        b = 1 / math.sqrt(i)
        P[name + ".weight"] = torch.zeros(o, i) if zero else (torch.rand(o, i, generator=g) * 2 - 1) * b
        if bias:
            P[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * b

    def ln(name, c):
        P[name + ".weight"] = 1 + 0.1 * torch.randn(c, generator=g)
        P[name + ".bias"] = 0.1 * torch.randn(c, generator=g)

    lin("img_proj", d_latent, d_img)
    pp = "perceiver."
    P[pp + "shared_queries"] = torch.randn(K, d_latent, generator=g) * 0.02
    lin(pp + "ts_proj", d_latent, cfg.tt_dim)
    for blk in ("img_cross", "img_self", "ts_cross", "ts_self"):
        for n in ("norm_q", "norm_kv", "norm_ff"):
            ln(pp + f"{blk}.{n}", d_latent)
        b = 1 / math.sqrt(d_latent)
        P[pp + f"{blk}.attn.in_proj_weight"] = (torch.rand(3 * d_latent, d_latent, generator=g) * 2 - 1) * b
        P[pp + f"{blk}.attn.in_proj_bias"] = 0.02 * torch.randn(3 * d_latent, generator=g)
        lin(pp + f"{blk}.attn.out_proj", d_latent, d_latent)
        lin(pp + f"{blk}.ff.0", 4 * d_latent, d_latent)
        lin(pp + f"{blk}.ff.3", d_latent, 4 * d_latent)
    for hd in ("image_head", "temporal_head"):
        lin(pp + f"{hd}.0", head_hidden, d_latent)
        lin(pp + f"{hd}.3", 1, head_hidden)
    ln(pp + "correction_head.0", d_latent)
    lin(pp + "correction_head.1", head_hidden, d_latent)
    P[pp + "correction_head.4.weight"] = 0.05 * torch.randn(1, head_hidden, generator=g)   # ref inits zeros; non-zero exercises the path
    P[pp + "beta"] = 1 + 0.1 * torch.randn(K, generator=g)
    P[pp + "image_label_bias"] = 0.1 * torch.randn(K, generator=g)
    P[pp + "temporal_label_bias"] = 0.1 * torch.randn(K, generator=g)
    return P


def perceiver_forward(Pt, ts_tokens, img_patches_proj, n_heads=4, ts_ablation="hourly_only"):
    """PatchDualPathologyPerceiver.forward (models/main_architecture_duett.py:595-654), dropout off."""
    pp = "perceiver."
    B = ts_tokens.shape[0]
    q0 = Pt[pp + "shared_queries"][None].expand(B, -1, -1)
    if ts_ablation == "hourly_only":
        ts_sel = ts_tokens[:, :-1]
    elif ts_ablation == "full":
        ts_sel = ts_tokens
    elif ts_ablation == "rep_only":
        ts_sel = ts_tokens[:, -1:]
    else:
        raise ValueError(ts_ablation)
    ts_kv = F.linear(ts_sel, Pt[pp + "ts_proj.weight"], Pt[pp + "ts_proj.bias"])
    I, img_attn = perceiver_block(Pt, pp + "img_cross", q0, img_patches_proj, n_heads)
    I, _ = perceiver_block(Pt, pp + "img_self", I, I, n_heads)
    Tt, ts_attn = perceiver_block(Pt, pp + "ts_cross", q0, ts_kv, n_heads)
    Tt, _ = perceiver_block(Pt, pp + "ts_self", Tt, Tt, n_heads)

    def head(prefix, x):
        return F.linear(F.gelu(F.linear(x, Pt[prefix + ".0.weight"], Pt[prefix + ".0.bias"])),
                        Pt[prefix + ".3.weight"], Pt[prefix + ".3.bias"]).squeeze(-1)

    img_logits = head(pp + "image_head", I) + Pt[pp + "image_label_bias"][None]
    ts_logits = head(pp + "temporal_head", Tt) + Pt[pp + "temporal_label_bias"][None]
    c = layer_norm(Tt, Pt[pp + "correction_head.0.weight"], Pt[pp + "correction_head.0.bias"])
    c = F.gelu(F.linear(c, Pt[pp + "correction_head.1.weight"], Pt[pp + "correction_head.1.bias"]))
    corr = F.linear(c, Pt[pp + "correction_head.4.weight"]).squeeze(-1)
    scaled = Pt[pp + "beta"][None] * corr
    fusion = img_logits.detach() + scaled
    return {"img_logits": img_logits, "ts_logits": ts_logits, "fusion_logits": fusion, "img_tokens": I,
            "ts_tokens": Tt, "ts_correction": corr, "scaled_correction": scaled, "img_attn": img_attn,
            "ts_attn": ts_attn}


def teacher_forward(P, Pt, cfg, xs_static, xs_feats, xs_times, img_patches, training=True, stats_out=None):
    """TeacherModel.forward, patch_dual branch, with the CXR encoder replaced by given patch embeddings
    [B,N,768] (models/main_architecture_duett.py:1093-1129)."""
    tok = encode(P, cfg, xs_static, xs_feats, xs_times, training, stats_out)
    proj = F.linear(img_patches, Pt["img_proj.weight"], Pt["img_proj.bias"])
    out = perceiver_forward(Pt, tok, proj)
    out["main_logit"] = out["fusion_logits"][:, 0]
    return out


# ------------------------------------------------------------------------------------------------------------------
# synthetic MIMIC-shaped data
# ------------------------------------------------------------------------------------------------------------------
def synth_batch(cfg: DuettConfig, B: int, seed: int, density=0.2):
    g = torch.Generator().manual_seed(seed)
    T, V, S = cfg.T, cfg.V, cfg.d_static_num
    obs = (torch.rand(B, T, V, generator=g) < density).float()
    vals = torch.randn(B, T, V, generator=g) * obs
    cnts = obs * torch.randint(1, 4, (B, T, V), generator=g).float()
    x_ts = torch.cat((vals, cnts), dim=2)
    x_static = torch.cat((torch.randn(B, 1, generator=g), (torch.rand(B, S - 1, generator=g) < 0.2).float()), dim=1)
    bin_ends = (torch.arange(1, T + 1).float() / 24.0)[None].expand(B, -1).contiguous()
    y = (torch.rand(B, generator=g) < 0.3).float()
    return {"x_ts": tuple(x_ts), "x_static": tuple(x_static), "bin_ends": tuple(bin_ends), "y": y}
