"""Fused DuETT backbone: DuettFeatureExtractor.encode (models/main_architecture_duett.py:31-94 == the body of
Model.forward, duett/duett.py:245-280) as ONE autograd.Function with hand-managed activations.

Data layout in HBM (act dtype = bf16 in the bf16 mode, f32 in the fp32 mode):
  psi / residual streams  [B, T+1, V+1, d] (time-major: token (b,t), features (v,dd)) or
                          [B, V+1, T+1, d] (event-major: token (b,v), features (t,dd)); each Encoder works on a dense
                          [N_tokens, dim] matrix so every contraction is a plain K-major / MN-major GEMM.
  per-token statistics    f32 vectors [N_tokens] (squared norms for ScaleNorm, <x,dy> row dots for its backward).
Per x_transformers Encoder (duett/duett.py:95-105) the kernel sequence is
  relayout(+final ScaleNorm of the previous encoder, + positional add, + row norms)          dx_relayout_fwd
  qkv = (x @ Wqkv^T) * s_a[row]                                                               dx_gemm (tcgen05)
  o   = softmax(q k^T / sqrt(dh)) v                                                           dx_attn_fwd
  x1  = x + o @ Wo^T                        (+ row norms of x1)                               dx_gemm
  h   = gelu((x1 @ W1^T) * s_f[row] + b1)                                                     dx_gemm
  x2  = x1 + h @ W2^T + b2                  (+ row norms of x2 for the final ScaleNorm)       dx_gemm
ScaleNorm is a per-row scalar, so it commutes with the contractions and is applied in their epilogues; its backward
projection term uses <x, dA> = <x W^T, dY> computed on the small [N,3d] / [N,F] tensors, so no extra pass over psi.
"""
from __future__ import annotations

import math

import torch

from . import ops
from .functional import BatchNorm2dFn, grad_sink, linear

ENC_KEYS = ("g_attn", "wqkv", "wo", "g_ff", "w1", "b1", "w2", "b2", "g_final")

# Registered by ddp.GradReducer.attach(): every hook is called as hook("time_transformers.3", ("w2", "b2", "g_final")) from
# inside the backward as soon as those parameter gradients of that encoder are final (all kernels writing them are
# enqueued), so their all-reduce bucket starts while the rest of the backward is still running.
GRAD_READY_HOOKS = []


def _notify(prefix, keys):
    for h in GRAD_READY_HOOKS:
        h(prefix, keys)


class EncoderCtx:
    """Saved activations of one Encoder forward."""
    __slots__ = ("x", "rowsq_x", "s_a", "qkv", "o", "lse", "x1", "rowsq1", "s_f", "fpre", "h", "x2", "rowsq2", "B", "S",
                 "dim", "drop_attn", "drop_ff")


class ZeroPool:
    """Zero-initialised f32 scratch handed out in slices: ONE memset for all the per-row accumulators (row norms, row dots,
    split-K partial sums) of a whole forward or backward pass instead of a fill kernel per buffer."""

    def __init__(self, numel, device):
        self.buf = torch.zeros(max(int(numel), 1), device=device, dtype=torch.float32)
        self.pos = 0

    def take(self, n):
        n8 = (int(n) + 7) // 8 * 8                      # slices stay 32 B aligned
        if self.pos + n8 > self.buf.numel():            # never hand out unzeroed memory
            return torch.zeros(int(n), device=self.buf.device, dtype=torch.float32)
        out = self.buf[self.pos:self.pos + int(n)]
        self.pos += n8
        return out


def _zeros(pool, n, dev):
    return pool.take(n) if pool is not None else torch.zeros(n, device=dev, dtype=torch.float32)


def encoder_fwd(x, rowsq_x, B, S, dim, p, heads, d, keep=True, drop_p=0.0, tag="", pool=None):
    """x: [B,S,dim] act tensor (dense), rowsq_x [B*S]. p: dict ENC_KEYS -> tensors (weights already in act dtype under
    'wqkv_c','wo_c','w1_c','w2_c'). Returns ctx with x2 / rowsq2 (un-normalised output + row norms)."""
    N = B * S
    dev, at = x.device, x.dtype
    F = p["w1"].shape[0]
    x2d = x.view(N, dim)
    s_a = ops.scalenorm_scale(rowsq_x, p["g_attn"], dim)
    qkv = torch.empty((N, 3 * d), device=dev, dtype=at)
    ops.gemm_(x2d, p["wqkv_c"], out=qkv, row_scale=s_a, act_dtype=at)
    q3 = qkv.view(B, S, 3 * d)
    # attn_dropout / ff_dropout of the x_transformers Encoder (duett/duett.py:98-99,104-105): on the attention probabilities
    # and on the FFN hidden after the GELU
    drop_attn = (drop_p, ops.drop_seed(tag + ".attn", drop_p, (B, heads, S, S))) if drop_p > 0 else None
    drop_ff = (drop_p, ops.drop_seed(tag + ".ff", drop_p, (N, F))) if drop_p > 0 else None
    o, lse = ops.attn_fwd(q3[:, :, :d], q3[:, :, d:2 * d], q3[:, :, 2 * d:], heads, drop_attn)
    x1 = torch.empty((N, dim), device=dev, dtype=at)
    rowsq1 = _zeros(pool, N, dev)
    ops.gemm_(o.view(N, d), p["wo_c"], out=x1, res=x2d, row_sumsq=rowsq1, act_dtype=at)
    s_f = ops.scalenorm_scale(rowsq1, p["g_ff"], dim)
    h = torch.empty((N, F), device=dev, dtype=at)
    fpre = torch.empty((N, F), device=dev, dtype=at)
    ops.gemm_(x1, p["w1_c"], out=h, out2=fpre, row_scale=s_f, bias=p["b1"], act=ops.ACT_GELU, act_dtype=at)
    if drop_ff:
        ops.dropout(h, drop_ff[0], drop_ff[1], out=h)       # in place: the saved h is the dropped one (dW2 needs that)
    x2 = torch.empty((N, dim), device=dev, dtype=at)
    rowsq2 = _zeros(pool, N, dev)
    ops.gemm_(h, p["w2_c"], out=x2, bias=p["b2"], res=x1, row_sumsq=rowsq2, act_dtype=at)
    c = EncoderCtx()
    c.x, c.rowsq_x, c.s_a, c.qkv, c.o, c.lse = x2d, rowsq_x, s_a, qkv, o, lse
    c.x1, c.rowsq1, c.s_f, c.fpre, c.h, c.x2, c.rowsq2 = x1, rowsq1, s_f, fpre, h, x2, rowsq2
    c.B, c.S, c.dim = B, S, dim
    c.drop_attn, c.drop_ff = drop_attn, drop_ff
    return c


def _skinny_split(N, d, K):
    """do = dx1 @ Wo has a [N, d] output: with few 128-row tiles and a deep K (time axis: 66 tiles, K = 16 512) the
    machine is half empty.  Those launches accumulate split-K partial sums into a zeroed f32 buffer (the launcher's
    split heuristic) and one small cast produces the bf16 result."""
    tiles = ((N + 127) // 128) * ((d + 127) // 128)
    return tiles < 100 and K >= 4096


def encoder_bwd(c: EncoderCtx, dx2, p, G, heads, d, tag=None, pool=None):
    """dx2: grad wrt the un-normalised encoder output x2 [N,dim] (act dtype).  G: dict key -> f32 grad sink tensor
    (accumulated).  Returns grad wrt the encoder input x [N,dim].  tag: encoder name for the gradient-ready hooks."""
    N, dim = c.x.shape
    dev, at = dx2.device, dx2.dtype
    F = p["w1"].shape[0]
    B, S = c.B, c.S
    # ---- FFN ------------------------------------------------------------------------------------------------
    if "b2" in G:
        ops.colsum(dx2, G["b2"], accumulate=True)
    if "w2" in G:
        ops.gemm_(dx2, c.h, a_mn=True, b_mn=True, out=G["w2"], accumulate=True)                 # dW2 += dx2^T h
    if tag:
        _notify(tag, ("w2", "b2", "g_final"))
    rowdot_f = _zeros(pool, N, dev)
    dfs = torch.empty((N, F), device=dev, dtype=at)      # s_f * dpre
    df = torch.empty((N, F), device=dev, dtype=at)       # dpre
    ops.gemm_(dx2, p["w2_c"], b_mn=True, out=dfs, out2=df, act=ops.ACT_GELU_BWD, aux=c.fpre, aux_bias=p["b1"],
              row_scale2=c.s_f, row_dot=rowdot_f, act_dtype=at)                                # dh = dx2 W2, through GELU'
    if c.drop_ff:
        # dpre = (dx2 W2) * mask/(1-p) * GELU'(pre): the mask commutes with the element-wise GELU' factor, so it is applied
        # to the epilogue's outputs; the row dot <dpre, pre - b1> has to be taken after it
        ops.dropout(df, c.drop_ff[0], c.drop_ff[1], out=df)
        ops.dropout(dfs, c.drop_ff[0], c.drop_ff[1], out=dfs)
        rowdot_f = ops.rowdot_bias(df, c.fpre, p["b1"])
    if "b1" in G:
        ops.colsum(df, G["b1"], accumulate=True)
    if "w1" in G:
        ops.gemm_(dfs, c.x1, a_mn=True, b_mn=True, out=G["w1"], accumulate=True)               # dW1 += (s_f dpre)^T x1
    if "g_ff" in G:
        _acc_g(G["g_ff"], rowdot_f, p["g_ff"])
    if tag:
        _notify(tag, ("g_ff", "w1", "b1"))
    dx1 = torch.empty((N, dim), device=dev, dtype=at)
    ops.gemm_(dfs, p["w1_c"], b_mn=True, out=dx1, res=dx2, cx=c.x1, coef_num=rowdot_f, coef_den=c.rowsq1, act_dtype=at)
    # ---- attention ----------------------------------------------------------------------------------------------
    if "wo" in G:
        ops.gemm_(dx1, c.o.view(N, d), a_mn=True, b_mn=True, out=G["wo"], accumulate=True)     # dWo += dx1^T o
    if at == torch.bfloat16 and _skinny_split(N, d, dim):
        do32 = _zeros(pool, N * d, dev).view(N, d)
        ops.gemm_(dx1, p["wo_c"], b_mn=True, out=do32, accumulate=True)                        # do = dx1 Wo (split-K, f32)
        do = ops.cast(do32, at)
    else:
        do = torch.empty((N, d), device=dev, dtype=at)
        ops.gemm_(dx1, p["wo_c"], b_mn=True, out=do, act_dtype=at)                             # do = dx1 Wo
    dqkv = torch.empty((N, 3 * d), device=dev, dtype=at)
    q3, g3 = c.qkv.view(B, S, 3 * d), dqkv.view(B, S, 3 * d)
    ops.attn_bwd(q3[:, :, :d], q3[:, :, d:2 * d], q3[:, :, 2 * d:], c.o, do.view(B, S, d), c.lse, heads,
                 g3[:, :, :d], g3[:, :, d:2 * d], g3[:, :, 2 * d:], c.drop_attn)
    rowdot_a = ops.rowdot_scale(c.qkv, dqkv, c.s_a)      # dqkv <- s_a * dqkv in place
    if "g_attn" in G:
        _acc_g(G["g_attn"], rowdot_a, p["g_attn"])
    if "wqkv" in G:
        ops.gemm_(dqkv, c.x, a_mn=True, b_mn=True, out=G["wqkv"], accumulate=True)             # dWqkv += (s_a dqkv)^T x
    if tag:
        _notify(tag, ("g_attn", "wqkv", "wo"))
    dx = torch.empty((N, dim), device=dev, dtype=at)
    ops.gemm_(dqkv, p["wqkv_c"], b_mn=True, out=dx, res=dx1, cx=c.x, coef_num=rowdot_a, coef_den=c.rowsq_x, act_dtype=at)
    return dx


def _acc_g(sink, rowdot, g):
    """dg += sum_rows rowdot / g   (ScaleNorm gain; see module docstring)."""
    ops.sum_div_acc(rowdot, g, sink)


# ----------------------------------------------------------------------------------------------------------------------
class DuettEncodeFn(torch.autograd.Function):
    """transformed[B,T+1,(V+1)d] = encode(xs_static, xs_feats, xs_times); parameters are passed positionally so
    autograd tracks them; `spec` describes shapes / names / mode."""

    @staticmethod
    def forward(ctx, spec, xs_feats, tab, te, *params):
        with ops.tf32_mode(spec.get("tf32", False)):
            return DuettEncodeFn._forward(ctx, spec, xs_feats, tab, te, *params)

    @staticmethod
    def backward(ctx, gout):
        with ops.tf32_mode(ctx.spec.get("tf32", False)):
            return DuettEncodeFn._backward(ctx, gout)

    @staticmethod
    def _forward(ctx, spec, xs_feats, tab, te, *params):
        names = spec["names"]
        P = dict(zip(names, params))
        cfgd, V, T, L = spec["d"], spec["V"], spec["T"], spec["n_layers"]
        heads, at, training = spec["heads"], spec["act_dtype"], spec["training"]
        B = xs_feats.shape[0]
        T1, V1 = T + 1, V + 1
        E, Ep = T1 * cfgd, V1 * cfgd
        det = {k: v.detach() for k, v in P.items()}
        psi0, emean, erstd, ehn = ops.embed_fwd(xs_feats, V, cfgd, det["emb.w0"], det["emb.b0"], det["emb.bn_w"],
                                                det["emb.bn_b"], spec["emb_rm"], spec["emb_rv"], det["emb.w4"], det["emb.b4"],
                                                det["n_obs_embedding.weight"].view(-1), det["special_embeddings.weight"],
                                                tab.detach().contiguous(), at, training, return_hidden=True)
        encs = []
        src, src_rowsq, g_prev = psi0, None, None
        pool = ZeroPool(2 * L * (B * V1 + B * T1) + 64, xs_feats.device)
        for l in range(L):
            pe = _enc_params(det, f"event_transformers.{l}", at, P)
            x_e, rsq = ops.relayout_fwd(src, B, T1, V1, cfgd, src_rowsq=src_rowsq, g=g_prev,
                                        pos_bcast=det["full_event_embedding.weight"])
            dp = float(spec.get("dropout", 0.0)) if training else 0.0
            ce = encoder_fwd(x_e.view(B, V1, E), rsq, B, V1, E, pe, heads, cfgd, drop_p=dp, tag=f"event_transformers.{l}",
                             pool=pool)
            pt = _enc_params(det, f"time_transformers.{l}", at, P)
            x_t, rsq = ops.relayout_fwd(ce.x2.view(B, V1, T1, cfgd), B, V1, T1, cfgd,
                                        src_rowsq=ce.rowsq2 if spec["final_norm"] else None, g=pe["g_final"],
                                        pos_batched=te.detach())
            ct = encoder_fwd(x_t.view(B, T1, Ep), rsq, B, T1, Ep, pt, heads, cfgd, drop_p=dp, tag=f"time_transformers.{l}",
                             pool=pool)
            encs.append((ce, ct))
            src, src_rowsq, g_prev = ct.x2.view(B, T1, V1, cfgd), (ct.rowsq2 if spec["final_norm"] else None), pt["g_final"]
        out, _ = ops.relayout_fwd(src.view(B * T1, 1, 1, Ep), B * T1, 1, 1, Ep, src_rowsq=src_rowsq, g=g_prev,
                                  want_rowsq=False)
        ctx.spec, ctx.encs, ctx.P, ctx.B, ctx.det = spec, encs, P, B, det
        ctx.emb_saved = (xs_feats, emean, erstd, ehn)
        ctx.te_requires_grad = te.requires_grad
        ctx.tab_requires_grad = tab.requires_grad
        return out.view(B, T1, Ep)

    @staticmethod
    def _backward(ctx, gout):
        spec, encs, P, B = ctx.spec, ctx.encs, ctx.P, ctx.B
        names = spec["names"]
        cfgd, V, T, L = spec["d"], spec["V"], spec["T"], spec["n_layers"]
        heads, at, training = spec["heads"], spec["act_dtype"], spec["training"]
        T1, V1 = T + 1, V + 1
        E, Ep = T1 * cfgd, V1 * cfgd
        dev = gout.device
        det = ctx.det       # detached parameters + act-dtype weight copies made by the forward
        sinks, rets = {}, {}
        for i, n in enumerate(names):
            if ctx.needs_input_grad[4 + i]:
                sinks[n], rets[n] = grad_sink(P[n])
        gout = gout.contiguous()
        if gout.dtype != at:
            gout = ops.cast(gout, at)
        fn = spec["final_norm"]
        # final ScaleNorm of the last time encoder (identity layout)
        ce, ct = encs[-1]
        gname = f"time_transformers.{L - 1}.g_final"
        dx2 = ops.relayout_bwd(gout.view(B * T1, 1, 1, Ep), B * T1, 1, 1, Ep, src=ct.x2 if fn else None,
                               src_rowsq=ct.rowsq2 if fn else None, g=det[gname] if fn else None,
                               dg=sinks.get(gname) if fn else None).view(B * T1, Ep)
        dx_ts = []     # time-encoder input gradients of every layer: their sum is the time-embedding gradient
        pool = ZeroPool(L * (B * V1 + B * T1) * (1 + cfgd) + 64, dev)       # row dots + split-K `do` partial sums
        for l in reversed(range(L)):
            ce, ct = encs[l]
            pt = _enc_params(det, f"time_transformers.{l}", at)
            dx_t = encoder_bwd(ct, dx2, pt, _enc_sinks(sinks, f"time_transformers.{l}"), heads, cfgd,
                               tag=f"time_transformers.{l}", pool=pool)                                 # [B*T1, Ep]
            if ctx.te_requires_grad:
                dx_ts.append(dx_t)
            # time-major grad -> event-major grad of the event encoder's un-normalised output (+ its final norm)
            gname = f"event_transformers.{l}.g_final"
            dx2e = ops.relayout_bwd(dx_t.view(B, T1, V1, cfgd), B, V1, T1, cfgd, src=ce.x2 if fn else None,
                                    src_rowsq=ce.rowsq2 if fn else None, g=det[gname] if fn else None,
                                    dg=sinks.get(gname) if fn else None).view(B * V1, E)
            pe = _enc_params(det, f"event_transformers.{l}", at)
            dx_e = encoder_bwd(ce, dx2e, pe, _enc_sinks(sinks, f"event_transformers.{l}"), heads, cfgd,
                               tag=f"event_transformers.{l}", pool=pool)                                # [B*V1, E]
            if "full_event_embedding.weight" in sinks:
                ops.colsum(dx_e.view(B, V1 * E), sinks["full_event_embedding.weight"].view(-1), accumulate=True)
            if l > 0:
                pce, pct = encs[l - 1]
                gname = f"time_transformers.{l - 1}.g_final"
                dx2 = ops.relayout_bwd(dx_e.view(B, V1, T1, cfgd), B, T1, V1, cfgd, src=pct.x2 if fn else None,
                                       src_rowsq=pct.rowsq2 if fn else None, g=det[gname] if fn else None,
                                       dg=sinks.get(gname) if fn else None).view(B * T1, Ep)
            else:
                dpsi0 = ops.relayout_bwd(dx_e.view(B, V1, T1, cfgd), B, T1, V1, cfgd)
        dte = None
        if ctx.te_requires_grad:
            dte = (dx_ts[0] if len(dx_ts) == 1 else ops.sum_n(dx_ts) if len(dx_ts) <= 8 else None)
            if dte is None:       # more than 8 layers: sum in groups
                dte = ops.sum_n(dx_ts[:8])
                for i in range(8, len(dx_ts), 7):
                    dte = ops.sum_n([dte] + dx_ts[i:i + 7])
            dte = dte.view(B, T1, Ep)
        dx_ts = None
        # embedding backward
        xs_feats, emean, erstd, ehn = ctx.emb_saved
        z = lambda n: sinks[n] if n in sinks else torch.zeros_like(P[n], dtype=torch.float32)
        eg = {"dW0": z("emb.w0"), "db0": z("emb.b0"), "dgamma": z("emb.bn_w"), "dbeta": z("emb.bn_b"), "dW4": z("emb.w4"),
              "db4": z("emb.b4"), "dnobs": z("n_obs_embedding.weight").view(-1), "dspecial": z("special_embeddings.weight")}
        dtab = ops.embed_bwd(xs_feats, V, cfgd, det["emb.w0"], det["emb.b0"], det["emb.bn_w"], det["emb.bn_b"],
                             det["emb.w4"], det["n_obs_embedding.weight"].view(-1), emean, erstd, dpsi0, eg, training, hn=ehn)
        ctx.encs = ctx.det = ctx.emb_saved = None
        grads = tuple(rets.get(n) for n in names)
        return (None, None, dtab if ctx.tab_requires_grad else None, dte) + grads


def _enc_params(det, prefix, at, P=None):
    """Encoder parameters + their act-dtype copies ('wqkv_c', ...).  bf16 mode: a parameter living in ddp.FlatParams with
    bf16 shadows enabled hands out its shadow view (refreshed by the optimizer kernel) instead of being cast here."""
    p = {k: det[f"{prefix}.{k}"] for k in ENC_KEYS}
    for k in ("wqkv", "wo", "w1", "w2"):
        ck = f"{prefix}.{k}@{at}"
        if ck not in det:
            sh = None
            if P is not None and at == torch.bfloat16:
                src = P[f"{prefix}.{k}"]
                sh = getattr(src, "_dx_shadow", None)
                if sh is not None and getattr(src, "_dx_shadow_version", -1) != src._version:
                    sh = None
            det[ck] = sh if sh is not None else ops.cast(p[k], at)
        p[k + "_c"] = det[ck]
    return p


def _enc_sinks(sinks, prefix):
    return {k: sinks[f"{prefix}.{k}"] for k in ENC_KEYS if f"{prefix}.{k}" in sinks}


class TimeEmbedAssembleFn(torch.autograd.Function):
    """te[b,t,:] = hid[b,t,:] @ W3^T + b3 for t < T ; te[b,T,:] = full_rep_embedding.weight[:,0]
    (second half of cve + the [REP] row, duett/duett.py:151-157,269-272) -> [B,T+1,E'] in the act dtype.
    hid: [B*T, h] f32 (output of Linear(1,h) -> tanh -> BatchNorm)."""

    @staticmethod
    def forward(ctx, hid, W3, b3, rep_w, B, T, at):
        h = hid.shape[1]
        Ep = W3.shape[0]
        T1 = T + 1
        hp = h if at == torch.float32 else ((h + 7) // 8) * 8      # TMA needs 16 B row pitch
        hid_pad = torch.zeros((B, T1, hp), device=hid.device, dtype=at)
        hid_pad[:, :T, :h].copy_(hid.detach().view(B, T, h))       # layout glue (tiny tensor)
        W3c = torch.zeros((Ep, hp), device=hid.device, dtype=at)
        W3c[:, :h].copy_(W3.detach())
        te = torch.empty((B * T1, Ep), device=hid.device, dtype=at)
        ops.gemm_(hid_pad.view(B * T1, hp), W3c, out=te, bias=b3.detach(), act_dtype=at)
        rep = rep_w.detach().view(1, Ep).expand(B, Ep).contiguous()
        off = (torch.arange(B, device=hid.device, dtype=torch.int64) * T1 + T) * Ep
        ops.scatter_vec(rep, off, te, accumulate=False)
        ctx.save_for_backward(hid_pad, W3c, W3, b3, rep_w)
        ctx.dims = (B, T, h, hp, Ep)
        return te.view(B, T1, Ep)

    @staticmethod
    def backward(ctx, dte):
        hid_pad, W3c, W3, b3, rep_w = ctx.saved_tensors
        B, T, h, hp, Ep = ctx.dims
        T1 = T + 1
        at = hid_pad.dtype
        dte = dte.contiguous()
        d2 = dte.view(B * T1, Ep)
        ghid = gW3 = gb3 = grep = None
        rep_rows = dte.view(B, T1 * Ep)[:, T * Ep:]                 # [B, Ep] strided view of the [REP] rows
        if ctx.needs_input_grad[3]:
            sink, grep = grad_sink(rep_w)
            ops.colsum(rep_rows, sink.view(-1), accumulate=True)
        if ctx.needs_input_grad[2]:
            sink, gb3 = grad_sink(b3)
            ops.colsum(d2, sink, accumulate=True)
            neg = torch.zeros(Ep, device=dte.device, dtype=torch.float32)
            ops.colsum(rep_rows, neg, accumulate=True)
            ops.axpy_f32(neg, sink, -1.0)
        if ctx.needs_input_grad[1]:
            sink, gW3 = grad_sink(W3)
            if hp == h:
                ops.gemm_(d2, hid_pad.view(B * T1, hp), a_mn=True, b_mn=True, out=sink, accumulate=True)
            else:
                tmp = torch.empty((Ep, hp), device=dte.device, dtype=torch.float32)
                ops.gemm_(d2, hid_pad.view(B * T1, hp), a_mn=True, b_mn=True, out=tmp)
                sink.add_(tmp[:, :h])                                 # unpad (tiny tensor)
        if ctx.needs_input_grad[0]:
            gh = torch.empty((B * T1, hp), device=dte.device, dtype=at)
            ops.gemm_(d2, W3c, b_mn=True, out=gh, act_dtype=at)
            ghid = gh.view(B, T1, hp)[:, :T, :h].float().reshape(B * T, h)
        return ghid, gW3, gb3, grep, None, None, None
