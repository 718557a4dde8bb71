"""TEST INFRASTRUCTURE ONLY — plain-torch restatement of the CONTRACT of every kernel wrapper in
multimodal_edema_prediction_b200/ops.py.

Two uses:
  * CPU tests (`-m "not gpu"`): `install()` monkeypatches the wrappers so the host orchestration (backbone.py,
    functional.py, the nn.Modules, state-dict mapping, engine) runs on CPU tensors and is checked against the oracle —
    this validates the fused-backward algebra without a GPU.
  * GPU tests (`-m gpu`): each CUDA kernel is compared against the function of the same name here.
Backward kernels are emulated by autograd through the emulated forward, so they are independent of the hand-derived
formulas in the CUDA code.  Nothing in the product imports this file.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

ACT_NONE, ACT_GELU, ACT_RELU, ACT_TANH, ACT_GELU_BWD, ACT_RELU_BWD, ACT_TANH_BWD = range(7)


def _f(t):
    return None if t is None else t.float()


def _gelu_grad(x):
    return 0.5 * (1 + torch.erf(x / math.sqrt(2))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)


def gemm_(a, b, *, a_mn=False, b_mn=False, out=None, out2=None, accumulate=False, act=ACT_NONE, act_dtype=None,
          row_scale=None, row_scale2=None, bias=None, res=None, aux=None, aux_bias=None, cx=None, coef_num=None,
          coef_den=None, row_sumsq=None, row_dot=None, force_simt=False, split_k=0, tf32=False):
    A = a.float().transpose(-1, -2) if a_mn else a.float()
    Bm = b.float().transpose(-1, -2) if b_mn else b.float()
    v = A @ Bm.transpose(-1, -2)
    if v.dim() == 3:      # grouped mode: plain / bias / accumulate epilogues only (what the embedding uses)
        assert act == ACT_NONE and res is None and cx is None and aux is None and row_scale is None
        if bias is not None:
            v = v + bias[:, None, :]
        if accumulate:
            out += v
        else:
            out.copy_(v)
        return
    if row_scale is not None:
        v = v * row_scale[:, None]
    if bias is not None:
        v = v + bias[None, :]
    if act == ACT_GELU:
        if out2 is not None:
            out2.copy_(v)
        v = F.gelu(v)
    elif act == ACT_RELU:
        v = torch.relu(v)
    elif act == ACT_TANH:
        v = torch.tanh(v)
    elif act == ACT_GELU_BWD:
        ax = aux.float()
        v = v * _gelu_grad(ax)
        if row_dot is not None:
            row_dot += (v * (ax - (aux_bias[None, :] if aux_bias is not None else 0))).sum(1)
        if out2 is not None:
            out2.copy_(v)
        if row_scale2 is not None:
            v = v * row_scale2[:, None]
    elif act == ACT_RELU_BWD:
        v = v * (aux.float() > 0)
    elif act == ACT_TANH_BWD:
        v = v * (1 - aux.float() ** 2)
    if res is not None:
        v = v + res.float()
    if cx is not None:
        v = v - cx.float() * (coef_num / coef_den.clamp_min(1e-24))[:, None]
    if row_sumsq is not None:
        row_sumsq += (v * v).sum(1)
    if out is not None:
        if accumulate:
            out += v
        else:
            out.copy_(v)


def _final_scale(src_rowsq, g, dim):
    return math.sqrt(dim) * g[0] / src_rowsq.sqrt().clamp_min(1e-12)


def relayout_fwd(src, B, P, Q, d, *, src_rowsq=None, g=None, pos_bcast=None, pos_batched=None, want_rowsq=True):
    x = src.float().reshape(B, P, Q, d)
    if src_rowsq is not None:
        x = x * _final_scale(src_rowsq, g, Q * d).reshape(B, P, 1, 1)
    y = x.permute(0, 2, 1, 3).reshape(B, Q, P * d)
    if pos_bcast is not None:
        y = y + pos_bcast.float().reshape(1, Q, P * d)
    if pos_batched is not None:
        y = y + pos_batched.float().reshape(B, Q, P * d)
    rowsq = (y * y).sum(-1).reshape(B * Q) if want_rowsq else None
    return y.reshape(B, Q, P, d).to(src.dtype).contiguous(), rowsq


def relayout_bwd(gdst, B, P, Q, d, *, src=None, src_rowsq=None, g=None, dg=None):
    gy = gdst.float().reshape(B, Q, P, d).permute(0, 2, 1, 3).reshape(B * P, Q * d)
    if src_rowsq is None:
        return gy.reshape(B, P, Q, d).to(gdst.dtype).contiguous()
    x = src.float().reshape(B * P, Q * d)
    nsq = src_rowsq.clamp_min(1e-24)
    c = math.sqrt(Q * d)
    dot = (gy * x).sum(1)
    s = c * g[0] / nsq.sqrt()
    if dg is not None:
        dg += (dot * c / nsq.sqrt()).sum()
    dx = s[:, None] * (gy - x * (dot / nsq)[:, None])
    return dx.reshape(B, P, Q, d).to(gdst.dtype).contiguous()


def colsum(x2d, out, accumulate=True):
    s = x2d.float().sum(0)
    if accumulate:
        out += s
    else:
        out.copy_(s)


def axpy(x, y, alpha=1.0, accumulate=True):
    xf = x.float().reshape(y.shape)
    if accumulate:
        y.copy_((y.float() + alpha * xf).to(y.dtype))
    else:
        y.copy_((alpha * xf).to(y.dtype))


def axpy_f32(x, y, alpha):
    y += alpha * x


def cast(x, dtype):
    return x if x.dtype == dtype else x.to(dtype)


def scalenorm_scale(rowsq, g, dim):
    return math.sqrt(dim) * g[0] / rowsq.sqrt().clamp_min(1e-12)


def rowdot_scale(a, g, row_scale):
    rd = (a.float() * g.float()).sum(1)
    if row_scale is not None:
        g.copy_((g.float() * row_scale[:, None]).to(g.dtype))
    return rd


def keep_factor(seed, n, p, step=0):
    """Restatement of dx_rng32 / dx_drop_factor (csrc/dx_common.cuh): float32 tensor [n] of keep/(1-p) over flat indices."""
    import numpy as np
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64)
        z = (np.uint64((int(seed) + int(step)) & 0xFFFFFFFFFFFFFFFF) + idx * np.uint64(0x9E3779B97F4A7C15)) & M
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M
        z = z ^ (z >> np.uint64(31))
        r = (z >> np.uint64(32)).astype(np.uint64)
    thresh = min(int(float(np.float32(p)) * 4294967296.0), 4294967295)
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return torch.from_numpy(np.where(r >= thresh, scale, np.float32(0)).astype(np.float32))


def dropout(x, p, seed, out=None):
    y = (x.float() * keep_factor(seed, x.numel(), p).reshape(x.shape)).to(x.dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def rowdot_bias(a, b, bias):
    bb = b.float() - (bias.float() if bias is not None else 0.0)
    return (a.float() * bb).sum(1)


def _attn(q, k, v, heads, drop=None):
    B, Sq, D = q.shape
    dh = D // heads
    qh = q.float().reshape(B, Sq, heads, dh).transpose(1, 2)
    kh = k.float().reshape(B, -1, heads, dh).transpose(1, 2)
    vh = v.float().reshape(B, -1, heads, dh).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(dh)
    lse = torch.logsumexp(s, -1)
    pr = torch.softmax(s, -1)
    if drop:
        pr = pr * keep_factor(drop[1], pr.numel(), drop[0]).reshape(pr.shape)      # index ((b*H+h)*Sq+q)*Sk+k
    o = (pr @ vh).transpose(1, 2).reshape(B, Sq, D)
    return o, lse


def attn_fwd(q, k, v, heads, drop=None):
    o, lse = _attn(q, k, v, heads, drop)
    return o.to(q.dtype).contiguous(), lse.contiguous()


def attn_probs_mean(q, k, lse, heads):
    """Contract of dx_attn_probs_mean: mean over heads of exp(q k^T / sqrt(dh) - lse)."""
    B, Sq, D = q.shape
    dh = D // heads
    qh = q.float().reshape(B, Sq, heads, dh).transpose(1, 2)
    kh = k.float().reshape(B, -1, heads, dh).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(dh)
    return torch.exp(s - lse[..., None]).mean(1)


def attn_bwd(q, k, v, o, go, lse, heads, dq, dk, dv, drop=None):
    with torch.enable_grad():
        qq, kk, vv = (t.detach().float().clone().requires_grad_(True) for t in (q, k, v))
        oo, _ = _attn(qq, kk, vv, heads, drop)
        gq, gk, gv = torch.autograd.grad(oo, (qq, kk, vv), go.float())
    dq.copy_(gq); dk.copy_(gk); dv.copy_(gv)


def _embed(xs, V, d, W0, b0, gamma, beta, W4, b4, nobs, special, tab, mean, rstd):
    B, T, _ = xs.shape
    val, cnt, step = xs[:, :, :V], xs[:, :, V:2 * V], xs[:, :, 2 * V]
    ce = nobs[cnt.long().clamp(0, 15)]
    h = torch.relu(val[..., None] * W0[None, None, :, :, 0] + ce[..., None] * W0[None, None, :, :, 1] + b0)   # [B,T,V,64]
    if mean is None:
        mean = h.mean((0, 1))
        var = h.var((0, 1), unbiased=False)
        rstd = torch.rsqrt(var + 1e-5)
    hn = (h - mean) * rstd * gamma + beta
    emb = torch.einsum("btvh,vdh->btvd", hn, W4) + b4
    psi = torch.zeros(B, T + 1, V + 1, d, dtype=emb.dtype)
    psi[:, :T, :V] = emb
    psi[:, :T, V] = tab[:, None, :]
    psi[:, T] = special[1]
    m = torch.zeros(B, T + 1, V + 1, dtype=torch.bool)
    m[:, :T, :] |= (step == 1)[:, :, None]
    m[:, :T, :V] |= cnt == -1
    m[:, T, :V] |= cnt[:, 0] == -1
    psi = torch.where(m[..., None], special[0], psi)
    return psi, mean, rstd, h


def embed_fwd(xs, V, d, W0, b0, gamma, beta, run_mean, run_var, W4, b4, nobs, special, tab, act_dtype, training,
              return_hidden=False):
    if training:
        psi, mean, rstd, h = _embed(xs, V, d, W0, b0, gamma, beta, W4, b4, nobs, special, tab, None, None)
        if run_mean is not None:
            R = xs.shape[0] * xs.shape[1]
            var = h.var((0, 1), unbiased=False)
            run_mean.mul_(0.9).add_(0.1 * mean)
            run_var.mul_(0.9).add_(0.1 * var * R / max(R - 1, 1))
    else:
        mean, rstd = run_mean.clone(), torch.rsqrt(run_var + 1e-5)
        psi, _, _, _ = _embed(xs, V, d, W0, b0, gamma, beta, W4, b4, nobs, special, tab, mean, rstd)
    if return_hidden:
        return psi.to(act_dtype).contiguous(), mean.contiguous(), rstd.contiguous(), None
    return psi.to(act_dtype).contiguous(), mean.contiguous(), rstd.contiguous()


def embed_bwd(xs, V, d, W0, b0, gamma, beta, W4, nobs, mean, rstd, dpsi, grads, training, hn=None):
    B = xs.shape[0]
    with torch.enable_grad():
        leaves = [t.detach().clone().requires_grad_(True) for t in (W0, b0, gamma, beta, W4)]
        b4 = torch.zeros(V, d, requires_grad=True)
        nb = nobs.detach().clone().requires_grad_(True)
        sp = torch.zeros(8, d, requires_grad=True)
        tab = torch.zeros(B, d, requires_grad=True)
        psi, _, _, _ = _embed(xs, V, d, leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], b4, nb, sp, tab,
                              None if training else mean, None if training else rstd)
        g = torch.autograd.grad(psi, leaves + [b4, nb, sp, tab], dpsi.float(), allow_unused=True)
    z = lambda t, like: torch.zeros_like(like) if t is None else t
    grads["dW0"] += z(g[0], W0); grads["db0"] += z(g[1], b0); grads["dgamma"] += z(g[2], gamma)
    grads["dbeta"] += z(g[3], beta); grads["dW4"] += z(g[4], W4); grads["db4"] += z(g[5], b4)
    grads["dnobs"] += z(g[6], nb); grads["dspecial"] += z(g[7], sp)
    return z(g[8], tab).contiguous()


def bn2d_fwd(x, w, b, run_mean, run_var, training):
    R = x.shape[0]
    if training:
        mean, var = x.mean(0), x.var(0, unbiased=False)
        if run_mean is not None:
            run_mean.mul_(0.9).add_(0.1 * mean)
            run_var.mul_(0.9).add_(0.1 * var * R / max(R - 1, 1))
        rstd = torch.rsqrt(var + 1e-5)
    else:
        mean, rstd = run_mean.clone(), torch.rsqrt(run_var + 1e-5)
    return (x - mean) * rstd * w + b, mean, rstd


def bn2d_bwd(dy, x, w, mean, rstd, dw, db, training, need_dx=True):
    xh = (x - mean) * rstd
    s1, s2 = dy.sum(0), (dy * xh).sum(0)
    db += s1
    dw += s2
    if not need_dx:
        return None
    R = x.shape[0]
    if training:
        return w * rstd * (dy - s1 / R - xh * s2 / R)
    return w * rstd * dy


def layernorm_fwd(x2d, w, b):
    x = x2d.float()
    mean, var = x.mean(1), x.var(1, unbiased=False)
    rstd = torch.rsqrt(var + 1e-5)
    y = (x - mean[:, None]) * rstd[:, None] * w + b
    return y.to(x2d.dtype), mean, rstd


def layernorm_bwd(dy, x2d, w, mean, rstd, dw, db, need_dx=True):
    x, g = x2d.float(), dy.float()
    xh = (x - mean[:, None]) * rstd[:, None]
    dw += (g * xh).sum(0)
    db += g.sum(0)
    if not need_dx:
        return None
    gw = g * w
    dx = rstd[:, None] * (gw - gw.mean(1, keepdim=True) - xh * (gw * xh).mean(1, keepdim=True))
    return dx.to(x2d.dtype)


def mean_rows(x3d, T):
    return x3d[:, :T].float().mean(1).contiguous()


def mean_rows_bwd(dy, T1, T, dtype):
    B, E = dy.shape
    dx = torch.zeros(B, T1, E, dtype=dtype)
    dx[:, :T] = (dy / T)[:, None, :].to(dtype)
    return dx


def gather_vec(src, offsets, Lv):
    """A negative offset selects nothing: zero row."""
    flat = src.reshape(-1)
    idx = offsets.clamp_min(0)[:, None] + torch.arange(Lv)[None, :]
    return flat[idx].float() * (offsets >= 0)[:, None]


def scatter_vec(src, offsets, dst, accumulate=True):
    """Rows with a negative offset are skipped."""
    flat = dst.view(-1)
    ok = offsets >= 0
    idx = (offsets[ok][:, None] + torch.arange(src.shape[1])[None, :]).reshape(-1)
    vals = src[ok].reshape(-1)
    if accumulate:
        flat[idx] = (flat[idx].float() + vals).to(dst.dtype)
    else:
        flat[idx] = vals.to(dst.dtype)


def _with_grad(fn, z, *args):
    with torch.enable_grad():
        zz = z.detach().clone().requires_grad_(True)
        out = fn(zz, *args)
        main = out[0] if isinstance(out, tuple) else out
        (dz,) = torch.autograd.grad(main, zz)
    return out, dz


def kd_loss(zs, zt, y, T, alpha, pos_weight, need_grad=True, eps=1e-7):
    def fn(z):
        pt = torch.sigmoid(zt / T).clamp(eps, 1 - eps)
        ps = torch.sigmoid(z / T).clamp(eps, 1 - eps)
        kd = T * T * (pt * (pt.log() - ps.log()) + (1 - pt) * ((1 - pt).log() - (1 - ps).log())).mean()
        pw = None if pos_weight is None else torch.tensor([pos_weight])
        bce = F.binary_cross_entropy_with_logits(z, y, pos_weight=pw)
        return alpha * bce + (1 - alpha) * kd, bce, kd
    (tot, bce, kd), dz = _with_grad(fn, zs)
    return torch.stack([tot, bce, kd]).detach(), dz


def bce_logits(z, y, w_pos=1.0, w_neg=1.0, need_grad=True):
    w = torch.where(y > 0, torch.tensor(w_pos), torch.tensor(w_neg))
    out, dz = _with_grad(lambda zz: F.binary_cross_entropy_with_logits(zz, y, w), z)
    return out.detach().reshape(1), dz


def masked_mse_bce(yhat, phat, y, m, w_presence, out2, need_grad=True):
    with torch.enable_grad():
        a, b = yhat.detach().clone().requires_grad_(True), phat.detach().clone().requires_grad_(True)
        l1 = F.mse_loss(a * m, y * m)
        l2 = F.binary_cross_entropy_with_logits(b, m) * w_presence
        d1, d2 = torch.autograd.grad(l1 + l2, (a, b))
    out2[0] += l1.detach()
    out2[1] += l2.detach()
    return d1, d2


def masked_bce_cols(z, y, m, pos_weight, coef, eps, need_grad=True):
    def fn(zz):
        l = F.binary_cross_entropy_with_logits(zz, y, reduction="none", pos_weight=pos_weight)
        return (l * m).sum(0) / (m.sum(0) + eps)
    with torch.enable_grad():
        zz = z.detach().clone().requires_grad_(True)
        per = fn(zz)
        c = torch.ones_like(per) if coef is None else coef
        (dz,) = torch.autograd.grad((per * c).sum(), zz)
    return per.detach(), dz


def aux_residual_kl(img_logits, scaled_corr, y, mask, eps=0.05, need_grad=True):
    def fn(c):
        ys = y * (1 - eps) + (1 - y) * eps
        p = torch.sigmoid(img_logits + c).clamp(1e-6, 1 - 1e-6)
        kl = ys * (ys.log() - p.log()) + (1 - ys) * ((1 - ys).log() - (1 - p).log())
        return (kl * mask).sum() / mask.sum().clamp(min=1.0)
    out, dc = _with_grad(fn, scaled_corr)
    return out.detach().reshape(1), dc


def require_device(t):
    pass


def act_fwd(x, act):
    xf = x.float()
    r = torch.relu(xf) if act == ACT_RELU else (torch.tanh(xf) if act == ACT_TANH else F.gelu(xf))
    return r.to(x.dtype)


def act_bwd(g, aux, code):
    a, gg = aux.float(), g.float()
    if code == ACT_RELU_BWD:
        r = gg * (a > 0)
    elif code == ACT_TANH_BWD:
        r = gg * (1 - a * a)
    else:
        r = gg * _gelu_grad(a)
    return r.to(g.dtype)


def scale_dev(x, s):
    s = s.reshape(-1).float()
    if s.numel() == 1:
        return x * s[0]
    return (x.reshape(-1, s.numel()) * s[None, :]).reshape(x.shape)


def sum_div_acc(x, g, sink):
    sink += x.sum() / g[0]


def fusion_logits(hi, ht, corr, bias_i, bias_t, beta):
    img = hi + bias_i[None]
    return img, ht + bias_t[None], beta[None] * corr, img + beta[None] * corr


def fusion_logits_bwd(d_img, d_ts, d_scaled, d_fus, corr, beta, dbeta, dbias_i, dbias_t):
    z = torch.zeros_like(corr)
    gs = (z if d_scaled is None else d_scaled) + (z if d_fus is None else d_fus)
    dbeta += (gs * corr).sum(0)
    if d_img is not None:
        dbias_i += d_img.sum(0)
    if d_ts is not None:
        dbias_t += d_ts.sum(0)
    return gs * beta[None]


def adamw(p, g, m, v, lr, betas, eps, weight_decay, step, grad_scale_dev=None, grad_scale=1.0, step_dev=None,
          lr_scale_dev=None, shadow=None):
    if step_dev is not None:
        step = int(step_dev[0])
    if lr_scale_dev is not None:
        lr = lr * float(lr_scale_dev[0])
    gs = grad_scale * (float(grad_scale_dev[0]) if grad_scale_dev is not None else 1.0)
    gi = g * gs
    p.mul_(1 - lr * weight_decay)
    m.mul_(betas[0]).add_((1 - betas[0]) * gi)
    v.mul_(betas[1]).add_((1 - betas[1]) * gi * gi)
    bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
    p.sub_((lr / bc1) * m / (v.sqrt() / math.sqrt(bc2) + eps))
    if shadow is not None:
        shadow.copy_(p)


def ssl_mask(xs, step, ev, keep):
    """Contract of dx_ssl_mask = the reference's index chain (duett/duett.py:198-233), written with torch indexing; a 2-D
    step [B,K] is the pretrain_masked_steps > 1 branch."""
    B, T, C = xs.shape
    V = (C - 1) // 2
    ar, st = torch.arange(B), step.long()
    if st.dim() == 2:
        ar = ar[:, None]
    y_ts = xs[ar, st, :V].clone()
    y_mask = xs[ar, st, V:2 * V].clip(0, 1)
    xc = xs.clone()
    xc[ar, st, :] = 0.
    xc[ar, st, -1] = 1.
    ar = torch.arange(B)
    y_ev = y_ev_mask = None
    if ev is not None:
        e = ev.long()
        y_ev = xs[ar, :, e].clone()
        y_ev_mask = xs[ar, :, e + V].clip(0, 1)
        xc[ar, :, e] = 0.
        xc[ar, :, e + V] = -1.
    if keep is not None:
        k = torch.logical_or(1 - (y_mask.sum(dim=1).clip(0, 1) if st.dim() == 2 else y_mask), keep.bool())
        k = torch.cat((k.tile(1, 2), torch.ones((B, 1))), dim=1)
        xc = xc * torch.logical_or(k.unsqueeze(1), xc == -1)
    return xc, y_ts, y_mask, y_ev, y_ev_mask


def bin_events(slot, vals, cnts, row_start, means, stds, T):
    """Contract of dx_bin_events = the reference's build_stay_tensor row walk (oracle/binning_oracle.py, pinned by g5)."""
    from oracle import binning_oracle
    return torch.from_numpy(binning_oracle.bin_events(slot.numpy(), vals.numpy(), cnts.numpy(), row_start.numpy(),
                                                      means.numpy(), stds.numpy(), int(T)))


def cast_into(x, y):
    y.copy_(x)
    return y


def sumsq(x, out):
    out += (x * x).sum()


def clip_factor(sumsq_t, max_norm, clip):
    clip[0] = min(1.0, max_norm / (float(sumsq_t[0]) ** 0.5 + 1e-6))


def sum_n(tensors):
    acc = tensors[0].float()
    for t in tensors[1:]:
        acc = acc + t.float()
    return acc.to(tensors[0].dtype)


def binary_auc(logits, labels, apply_sigmoid=True):
    """Contract of dx_binary_auc: sklearn's definitions on sigmoid(logits) (training_duett/evaluator.py:22-35)."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    z = logits.reshape(-1).float()
    p = (torch.sigmoid(z) if apply_sigmoid else z).numpy()
    y = (labels.reshape(-1).float() > 0.5).float().numpy()
    n_pos = float(y.sum())
    one_class = n_pos == 0 or n_pos == len(y)
    auroc = float("nan") if one_class else float(roc_auc_score(y, p))
    auprc = float("nan") if n_pos == 0 else float(average_precision_score(y, p))
    return torch.tensor([auroc, auprc, n_pos, float(len(y))], dtype=torch.float64)


EMULATED = ["binary_auc", "sum_n", "dropout", "rowdot_bias", "gemm_", "relayout_fwd", "relayout_bwd", "colsum", "axpy", "axpy_f32", "cast", "scalenorm_scale", "rowdot_scale",
            "attn_fwd", "attn_bwd", "attn_probs_mean", "embed_fwd", "embed_bwd", "bn2d_fwd", "bn2d_bwd", "layernorm_fwd", "layernorm_bwd",
            "mean_rows", "mean_rows_bwd", "gather_vec", "scatter_vec", "kd_loss", "bce_logits", "masked_mse_bce",
            "masked_bce_cols", "aux_residual_kl", "require_device", "act_bwd", "act_fwd", "scale_dev", "sum_div_acc", "fusion_logits",
            "fusion_logits_bwd", "adamw", "sumsq", "clip_factor", "cast_into", "ssl_mask", "bin_events"]


def install(monkeypatch):
    """Route multimodal_edema_prediction_b200.ops through this emulator (CPU tests only)."""
    from multimodal_edema_prediction_b200 import ops
    g = globals()
    for name in EMULATED:
        monkeypatch.setattr(ops, name, g[name])
