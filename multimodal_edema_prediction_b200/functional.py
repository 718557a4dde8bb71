"""autograd.Function wrappers whose forward AND backward are hand-written CUDA kernels (ops.py -> C ABI).

These serve the parts of the path that sit around the fused DuETT backbone (backbone.py): tab_encoder / time-embedding
front / heads (nn.Linear + BatchNormLastDim stacks, duett/duett.py:24-39), the perceiver fusion head
(models/main_architecture_duett.py:536-774) and the losses (loss/losses_duett.py, duett/duett.py:337-365).
Parameter gradients are produced by the kernels; torch only routes tensors.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import lib  # noqa: F401  (import fails loudly when the CUDA library is missing)

_BWD_OF = {ops.ACT_RELU: ops.ACT_RELU_BWD, ops.ACT_TANH: ops.ACT_TANH_BWD, ops.ACT_GELU: ops.ACT_GELU_BWD}


def grad_sink(param: torch.Tensor):
    """Where a parameter gradient is written: straight into an existing fp32 .grad (flat-buffer / DDP layout; autograd
    then receives None) or into a fresh zero tensor that is handed back to autograd."""
    if param.grad is not None and param.grad.dtype == torch.float32 and param.grad.is_contiguous():
        return param.grad, None
    g = torch.zeros(param.shape, device=param.device, dtype=torch.float32)
    return g, g


def _dense2d(x, K):
    x2 = x.reshape(-1, K)
    return x2 if x2.is_contiguous() else x2.contiguous()


class LinearFn(torch.autograd.Function):
    """y = act(x @ W^T + b) (+ res) over the last dim.  x: [..., K] f32 or bf16; W: [N,K] f32 master weight.
    bf16 activations run on the tcgen05 kernel (weight cast per call), f32 on the FFMA kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, res):
        K = x.shape[-1]
        x2 = _dense2d(x, K)
        N = weight.shape[0]
        Wc = ops.cast(weight.detach(), x2.dtype)
        if x2.dtype == torch.bfloat16 and K % 8:
            raise ops.L.DxError(f"bf16 Linear needs in_features % 8 == 0 (got {K})")
        M = x2.shape[0]
        r2 = None if res is None else _dense2d(res, N)
        split = 0
        if x2.dtype == torch.float32 and res is None and K >= 2048 and ((M + 63) // 64) * ((N + 127) // 128) <= 32:
            split = min(64, K // 256)          # skinny output, huge K (pooled features -> head): spread K over the SMs
        if split > 1:
            pre = torch.zeros((M, N), device=x.device, dtype=torch.float32)
            ops.gemm_(x2, Wc, out=pre, bias=None if bias is None else bias.detach(), accumulate=True, split_k=split)
            y = pre if act == ops.ACT_NONE else ops.act_fwd(pre, act)
            if act != ops.ACT_GELU:
                pre = None
        else:
            y = torch.empty((M, N), device=x.device, dtype=x2.dtype)
            pre = torch.empty_like(y) if act == ops.ACT_GELU else None
            ops.gemm_(x2, Wc, out=y, out2=pre, bias=None if bias is None else bias.detach(), act=act, res=r2,
                      act_dtype=x2.dtype)
        ctx.act, ctx.lead, ctx.has_bias, ctx.has_res = act, x.shape[:-1], bias is not None, res is not None
        aux = pre if pre is not None else y
        if act != ops.ACT_NONE and res is not None:
            raise ops.L.DxError("LinearFn: activation + residual in one call is not supported")
        ctx.save_for_backward(x2, weight, bias if bias is not None else weight, aux, Wc)
        return y.reshape(*ctx.lead, N)

    @staticmethod
    def backward(ctx, gy):
        x2, weight, bias, aux, Wc = ctx.saved_tensors
        N, K = weight.shape
        g2 = _dense2d(gy, N)
        if g2.dtype != x2.dtype:
            g2 = ops.cast(g2, x2.dtype)
        gres = gy if ctx.has_res else None
        if ctx.act != ops.ACT_NONE:
            g2 = ops.act_bwd(g2, aux, _BWD_OF[ctx.act])
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx2 = torch.empty_like(x2)
            ops.gemm_(g2, Wc, b_mn=True, out=gx2, act_dtype=x2.dtype)          # dX = dY @ W   (W as [K=N_out, N=K_in], MN-major)
            gx = gx2.reshape(*ctx.lead, K)
        if ctx.needs_input_grad[1]:
            sink, gw = grad_sink(weight)
            split, rows = 0, g2.shape[0]
            if x2.dtype == torch.float32 and rows >= 2048 and ((N + 63) // 64) * ((K + 127) // 128) <= 32:
                split = min(64, rows // 256)   # few output tiles, long reduction over rows (e.g. Linear(1, 128) over B*T rows)
            ops.gemm_(g2, x2, a_mn=True, b_mn=True, out=sink, accumulate=True, split_k=split)  # dW += dY^T @ X
        if ctx.has_bias and ctx.needs_input_grad[2]:
            sink, gb = grad_sink(bias)
            ops.colsum(g2, sink, accumulate=True)
        return gx, gw, gb, None, gres


def linear(x, weight, bias=None, act=ops.ACT_NONE, res=None):
    return LinearFn.apply(x, weight, bias, act, res)


class BatchNorm2dFn(torch.autograd.Function):
    """BatchNormLastDim on a [R,C] f32 matrix (duett/duett.py:11-22); running buffers are updated in the kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias, run_mean, run_var, training):
        x = x.contiguous()
        y, mean, rstd = ops.bn2d_fwd(x, weight.detach(), bias.detach(), run_mean, run_var, training)
        ctx.training = training
        ctx.save_for_backward(x, weight, bias, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, bias, mean, rstd = ctx.saved_tensors
        sw, gw = grad_sink(weight)
        sb, gb = grad_sink(bias)
        gx = ops.bn2d_bwd(gy.contiguous(), x, weight, mean, rstd, sw, sb, ctx.training, need_dx=ctx.needs_input_grad[0])
        return gx, gw, gb, None, None, None


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        C = x.shape[-1]
        x2 = _dense2d(x, C)
        y, mean, rstd = ops.layernorm_fwd(x2, weight.detach(), bias.detach())
        ctx.lead = x.shape[:-1]
        ctx.save_for_backward(x2, weight, bias, mean, rstd)
        return y.reshape(*ctx.lead, C)

    @staticmethod
    def backward(ctx, gy):
        x2, weight, bias, mean, rstd = ctx.saved_tensors
        C = x2.shape[-1]
        sw, gw = grad_sink(weight)
        sb, gb = grad_sink(bias)
        gx = ops.layernorm_bwd(_dense2d(gy, C), x2, weight, mean, rstd, sw, sb, need_dx=ctx.needs_input_grad[0])
        return (None if gx is None else gx.reshape(*ctx.lead, C)), gw, gb


def layer_norm(x, weight, bias):
    return LayerNormFn.apply(x, weight, bias)


class MeanRowsFn(torch.autograd.Function):
    """[B,T1,E] -> [B,E] f32 mean over the first T rows (hourly tokens, [REP] excluded)."""

    @staticmethod
    def forward(ctx, x, T):
        ctx.T1, ctx.T, ctx.dtype = x.shape[1], T, x.dtype
        return ops.mean_rows(x.contiguous(), T)

    @staticmethod
    def backward(ctx, gy):
        return ops.mean_rows_bwd(gy.contiguous(), ctx.T1, ctx.T, ctx.dtype), None


class GatherVecFn(torch.autograd.Function):
    """out[i,:] = src.flatten()[off[i] : off[i]+L] as f32; backward scatters into a zero tensor of src's shape."""

    @staticmethod
    def forward(ctx, src, offsets, L):
        src = src.contiguous()
        ctx.shape, ctx.dtype = src.shape, src.dtype
        ctx.save_for_backward(offsets)
        return ops.gather_vec(src, offsets, L)

    @staticmethod
    def backward(ctx, gy):
        (offsets,) = ctx.saved_tensors
        gsrc = torch.zeros(ctx.shape, device=gy.device, dtype=ctx.dtype)
        ops.scatter_vec(gy.contiguous(), offsets, gsrc, accumulate=False)
        return gsrc, None, None


class CastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src = x.dtype
        return ops.cast(x.contiguous(), dtype)

    @staticmethod
    def backward(ctx, gy):
        return ops.cast(gy.contiguous(), ctx.src), None


def cast(x, dtype):
    return x if x.dtype == dtype else CastFn.apply(x, dtype)


class AttentionFn(torch.autograd.Function):
    """dropout(softmax(q k^T / sqrt(dh))) v, no mask. q [B,Sq,D], k/v [B,Sk,D]; last dim dense, other strides free.
    drop_p > 0: dropout on the attention probabilities (nn.MultiheadAttention(dropout=...) in training mode)."""

    @staticmethod
    def forward(ctx, q, k, v, heads, drop_p=0.0, return_lse=False):
        drop = (drop_p, ops.drop_seed("attn", drop_p, (q.shape[0], heads, q.shape[1], k.shape[1]))) if drop_p > 0 else None
        o, lse = ops.attn_fwd(q, k, v, heads, drop)
        ctx.heads, ctx.drop = heads, drop
        ctx.save_for_backward(q, k, v, o, lse)
        if return_lse:               # row log-sum-exp [B,h,Sq] for ops.attn_probs_mean (attention-map visualisation)
            ctx.mark_non_differentiable(lse)
            return o, lse
        return o

    @staticmethod
    def backward(ctx, go, *_):
        q, k, v, o, lse = ctx.saved_tensors
        go = go.contiguous()
        dq, dk, dv = torch.empty_like(q, memory_format=torch.contiguous_format), \
            torch.empty_like(k, memory_format=torch.contiguous_format), \
            torch.empty_like(v, memory_format=torch.contiguous_format)
        ops.attn_bwd(q, k, v, o, go, lse, ctx.heads, dq, dk, dv, ctx.drop)
        return dq, dk, dv, None, None, None


class DropoutFn(torch.autograd.Function):
    """nn.Dropout in training mode: y = x * keep / (1-p); the backward regenerates the mask from the saved seed."""

    @staticmethod
    def forward(ctx, x, p, tag):
        x = x.contiguous()
        ctx.p, ctx.seed = p, ops.drop_seed(tag, p, x.shape)
        return ops.dropout(x, p, ctx.seed)

    @staticmethod
    def backward(ctx, gy):
        return ops.dropout(gy.contiguous(), ctx.p, ctx.seed), None, None


def dropout(x, p, training, tag="dropout"):
    if not training or p <= 0:
        return x
    if p >= 1:
        raise ops.L.DxError("dropout probability must be < 1")
    return DropoutFn.apply(x, float(p), tag)


# ---- losses (each returns scalars; gradients w.r.t. logits come from the same kernel launch) ----------------------------
class KDLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z_s, z_t, y, T, alpha, pos_weight, eps=1e-7):
        out, dz = ops.kd_loss(z_s.contiguous().float(), z_t.contiguous().float(), y.contiguous().float(), T, alpha, pos_weight,
                              eps=eps)
        ctx.save_for_backward(dz)
        total, bce, kd = out[0], out[1], out[2]
        ctx.mark_non_differentiable(bce, kd)
        return total, bce, kd

    @staticmethod
    def backward(ctx, g_total, g_bce, g_kd):
        (dz,) = ctx.saved_tensors
        return ops.scale_dev(dz, g_total), None, None, None, None, None, None


class BCELogitsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, y, w_pos, w_neg):
        out, dz = ops.bce_logits(z.contiguous().float(), y.contiguous().float(), w_pos, w_neg)
        ctx.save_for_backward(dz)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return ops.scale_dev(dz, g), None, None, None


class MaskedMseBceFn(torch.autograd.Function):
    """mse(yhat*m, y*m) + w * bce_with_logits(phat, m)  (both means over all elements) -> scalar."""

    @staticmethod
    def forward(ctx, yhat, phat, y, m, w_presence):
        out2 = torch.zeros(2, device=yhat.device, dtype=torch.float32)
        d1, d2 = ops.masked_mse_bce(yhat.contiguous().float(), phat.contiguous().float(), y.contiguous().float(),
                                    m.contiguous().float(), w_presence, out2)
        ctx.save_for_backward(d1, d2)
        return out2[0] + out2[1]

    @staticmethod
    def backward(ctx, g):
        d1, d2 = ctx.saved_tensors
        return ops.scale_dev(d1, g), ops.scale_dev(d2, g), None, None, None


class MaskedBceColsFn(torch.autograd.Function):
    """per-pathology masked BCE [K]; gradient of sum_k coef[k]*per[k] flows to the logits."""

    @staticmethod
    def forward(ctx, z, y, m, pos_weight, eps):
        per, dz = ops.masked_bce_cols(z.contiguous().float(), y.contiguous().float(), m.contiguous().float(), pos_weight,
                                      None, eps)
        ctx.save_for_backward(dz)
        return per

    @staticmethod
    def backward(ctx, gper):
        (dz,) = ctx.saved_tensors
        return ops.scale_dev(dz, gper), None, None, None, None


class AuxResidualKLFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img_logits, scaled_corr, y, mask, eps):
        out, dc = ops.aux_residual_kl(img_logits.detach().contiguous().float(), scaled_corr.contiguous().float(),
                                      y.contiguous().float(), mask.contiguous().float(), eps)
        ctx.save_for_backward(dc)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (dc,) = ctx.saved_tensors
        return None, ops.scale_dev(dc, g), None, None, None


class FusionLogitsFn(torch.autograd.Function):
    """img = hi + bias_i; ts = ht + bias_t; scaled = beta*corr; fusion = img.detach() + scaled."""

    @staticmethod
    def forward(ctx, hi, ht, corr, bias_i, bias_t, beta):
        hi, ht, corr = hi.contiguous().float(), ht.contiguous().float(), corr.contiguous().float()
        img, ts, scaled, fusion = ops.fusion_logits(hi, ht, corr, bias_i.detach(), bias_t.detach(), beta.detach())
        ctx.save_for_backward(corr, beta, bias_i, bias_t)
        return img, ts, scaled, fusion

    @staticmethod
    def backward(ctx, d_img, d_ts, d_scaled, d_fus):
        corr, beta, bias_i, bias_t = ctx.saved_tensors
        sbe, gbe = grad_sink(beta)
        sbi, gbi = grad_sink(bias_i)
        sbt, gbt = grad_sink(bias_t)
        d_corr = ops.fusion_logits_bwd(d_img, d_ts, d_scaled, d_fus, corr, beta.detach(), sbe, sbi, sbt)
        return d_img, d_ts, d_corr, gbi, gbt, gbe
