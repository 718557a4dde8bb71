"""Thin Python wrappers over the C ABI (include/duett_b200.h).  Every function launches hand-written sm_100a
kernels on the current CUDA stream; tensors are only used for their device memory.  No CPU fallback exists: a
missing library or a non-zero status raises."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from ._lib import ACT_GELU, ACT_GELU_BWD, ACT_NONE, ACT_RELU, ACT_RELU_BWD, ACT_TANH, ACT_TANH_BWD, gemm  # noqa: F401

_launches = 0   # kernels-launching C calls issued (bench.py reports it as gpu_launches)


def gemm_reserve_sms(n: int) -> int:
    """Persistent GEMM grids leave n SMs free (for the NCCL kernels of an overlapped all-reduce); returns the previous n."""
    return int(L.lib().dx_gemm_reserve_sms(int(n)))


def launches() -> int:
    return _launches


def _p(t):
    return None if t is None else t.data_ptr()


def _call(name, *args):
    global _launches
    _launches += 1
    L.check(getattr(L.lib(), name)(*args, torch.cuda.current_stream().cuda_stream))


def _dt(t: torch.Tensor) -> int:
    return L.dtype_code(t.dtype)


def _chk(t, dtype=None, contiguous=True):
    assert t.is_cuda, "duett_b200 ops need CUDA tensors (there is no CPU path)"
    if contiguous:
        assert t.is_contiguous(), "tensor must be contiguous"
    if dtype is not None:
        assert t.dtype == dtype, (t.dtype, dtype)
    return t


PROFILE = None   # bench.py sets this to a list: every GEMM / attention launch then records (kind, tag, flops, bytes, start_evt, end_evt)


def _prof_begin():
    if PROFILE is None:
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    return e0


def _prof_end(e0, kind, tag, flops, nbytes):
    if e0 is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    PROFILE.append((kind, tag, flops, nbytes, e0, e1))


TF32 = False     # fp32-operand GEMMs on tcgen05 kind::tf32 instead of exact FFMA (set through tf32_mode)


class tf32_mode:
    """with ops.tf32_mode(True): every fp32 dx_gemm issued inside may use the tensor cores (kind::tf32) — the "tf32" precision
    of the modules = the reference's own SSL / fine-tune precision (torch.set_float32_matmul_precision('high'),
    duett/duett.py:9).  The exact FFMA mode ("fp32") stays the 1e-3 parity mode."""

    def __init__(self, on):
        self.on = bool(on)

    def __enter__(self):
        global TF32
        self.prev, TF32 = TF32, self.on
        return self

    def __exit__(self, *a):
        global TF32
        TF32 = self.prev


def gemm_(a, b, **kw):
    global _launches
    _launches += 1
    if TF32 and a.dtype == torch.float32 and "tf32" not in kw:
        kw["tf32"] = True
    if PROFILE is None:
        return gemm(a, b, **kw)
    a_mn, b_mn = kw.get("a_mn", False), kw.get("b_mn", False)
    G = a.shape[0] if a.dim() == 3 else 1
    M, K = (a.shape[-1], a.shape[-2]) if a_mn else (a.shape[-2], a.shape[-1])
    N = b.shape[-1] if b_mn else b.shape[-2]
    tc = (a.dtype == torch.bfloat16 or kw.get("tf32", False)) and not kw.get("force_simt", False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gemm(a, b, **kw)
    e1.record()
    esz = a.element_size()
    nbytes = G * (M * K + N * K) * esz
    for k in ("out", "out2", "res", "aux", "cx"):
        t = kw.get(k)
        if t is not None:
            nbytes += t.numel() * t.element_size() * (2 if (k == "out" and kw.get("accumulate")) else 1)
    PROFILE.append((("tc" if tc else "ffma"), f"{G}*" * (G > 1) + f"{M}x{N}x{K}{'a' if a_mn else ''}{'b' if b_mn else ''}",
                    2.0 * G * M * N * K, nbytes, e0, e1))


# ---- relayout / norms -------------------------------------------------------------------------------------------------
def relayout_fwd(src, B, P, Q, d, *, src_rowsq=None, g=None, pos_bcast=None, pos_batched=None, want_rowsq=True):
    """dst[b,q,p,:] = src[b,p,q,:] * final_norm_scale[b,p] + pos ; returns (dst [B,Q,P,d], rowsq [B*Q] or None)."""
    _chk(src)
    dst = torch.empty((B, Q, P, d), device=src.device, dtype=src.dtype)
    rowsq = torch.empty(B * Q, device=src.device, dtype=torch.float32) if want_rowsq else None
    if pos_batched is not None:
        assert pos_batched.dtype == src.dtype and pos_batched.is_contiguous()
    _call("dx_relayout_fwd", _p(src), _p(src_rowsq), _p(g), _p(pos_bcast), _p(pos_batched), _p(dst), _p(rowsq), B, P, Q, d,
          _dt(src))
    return dst, rowsq


def relayout_bwd(gdst, B, P, Q, d, *, src=None, src_rowsq=None, g=None, dg=None):
    """grad wrt src [B,P,Q,d] given grad wrt dst [B,Q,P,d]; applies the final-ScaleNorm backward when src_rowsq is given."""
    _chk(gdst)
    dsrc = torch.empty((B, P, Q, d), device=gdst.device, dtype=gdst.dtype)
    _call("dx_relayout_bwd", _p(gdst), _p(src), _p(src_rowsq), _p(g), _p(dsrc), _p(dg), B, P, Q, d, _dt(gdst))
    return dsrc


def colsum(x2d, out, accumulate=True):
    """out[n] (+)= sum_m x2d[m, n]; x2d may be a row-strided 2-D view."""
    assert x2d.dim() == 2 and x2d.stride(1) == 1
    _chk(out, torch.float32)
    _call("dx_colsum", _p(x2d), x2d.stride(0), x2d.shape[0], x2d.shape[1], _p(out), int(accumulate), _dt(x2d))


def axpy(x, y, alpha=1.0, accumulate=True):
    _chk(x); _chk(y)
    assert x.numel() == y.numel() and x.dtype == y.dtype
    _call("dx_axpy", _p(x), _p(y), x.numel(), float(alpha), int(accumulate), _dt(x))


def sum_n(tensors):
    """Sum of 1..8 same-shape tensors (fp32 accumulation, one rounding) -> new tensor."""
    import ctypes
    t0 = tensors[0]
    for t in tensors:
        _chk(t)
        assert t.shape == t0.shape and t.dtype == t0.dtype
    y = torch.empty_like(t0)
    arr = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    _call("dx_sum_n", arr, len(tensors), _p(y), t0.numel(), _dt(t0))
    return y


def axpy_f32(x, y, alpha):
    """y += alpha * x for small f32 vectors of any length (dx_axpy needs n % 8 == 0, so go through dx_scale/dx_colsum)."""
    _chk(x, torch.float32); _chk(y, torch.float32)
    n = x.numel()
    if n % 8 == 0:
        return axpy(x, y, alpha, accumulate=True)
    xs = x.reshape(1, n) if alpha == 1.0 else (x * alpha).reshape(1, n)
    colsum(xs, y, accumulate=True)


def cast(x, dtype):
    _chk(x)
    if x.dtype == dtype:
        return x
    y = torch.empty(x.shape, device=x.device, dtype=dtype)
    _call("dx_cast", _p(x), _dt(x), _p(y), _dt(y), x.numel())
    return y


def cast_into(x, y):
    """y[:] = x converted to y's dtype (same numel)."""
    _chk(x); _chk(y)
    assert x.numel() == y.numel()
    _call("dx_cast", _p(x), _dt(x), _p(y), _dt(y), x.numel())
    return y


def scalenorm_scale(rowsq, g, dim):
    out = torch.empty_like(rowsq)
    _call("dx_scalenorm_scale", _p(rowsq), _p(g), float(dim) ** 0.5, _p(out), rowsq.numel())
    return out


def rowdot_scale(a, g, row_scale):
    """returns rowdot[n] = <a[n], g[n]>; scales g[n] by row_scale[n] in place."""
    _chk(a); _chk(g)
    N, Cc = a.shape
    rd = torch.empty(N, device=a.device, dtype=torch.float32)
    _call("dx_rowdot_scale", _p(a), _p(g), _p(row_scale), _p(rd), N, Cc, _dt(a))
    return rd


# ---- attention ---------------------------------------------------------------------------------------------------------
def _view3(t):
    """(ptr tensor, batch stride, row stride) of a [B,S,*] tensor whose last dim is dense."""
    assert t.dim() == 3 and t.stride(2) == 1
    return t, t.stride(0), t.stride(1)


# ---- dropout seeds --------------------------------------------------------------------------------------------------
# Every dropout site (attention probabilities, FFN hidden, nn.Dropout modules) draws a unique constant seed when it runs;
# the kernels add the device-resident step counter to it, so a captured CUDA graph draws fresh masks on every replay once
# the training loop calls advance_drop_step() (captured with the step).  The forward call's seed is kept by the autograd
# node and handed to the backward call, which regenerates the same mask.  Determinism follows torch.manual_seed.
_DROP = {"base": None, "site": 0, "dev": None}
DROP_LOG = None      # tests set this to a list: every site appends (tag, p, seed, shape)


def drop_seed(tag="", p=0.0, shape=None):
    if _DROP["base"] is None:
        rank = 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            rank = torch.distributed.get_rank()      # data-parallel replicas draw different masks (each sees its own shard)
        _DROP["base"] = (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019
                         + rank * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF
    _DROP["site"] += 1
    seed = (_DROP["base"] + _DROP["site"] * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    if DROP_LOG is not None:
        DROP_LOG.append((tag, float(p), seed, tuple(shape) if shape is not None else None))
    return seed


def reset_drop_seeds(seed=None):
    """Restart the site sequence (tests; torch.manual_seed alone does not rewind it)."""
    _DROP["base"], _DROP["site"] = (None if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF), 0
    if _DROP["dev"] is not None:
        _DROP["dev"].zero_()


def drop_step_counter(device):
    """Device-resident uint64 (stored as int64) added to every site seed."""
    if _DROP["dev"] is None or _DROP["dev"].device != torch.device(device):
        _DROP["dev"] = torch.zeros(1, device=device, dtype=torch.int64)
    return _DROP["dev"]


def advance_drop_step():
    """Call once per optimisation step, outside forward..backward (capturable)."""
    if _DROP["dev"] is not None:
        _DROP["dev"].add_(1)


def dropout(x, p, seed, out=None):
    """out = x * keep / (1-p); calling it on the upstream gradient with the same (p, seed) is the backward pass."""
    _chk(x)
    y = x if out is x else (torch.empty_like(x) if out is None else out)
    _call("dx_dropout", _p(x), _p(y), x.numel(), float(p), int(seed), _p(drop_step_counter(x.device)), _dt(x))
    return y


def rowdot_bias(a, b, bias):
    """out[n] = sum_c a[n,c] * (b[n,c] - bias[c])."""
    _chk(a); _chk(b)
    N, C = a.shape
    out = torch.empty(N, device=a.device, dtype=torch.float32)
    _call("dx_rowdot_bias", _p(a), _p(b), _p(bias), _p(out), N, C, _dt(a))
    return out


def attn_fwd(q, k, v, heads, drop=None):
    """q [B,Sq,h*dh], k/v [B,Sk,h*dh] (strided views allowed) -> o [B,Sq,h*dh], lse [B,h,Sq].
    drop = (p, seed): dropout on the attention probabilities."""
    B, Sq, D = q.shape
    Sk = k.shape[1]
    dh = D // heads
    o = torch.empty((B, Sq, D), device=q.device, dtype=q.dtype)
    lse = torch.empty((B, heads, Sq), device=q.device, dtype=torch.float32)
    args = []
    for t in (q, k, v, o):
        t_, bs, rs = _view3(t)
        args += [_p(t_), bs, rs]
    dp, dseed = (float(drop[0]), int(drop[1])) if drop else (0.0, 0)
    ev = _prof_begin()
    _call("dx_attn_fwd", *args, _p(lse), B, heads, Sq, Sk, dh, _dt(q), dp, dseed,
          _p(drop_step_counter(q.device)) if drop else None)
    _prof_end(ev, "attn_fwd", f"B{B}xH{heads}xS{Sq}x{Sk}xdh{dh}", 4.0 * B * heads * Sq * Sk * dh,
              (q.numel() + k.numel() + v.numel() + o.numel()) * q.element_size())
    return o, lse


def attn_probs_mean(q, k, lse, heads):
    """Head-averaged attention probabilities [B,Sq,Sk] f32 from q [B,Sq,h*dh], k [B,Sk,h*dh] and the lse [B,h,Sq] of a
    dropout-free attn_fwd call: nn.MultiheadAttention(need_weights=True, average_attn_weights=True)'s second output."""
    B, Sq, D = q.shape
    Sk = k.shape[1]
    for t in (q, k, lse):
        require_device(t)
    assert k.dtype == q.dtype and lse.dtype == torch.float32 and lse.is_contiguous() and lse.shape == (B, heads, Sq)
    out = torch.empty((B, Sq, Sk), device=q.device, dtype=torch.float32)
    q_, qbs, qrs = _view3(q)
    k_, kbs, krs = _view3(k)
    _call("dx_attn_probs_mean", _p(q_), qbs, qrs, _p(k_), kbs, krs, _p(lse), _p(out), B, heads, Sq, Sk, D // heads, _dt(q))
    return out


def attn_bwd(q, k, v, o, go, lse, heads, dq, dk, dv, drop=None):
    B, Sq, D = q.shape
    Sk = k.shape[1]
    dh = D // heads
    Dws = torch.empty((B, heads, Sq), device=q.device, dtype=torch.float32)
    args = []
    for t in (q, k, v, o, go, dq, dk, dv):
        t_, bs, rs = _view3(t)
        args += [_p(t_), bs, rs]
    dp, dseed = (float(drop[0]), int(drop[1])) if drop else (0.0, 0)
    ev = _prof_begin()
    _call("dx_attn_bwd", *args, _p(lse), _p(Dws), B, heads, Sq, Sk, dh, _dt(q), dp, dseed,
          _p(drop_step_counter(q.device)) if drop else None)
    _prof_end(ev, "attn_bwd", f"B{B}xH{heads}xS{Sq}x{Sk}xdh{dh}", 10.0 * B * heads * Sq * Sk * dh,
              (2 * q.numel() + 2 * k.numel() + 2 * v.numel() + 2 * o.numel()) * q.element_size())


# ---- embedding ---------------------------------------------------------------------------------------------------------
def _embed_hidden(xs, V, W0, b0, nobs, gamma, beta, mean, rstd, act_dtype):
    B, T, _ = xs.shape
    hn = torch.empty((V, B * (T + 1), 64), device=xs.device, dtype=act_dtype)
    _call("dx_embed_hidden", _p(xs), B, T, V, _p(W0), _p(b0), _p(nobs), _p(gamma), _p(beta), _p(mean), _p(rstd), _p(hn),
          L.dtype_code(act_dtype))
    return hn


def _psi_groups(psi, B, T1, V, d):
    """[V, B*T1, d] strided view of psi[B,T1,V+1,d]: group v = variable v's column (row stride (V+1)*d, group stride d)."""
    return psi.view(B * T1, V + 1, d)[:, :V, :].permute(1, 0, 2)


def embed_fwd(xs, V, d, W0, b0, gamma, beta, run_mean, run_var, W4, b4, nobs, special, tab, act_dtype, training,
              return_hidden=False):
    """psi[B,T+1,V+1,d] = embedding of the binned grid: stats -> hidden -> ONE grouped tensor-core GEMM over the V
    variables (64 -> d, written into the strided psi view) -> special-cell substitution."""
    B, T, _ = xs.shape
    dev = xs.device
    for t in (xs, W0, b0, gamma, beta, W4, b4, nobs, special, tab):
        _chk(t, torch.float32)
    stats = torch.empty((V, 64, 2), device=dev, dtype=torch.float64) if training else None
    mean = torch.empty((V, 64), device=dev, dtype=torch.float32)
    rstd = torch.empty((V, 64), device=dev, dtype=torch.float32)
    _call("dx_embed_stats", _p(xs), B, T, V, _p(W0), _p(b0), _p(nobs), _p(run_mean), _p(run_var), _p(stats), _p(mean), _p(rstd),
          int(training))
    hn = _embed_hidden(xs, V, W0, b0, nobs, gamma, beta, mean, rstd, act_dtype)
    psi = torch.empty((B, T + 1, V + 1, d), device=dev, dtype=act_dtype)
    gemm_(hn, cast(W4, act_dtype), out=_psi_groups(psi, B, T + 1, V, d), bias=b4, act_dtype=act_dtype)
    _call("dx_embed_special", _p(xs), B, T, V, d, _p(special), _p(tab), _p(psi), L.dtype_code(act_dtype))
    if return_hidden:            # the backward's dW4 GEMM takes hn as an operand: keeping it saves one recomputation pass
        return psi, mean, rstd, hn
    return psi, mean, rstd


def embed_bwd(xs, V, d, W0, b0, gamma, beta, W4, nobs, mean, rstd, dpsi, grads, training, hn=None):
    """grads: dict with f32 tensors dW0, db0, dgamma, dbeta, dW4, db4, dnobs, dspecial (accumulated). Returns dtab [B,d].
    dpsi is consumed (its special cells are zeroed in place)."""
    B, T, _ = xs.shape
    T1 = T + 1
    dev, at = xs.device, dpsi.dtype
    _chk(dpsi)
    dtab = torch.empty((B, d), device=dev, dtype=torch.float32)
    _call("dx_embed_special_bwd", _p(xs), B, T, V, d, _p(dpsi), _dt(dpsi), _p(grads["dspecial"]), _p(dtab))
    if hn is None:
        hn = _embed_hidden(xs, V, W0, b0, nobs, gamma, beta, mean, rstd, at)
    dv = _psi_groups(dpsi, B, T1, V, d)                                           # [V, B*T1, d]
    gemm_(dv, hn, a_mn=True, b_mn=True, out=grads["dW4"], accumulate=True)         # dW4[v] += dout_v^T hn_v
    colsum(dpsi.view(B * T1, (V + 1) * d)[:, :V * d], grads["db4"].view(-1), accumulate=True)
    dhn = torch.empty((V, B * T1, 64), device=dev, dtype=at)
    gemm_(dv, cast(W4, at), b_mn=True, out=dhn, act_dtype=at)                      # dhn_v = dout_v W4_v
    dgb = torch.empty((2, V, 64), device=dev, dtype=torch.float32)
    _call("dx_embed_bn_reduce", _p(xs), B, T, V, _p(W0), _p(b0), _p(nobs), _p(mean), _p(rstd), _p(dhn), _dt(dhn), _p(dgb))
    _call("dx_embed_bwd_front", _p(xs), B, T, V, _p(W0), _p(b0), _p(nobs), _p(gamma), _p(mean), _p(rstd), _p(dhn), _dt(dhn),
          _p(dgb), _p(grads["dW0"]), _p(grads["db0"]), _p(grads["dnobs"]), int(training))
    axpy(dgb[0], grads["dgamma"], 1.0, accumulate=True)
    axpy(dgb[1], grads["dbeta"], 1.0, accumulate=True)
    return dtab


# ---- norms -------------------------------------------------------------------------------------------------------------
def bn2d_fwd(x, w, b, run_mean, run_var, training):
    _chk(x, torch.float32)
    R, Cc = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(Cc, device=x.device, dtype=torch.float32)
    rstd = torch.empty(Cc, device=x.device, dtype=torch.float32)
    ws = torch.empty((Cc, 2), device=x.device, dtype=torch.float64) if training else None
    _call("dx_bn2d_fwd", _p(x), R, Cc, _p(w), _p(b), _p(run_mean), _p(run_var), _p(y), _p(mean), _p(rstd), _p(ws), int(training))
    return y, mean, rstd


def bn2d_bwd(dy, x, w, mean, rstd, dw, db, training, need_dx=True):
    _chk(dy, torch.float32)
    R, Cc = x.shape
    dx = torch.empty_like(x) if need_dx else None
    ws = torch.empty((Cc, 2), device=x.device, dtype=torch.float32)
    _call("dx_bn2d_bwd", _p(dy), _p(x), R, Cc, _p(w), _p(mean), _p(rstd), _p(dx), _p(dw), _p(db), _p(ws), int(training))
    return dx


def layernorm_fwd(x2d, w, b):
    _chk(x2d)
    R, Cc = x2d.shape
    y = torch.empty_like(x2d)
    mean = torch.empty(R, device=x2d.device, dtype=torch.float32)
    rstd = torch.empty(R, device=x2d.device, dtype=torch.float32)
    _call("dx_layernorm_fwd", _p(x2d), R, Cc, _p(w), _p(b), _p(y), _p(mean), _p(rstd), _dt(x2d))
    return y, mean, rstd


def layernorm_bwd(dy, x2d, w, mean, rstd, dw, db, need_dx=True):
    _chk(dy); _chk(x2d)
    R, Cc = x2d.shape
    dx = torch.empty_like(x2d) if need_dx else None
    _call("dx_layernorm_bwd", _p(dy), _p(x2d), R, Cc, _p(w), _p(mean), _p(rstd), _p(dx), _p(dw), _p(db), _dt(x2d))
    return dx


# ---- pooling / gathers ---------------------------------------------------------------------------------------------------
def mean_rows(x3d, T):
    _chk(x3d)
    B, T1, E = x3d.shape
    y = torch.empty((B, E), device=x3d.device, dtype=torch.float32)
    _call("dx_mean_rows", _p(x3d), _p(y), B, T1, T, E, _dt(x3d))
    return y


def mean_rows_bwd(dy, T1, T, dtype):
    _chk(dy, torch.float32)
    B, E = dy.shape
    dx = torch.empty((B, T1, E), device=dy.device, dtype=dtype)
    _call("dx_mean_rows_bwd", _p(dy), _p(dx), B, T1, T, E, L.dtype_code(dtype))
    return dx


def gather_vec(src, offsets, Lv):
    _chk(src); _chk(offsets, torch.int64)
    n = offsets.numel()
    out = torch.empty((n, Lv), device=src.device, dtype=torch.float32)
    _call("dx_gather_vec", _p(src), _p(offsets), _p(out), n, Lv, _dt(src))
    return out


def scatter_vec(src, offsets, dst, accumulate=True):
    _chk(src, torch.float32); _chk(offsets, torch.int64); _chk(dst)
    n, Lv = src.shape
    _call("dx_scatter_vec", _p(src), _p(offsets), _p(dst), n, Lv, int(accumulate), _dt(dst))


# ---- losses --------------------------------------------------------------------------------------------------------------
def kd_loss(zs, zt, y, T, alpha, pos_weight, need_grad=True, eps=1e-7):
    for t in (zs, zt, y):
        _chk(t, torch.float32)
    out = torch.empty(3, device=zs.device, dtype=torch.float32)
    dz = torch.empty_like(zs) if need_grad else None
    _call("dx_kd_loss", _p(zs), _p(zt), _p(y), zs.numel(), float(T), float(alpha),
          -1.0 if pos_weight is None else float(pos_weight), float(eps), _p(out), _p(dz))
    return out, dz


def bce_logits(z, y, w_pos=1.0, w_neg=1.0, need_grad=True):
    _chk(z, torch.float32); _chk(y, torch.float32)
    out = torch.empty(1, device=z.device, dtype=torch.float32)
    dz = torch.empty_like(z) if need_grad else None
    _call("dx_bce_logits", _p(z), _p(y), z.numel(), float(w_pos), float(w_neg), _p(out), _p(dz))
    return out, dz


def masked_mse_bce(yhat, phat, y, m, w_presence, out2, need_grad=True):
    for t in (yhat, phat, y, m):
        _chk(t, torch.float32)
    d1 = torch.empty_like(yhat) if need_grad else None
    d2 = torch.empty_like(phat) if need_grad else None
    _call("dx_masked_mse_bce", _p(yhat), _p(phat), _p(y), _p(m), yhat.numel(), float(w_presence), _p(out2), _p(d1), _p(d2))
    return d1, d2


def masked_bce_cols(z, y, m, pos_weight, coef, eps, need_grad=True):
    for t in (z, y, m):
        _chk(t, torch.float32)
    B, K = z.shape
    per = torch.empty(K, device=z.device, dtype=torch.float32)
    dz = torch.empty_like(z) if need_grad else None
    _call("dx_masked_bce_cols", _p(z), _p(y), _p(m), _p(pos_weight), _p(coef), B, K, float(eps), _p(per), _p(dz))
    return per, dz


def aux_residual_kl(img_logits, scaled_corr, y, mask, eps=0.05, need_grad=True):
    for t in (img_logits, scaled_corr, y, mask):
        _chk(t, torch.float32)
    out = torch.empty(1, device=y.device, dtype=torch.float32)
    dc = torch.empty_like(scaled_corr) if need_grad else None
    _call("dx_aux_residual_kl", _p(img_logits), _p(scaled_corr), _p(y), _p(mask), y.numel(), float(eps), _p(out), _p(dc))
    return out, dc


def require_device(t):
    if not t.is_cuda:
        raise L.DxError("duett_b200 has no CPU path: inputs must be CUDA tensors on a B200")


def act_fwd(x, act):
    """act(x) for ACT_GELU / ACT_RELU / ACT_TANH (after a split-K GEMM, whose epilogue cannot apply the activation)."""
    _chk(x)
    out = torch.empty_like(x)
    _call("dx_act_fwd", _p(x), _p(out), x.numel(), act, _dt(x))
    return out


def act_bwd(g, aux, act_bwd_code):
    """g * act'(aux) (RELU_BWD / TANH_BWD: aux = activation output; GELU_BWD: aux = pre-activation)."""
    _chk(g); _chk(aux)
    out = torch.empty_like(g)
    _call("dx_act_bwd", _p(g), _p(aux), _p(out), g.numel(), act_bwd_code, _dt(g))
    return out


def scale_dev(x, s):
    """x * s with s a device tensor: 1 element (upstream loss gradient, no host sync) or [K] (per-column scale)."""
    _chk(x, torch.float32)
    out = torch.empty_like(x)
    s = s.reshape(-1).float().contiguous()
    _call("dx_scale_dev", _p(x), _p(s), _p(out), x.numel(), s.numel())
    return out


def sum_div_acc(x, g, sink):
    """sink[0] += sum(x) / g[0]."""
    _call("dx_sum_div_acc", _p(x), x.numel(), _p(g), _p(sink))


def fusion_logits(hi, ht, corr, bias_i, bias_t, beta):
    B, K = hi.shape
    img, ts, scaled, fusion = (torch.empty_like(hi) for _ in range(4))
    _call("dx_fusion_logits", _p(hi), _p(ht), _p(corr), _p(bias_i), _p(bias_t), _p(beta), _p(img), _p(ts), _p(scaled),
          _p(fusion), B, K)
    return img, ts, scaled, fusion


def fusion_logits_bwd(d_img, d_ts, d_scaled, d_fus, corr, beta, dbeta, dbias_i, dbias_t):
    B, K = corr.shape
    d_corr = torch.empty_like(corr)
    c = lambda t: None if t is None else t.contiguous()
    d_img, d_ts, d_scaled, d_fus = c(d_img), c(d_ts), c(d_scaled), c(d_fus)
    _call("dx_fusion_logits_bwd", _p(d_img), _p(d_ts), _p(d_scaled), _p(d_fus), _p(corr), _p(beta), _p(d_corr), _p(dbeta),
          _p(dbias_i), _p(dbias_t), B, K)
    return d_corr


# ---- optimizer -------------------------------------------------------------------------------------------------------------
def adamw(p, g, m, v, lr, betas, eps, weight_decay, step, grad_scale_dev=None, grad_scale=1.0, step_dev=None,
          lr_scale_dev=None, shadow=None):
    for t in (p, g, m, v):
        _chk(t, torch.float32)
    if shadow is not None:
        _chk(shadow, torch.bfloat16)
        assert shadow.numel() == p.numel()
    _call("dx_adamw", _p(p), _p(g), _p(m), _p(v), p.numel(), float(lr), float(betas[0]), float(betas[1]), float(eps),
          float(weight_decay), int(step), _p(grad_scale_dev), float(grad_scale), _p(step_dev), _p(lr_scale_dev), _p(shadow))


def sumsq(x, out):
    _chk(x, torch.float32)
    _call("dx_sumsq", _p(x), x.numel(), _p(out))


def bin_events(slot, vals, cnts, row_start, means, stds, T):
    """Dense [B,T,2V] f32 grid from the concatenated event rows of B stays (duett/mimic_dataset.py:33-46 semantics).
    slot [R] int32, vals / cnts [R,V] f64, row_start [B+1] int64, means / stds [V] f64 - all on the device."""
    for t, dt in ((slot, torch.int32), (vals, torch.float64), (cnts, torch.float64), (row_start, torch.int64),
                  (means, torch.float64), (stds, torch.float64)):
        _chk(t, dt)
    B, V = row_start.numel() - 1, means.numel()
    x = torch.empty((B, T, 2 * V), device=vals.device, dtype=torch.float32)
    _call("dx_bin_events", _p(slot), _p(vals), _p(cnts), _p(row_start), _p(means), _p(stds), B, T, V, _p(x))
    return x


def ssl_mask(xs, step, ev, keep):
    """Model.pretrain_prep_batch's masking in one launch (duett/duett.py:189-237).  xs [B,T,2V+1] f32, step [B] or [B,K]
    int32 (K = pretrain_masked_steps draws per sample, repeats allowed), ev [B] int32 (None: no event prediction), keep
    [B,V] uint8 or None -> (x_c, y_ts, y_ts_masks, y_events, y_events_mask); y_ts / y_ts_masks are [B,V] for a 1-D step and
    [B,K,V] for a 2-D one."""
    _chk(xs, torch.float32); _chk(step, torch.int32)
    B, T, C = xs.shape
    V = (C - 1) // 2
    K = 1 if step.dim() == 1 else step.shape[1]
    assert step.shape[0] == B and step.dim() in (1, 2)
    xc = torch.empty_like(xs)
    tshape = (B, V) if step.dim() == 1 else (B, K, V)
    y_ts = torch.empty(tshape, device=xs.device, dtype=torch.float32)
    y_mask = torch.empty(tshape, device=xs.device, dtype=torch.float32)
    y_ev = y_ev_mask = None
    if ev is not None:
        _chk(ev, torch.int32)
        y_ev = torch.empty((B, T), device=xs.device, dtype=torch.float32)
        y_ev_mask = torch.empty((B, T), device=xs.device, dtype=torch.float32)
    if keep is not None:
        _chk(keep, torch.uint8)
        assert keep.shape == (B, V)
    _call("dx_ssl_mask", _p(xs), _p(step), _p(ev), _p(keep), B, T, V, K, _p(xc), _p(y_ts), _p(y_mask), _p(y_ev), _p(y_ev_mask))
    return xc, y_ts, y_mask, y_ev, y_ev_mask


def binary_auc(logits, labels, apply_sigmoid=True):
    """AUROC / AUPRC of a binary scorer on the device -> float64 tensor [auroc, auprc, n_pos, n] (training_duett/
    evaluator.py:22-35 semantics: sklearn roc_auc_score / average_precision_score of sigmoid(logits))."""
    logits = logits.reshape(-1).float().contiguous()
    labels = labels.reshape(-1).float().contiguous()
    _chk(logits, torch.float32)
    n = logits.numel()
    if labels.numel() != n or n == 0:
        raise L.DxError(f"binary_auc: {n} logits vs {labels.numel()} labels")
    npad = 1 << max(1, (n - 1).bit_length())
    ws = torch.empty((2, npad), device=logits.device, dtype=torch.float32)
    out = torch.empty(4, device=logits.device, dtype=torch.float64)
    _call("dx_binary_auc", _p(logits), _p(labels), n, _p(ws[0]), _p(ws[1]), npad, int(bool(apply_sigmoid)), _p(out))
    return out


def clip_factor(sumsq_t, max_norm, clip):
    _call("dx_clip_factor", _p(sumsq_t), float(max_norm), _p(clip))
