"""ctypes binding of libduett_b200.so — the C-ABI boundary declared in include/duett_b200.h.

The library is built in-tree by ``__graft_entry__.build()`` (``make -C csrc``).  There is no CPU fallback:
importing this module without the built library raises, and every op raises ``RuntimeError`` when the
library reports a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libduett_b200.so")

DX_F32, DX_BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_RELU, ACT_TANH, ACT_GELU_BWD, ACT_RELU_BWD, ACT_TANH_BWD = range(7)

_lib = None


class DxError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DxError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C multimodal_edema_prediction_b200/csrc`). There is no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.dx_last_error.restype = C.c_char_p
        _declare(_lib)
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise DxError(f"duett_b200 error {rc}: {lib().dx_last_error().decode()}")


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(t: torch.dtype) -> int:
    if t == torch.float32:
        return DX_F32
    if t == torch.bfloat16:
        return DX_BF16
    raise DxError(f"unsupported dtype {t}")


def ptr(t: torch.Tensor | None) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("in_dtype", C.c_int32), ("a_mn", C.c_int32), ("b_mn", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64),
        ("B", C.c_void_p), ("ldb", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("out_dtype", C.c_int32), ("accumulate", C.c_int32),
        ("out2", C.c_void_p), ("ldo2", C.c_int64),
        ("act", C.c_int32), ("act_dtype", C.c_int32),
        ("row_scale", C.c_void_p), ("row_scale2", C.c_void_p), ("bias", C.c_void_p),
        ("res", C.c_void_p), ("ldr", C.c_int64),
        ("aux", C.c_void_p), ("ldx", C.c_int64), ("aux_bias", C.c_void_p),
        ("cx", C.c_void_p), ("ldc", C.c_int64),
        ("coef_num", C.c_void_p), ("coef_den", C.c_void_p),
        ("row_sumsq", C.c_void_p), ("row_dot", C.c_void_p),
        ("force_simt", C.c_int32), ("batch", C.c_int32),
        ("a_bs", C.c_int64), ("b_bs", C.c_int64), ("out_bs", C.c_int64), ("out2_bs", C.c_int64), ("res_bs", C.c_int64),
        ("aux_bs", C.c_int64), ("cx_bs", C.c_int64), ("bias_bs", C.c_int32), ("rowvec_bs", C.c_int32),
        ("split_k", C.c_int32), ("allow_tf32", C.c_int32),
    ]


def _declare(L: C.CDLL) -> None:
    L.dx_version.restype = C.c_int
    L.dx_device_ok.restype = C.c_int
    L.dx_gemm.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
    L.dx_gemm.restype = C.c_int
    L.dx_gemm_tc_debug.argtypes = [C.POINTER(GemmDesc)] + [C.c_int32] * 6 + [C.c_void_p]
    L.dx_gemm_tc_debug.restype = C.c_int
    L.dx_gemm_reserve_sms.argtypes = [C.c_int]
    L.dx_gemm_reserve_sms.restype = C.c_int
    from . import _decl
    _decl.declare(L)


def _ld(t: torch.Tensor) -> int:
    assert t.dim() in (2, 3) and t.stride(-1) == 1, f"need row-major 2-D (or batched 3-D) view, got strides {t.stride()}"
    return t.stride(-2)


def make_gemm_desc(a: torch.Tensor, b: torch.Tensor, *, a_mn=False, b_mn=False, out=None, out2=None,
                   accumulate=False, act=ACT_NONE, act_dtype=None, row_scale=None, row_scale2=None, bias=None,
                   res=None, aux=None, aux_bias=None, cx=None, coef_num=None, coef_den=None, row_sumsq=None,
                   row_dot=None, force_simt=False, split_k=0, tf32=False) -> GemmDesc:
    """a: [M,K] (or [K,M] when a_mn), b: [N,K] (or [K,N] when b_mn); both 2-D with unit inner stride.
    Grouped mode: every matrix argument carries a leading batch dim ([G,M,K], [G,N,K], out [G,M,N], ...; arbitrary batch
    stride), bias is [G,N] and row vectors are [G,M]."""
    batch = a.shape[0] if a.dim() == 3 else 1
    M, K = (a.shape[-1], a.shape[-2]) if a_mn else (a.shape[-2], a.shape[-1])
    N, Kb = (b.shape[-1], b.shape[-2]) if b_mn else (b.shape[-2], b.shape[-1])
    if K != Kb:
        raise DxError(f"dx_gemm: contraction mismatch {K} vs {Kb}")
    if a.dtype != b.dtype:
        raise DxError("dx_gemm: A and B dtypes differ")
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.in_dtype = dtype_code(a.dtype)
    d.a_mn, d.b_mn = int(a_mn), int(b_mn)
    d.A, d.lda = a.data_ptr(), _ld(a)
    d.B, d.ldb = b.data_ptr(), _ld(b)
    d.batch = batch
    if batch > 1:
        assert b.dim() == 3 and b.shape[0] == batch
        d.a_bs, d.b_bs = a.stride(0), b.stride(0)
        d.bias_bs, d.rowvec_bs = N, M
    if out is not None:
        assert tuple(out.shape[-2:]) == (M, N), (out.shape, M, N)
        d.out, d.ldo, d.out_dtype = out.data_ptr(), _ld(out), dtype_code(out.dtype)
        if batch > 1:
            d.out_bs = out.stride(0)
    d.accumulate = int(accumulate)
    ad = act_dtype
    for t in (out2, res, aux, cx):
        if t is not None:
            assert tuple(t.shape[-2:]) == (M, N), (t.shape, M, N)
            ad = t.dtype if ad is None else ad
            assert t.dtype == ad, "out2/res/aux/cx must share the activation dtype"
    d.act = act
    d.act_dtype = dtype_code(ad if ad is not None else a.dtype)
    bs = (lambda t: t.stride(0)) if batch > 1 else (lambda t: 0)
    if out2 is not None:
        d.out2, d.ldo2, d.out2_bs = out2.data_ptr(), _ld(out2), bs(out2)
    if res is not None:
        d.res, d.ldr, d.res_bs = res.data_ptr(), _ld(res), bs(res)
    if aux is not None:
        d.aux, d.ldx, d.aux_bs = aux.data_ptr(), _ld(aux), bs(aux)
    if cx is not None:
        d.cx, d.ldc, d.cx_bs = cx.data_ptr(), _ld(cx), bs(cx)
    for name, t, n in (("row_scale", row_scale, M), ("row_scale2", row_scale2, M), ("bias", bias, N),
                       ("aux_bias", aux_bias, N), ("coef_num", coef_num, M), ("coef_den", coef_den, M),
                       ("row_sumsq", row_sumsq, M), ("row_dot", row_dot, M)):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n * batch, (name, t.shape, t.dtype, n)
            setattr(d, name, t.data_ptr())
    d.force_simt = int(force_simt)
    d.split_k = int(split_k)
    d.allow_tf32 = int(bool(tf32))
    return d


def gemm(a: torch.Tensor, b: torch.Tensor, **kw) -> None:
    d = make_gemm_desc(a, b, **kw)
    check(lib().dx_gemm(C.byref(d), stream_ptr()))
