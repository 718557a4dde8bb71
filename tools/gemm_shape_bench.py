"""Times one dx_gemm shape/epilogue with CUDA events (and is the command profiled under ncu).
usage: python tools/gemm_shape_bench.py M N K [epi] [iters]   epi in plain|bias|res_rowsq|dx|resid|ffn_in|gelu_bwd|dw"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodal_edema_prediction_b200 import ops

M, N, K = (int(x) for x in sys.argv[1:4])
epi = sys.argv[4] if len(sys.argv) > 4 else "plain"
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 20
bf = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
kw, nb = {}, 0
if epi == "dw":      # A^T B with both operands MN-major, fp32 accumulate (M,N small; K = tokens)
    a, b = rn(K, M).to(bf), rn(K, N).to(bf)
    out = torch.zeros(M, N, device="cuda")
    kw = dict(a_mn=True, b_mn=True, out=out, accumulate=True)
    nbytes = (K * M + K * N) * 2 + M * N * 8
else:
    a = rn(M, K).to(bf)
    b = rn(K, N).to(bf) if epi in ("dx", "gelu_bwd") else rn(N, K).to(bf)
    out = torch.empty(M, N, device="cuda", dtype=bf)
    nbytes = (M * K + N * K + M * N) * 2
    kw = dict(out=out, act_dtype=bf, b_mn=epi in ("dx", "gelu_bwd"))
    if epi == "dx":
        kw.update(res=rn(M, N).to(bf), cx=rn(M, N).to(bf), coef_num=rn(M), coef_den=torch.rand(M, device="cuda") + 1)
        nbytes += 2 * M * N * 2
    elif epi == "resid":
        kw.update(res=rn(M, N).to(bf), bias=rn(N), row_sumsq=torch.zeros(M, device="cuda"))
        nbytes += M * N * 2
    elif epi == "res_rowsq":
        kw.update(res=rn(M, N).to(bf), row_sumsq=torch.zeros(M, device="cuda"))
        nbytes += M * N * 2
    elif epi == "bias":
        kw.update(bias=rn(N))
    elif epi == "ffn_in":
        kw.update(row_scale=torch.rand(M, device="cuda"), bias=rn(N), act=ops.ACT_GELU, out2=torch.empty(M, N, device="cuda", dtype=bf))
        nbytes += M * N * 2
    elif epi == "gelu_bwd":
        kw.update(act=ops.ACT_GELU_BWD, aux=rn(M, N).to(bf), aux_bias=rn(N), row_scale2=torch.rand(M, device="cuda"),
                  out2=torch.empty(M, N, device="cuda", dtype=bf), row_dot=torch.zeros(M, device="cuda"))
        nbytes += 2 * M * N * 2
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for _ in range(3):
    ops.gemm_(a, b, **kw)
torch.cuda.synchronize()
ts = []
for _ in range(iters):
    flush.zero_()                                   # evict L2 between timed launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm_(a, b, **kw); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
ms = ts[len(ts) // 2]
print(f"{M}x{N}x{K} {epi}: {ms:.4f} ms  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s  {nbytes/ms/1e6:.0f} GB/s (algorithmic)")
