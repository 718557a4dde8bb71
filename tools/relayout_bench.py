"""Times dx_relayout_fwd / dx_relayout_bwd at the bench shapes (CUDA events, L2 flushed between launches).
usage: python tools/relayout_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodal_edema_prediction_b200 import ops

B, T1, V1, d = 256, 33, 129, 128
bf = torch.bfloat16
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def t(fn, n=10):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3


for name, P, Q in (("time->event", T1, V1), ("event->time", V1, T1)):
    src = torch.randn(B, P, Q, d, device="cuda").to(bf)
    rowsq = torch.rand(B * P, device="cuda") * 100 + 1
    g = torch.ones(1, device="cuda")
    pos_b = torch.randn(Q, P * d, device="cuda")
    pos_n = torch.randn(B, Q, P * d, device="cuda").to(bf)
    nb = src.numel() * 2
    f1 = t(lambda: ops.relayout_fwd(src, B, P, Q, d, src_rowsq=rowsq, g=g, pos_bcast=pos_b))
    f2 = t(lambda: ops.relayout_fwd(src, B, P, Q, d, src_rowsq=rowsq, g=g, pos_batched=pos_n))
    gd = torch.randn(B, Q, P, d, device="cuda").to(bf)
    dg = torch.zeros(1, device="cuda")
    b1 = t(lambda: ops.relayout_bwd(gd, B, P, Q, d, src=src, src_rowsq=rowsq, g=g, dg=dg))
    b2 = t(lambda: ops.relayout_bwd(gd, B, P, Q, d))
    print(f"{name}: fwd+pos_bcast {f1:.0f} us ({2 * nb / f1 / 1e6:.2f} TB/s)  fwd+pos_batched {f2:.0f} us ({3 * nb / f2 / 1e6:.2f} TB/s)  "
          f"bwd+norm {b1:.0f} us ({3 * nb / b1 / 1e6:.2f} TB/s)  bwd plain {b2:.0f} us ({2 * nb / b2 / 1e6:.2f} TB/s)")
