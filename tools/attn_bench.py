"""Times dx_attn_fwd / dx_attn_bwd at the bench shapes (CUDA events).  usage: python tools/attn_bench.py [B S D H]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodal_edema_prediction_b200 import ops

shapes = [(256, 129, 128, 2), (256, 33, 128, 2)] if len(sys.argv) < 5 else [tuple(int(x) for x in sys.argv[1:5])]
for B, S, D, H in shapes:
    qkv = (torch.randn(B, S, 3 * D, device="cuda") * 0.5).bfloat16()
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    go = torch.randn(B, S, D, device="cuda").bfloat16()
    dqkv = torch.empty_like(qkv)
    o, lse = ops.attn_fwd(q, k, v, H)

    def t(fn, n=50):
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    tf = t(lambda: ops.attn_fwd(q, k, v, H))
    tb = t(lambda: ops.attn_bwd(q, k, v, o, go, lse, H, dqkv[:, :, :D], dqkv[:, :, D:2 * D], dqkv[:, :, 2 * D:]))
    print(f"B={B} S={S} D={D} H={H}: fwd {tf:.1f} us  bwd {tb:.1f} us")
