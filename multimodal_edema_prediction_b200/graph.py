"""Whole-step CUDA-graph capture.  The DuETT step is ~270 kernel launches; at the reference's real MIMIC shape (and, once
the kernels are fast, at the base shape too) the Python/ctypes enqueue time exceeds the GPU time, so the step is
captured once (forward + loss + backward + gradient all-reduce hooks + fused AdamW) and replayed.

    step = CudaGraphStep(fn, static_inputs)      # fn() reads ONLY the static input tensors, returns tensor(s)
    out = step(x=new_x, y=new_y)                 # copies into the static inputs (async H2D from pinned memory), replays

Requirements on fn: static shapes, no host synchronisation (no .item()/.cpu()/torch.tensor(list, device=cuda)), every
kernel launched on the current stream — all true for this package's ops (they take torch.cuda.current_stream()).
The AdamW step counter and LR multiplier live on the device (ddp.FusedAdamW) so replays stay correct.
"""
from __future__ import annotations

import torch


class CudaGraphStep:
    def __init__(self, fn, static_inputs: dict, warmup: int = 3, pool=None):
        self.fn, self.static = fn, static_inputs
        self.graph = None
        # Warm-up on the CURRENT stream (lazy allocations, cudaFuncSetAttribute).  The usual side-stream warm-up makes the
        # autograd engine record cross-stream dependencies on the flat gradient buffer, which a later capture rejects
        # ("dependency created on uncaptured work in another stream").
        out = None
        for _ in range(warmup):          # warmup=0: a second graph of an already warmed-up step (shares `pool`, no eager copy of
            out = fn()                   # the activations next to the first graph's pool)
        del out                          # no eager autograd graph may be released in the middle of the capture
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool):      # pool: share another graph's memory pool (graphs never replayed concurrently)
            self.out = fn()
        self.graph = g

    def __call__(self, **inputs):
        for k, v in inputs.items():
            self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.out
