"""Data-parallel plumbing for the DuETT path: flat fp32 parameter/gradient storage, a bucketed gradient all-reduce
that is launched while backward is still running, and the fused AdamW step over the flat buffers.

The reference shards only by data (HF accelerate / Lightning DDP: replicated weights, bucketed gradient all-reduce
overlapped with backward, per-rank BatchNorm statistics — training_duett/trainer.py:217-218,418-419,
duett/train_duett_ssl.py:188-195).  Same algorithm here, driven explicitly:

  * FlatParams lays every trainable parameter out in ONE fp32 buffer in *reverse backward-completion order* (heads first,
    then time/event encoders from the last layer to the first, embeddings last) and makes p.data / p.grad views of it.
    The backbone's kernels accumulate straight into those .grad views (functional.grad_sink), so there is no gradient
    copy and every bucket is a contiguous slice.
  * GradReducer receives "encoder l finished" notifications from backbone.DuettEncodeFn.backward and issues
    torch.distributed.all_reduce(async_op=True) on the finished slice — NCCL over NVLink 5 / NVSwitch on its own
    stream, overlapped with the remaining backward kernels.  The 1/world scaling is folded into the optimizer kernel.
  * FusedAdamW: dx_adamw over contiguous ranges (one launch per LR group), optional global-norm clipping
    (dx_sumsq + dx_clip_factor, no host sync).
"""
from __future__ import annotations

import re

import torch
import torch.distributed as dist

from . import backbone, ops


def _order_key(name: str):
    """Reverse backward-completion order: larger key = finished later in backward."""
    m = re.search(r"(event|time)_transformers\.(\d+)\.", name)
    if m:
        layer = int(m.group(2))
        return (1, -layer, 0 if m.group(1) == "time" else 1)
    if any(s in name for s in ("embedding_layers", "special_embeddings", "n_obs_embedding", "tab_encoder",
                               "full_time_embedding", "full_rep_embedding", "full_event_embedding")):
        return (2, 0, 0)
    return (0, 0, 0)          # heads / perceiver / everything after the backbone: gradients arrive first


class FlatParams:
    def __init__(self, module: torch.nn.Module, device=None):
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        named.sort(key=lambda np_: _order_key(np_[0]))        # stable: keeps registration order inside a class
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        device = device or self.params[0].device
        sizes = [((p.numel() + 7) // 8) * 8 for p in self.params]      # 32 B aligned slices (vector kernels, NCCL)
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + s)
        n = self.offsets[-1]
        self.data = torch.zeros(n, device=device, dtype=torch.float32)
        self.grad = torch.zeros(n, device=device, dtype=torch.float32)
        for p, off in zip(self.params, self.offsets):
            v = self.data[off:off + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
            p.grad = self.grad[off:off + p.numel()].view(p.shape)
        self.numel = n

    def range_of(self, pred):
        """[lo, hi) covering every parameter whose name satisfies pred (they are contiguous by construction)."""
        idx = [i for i, n in enumerate(self.names) if pred(n)]
        if not idx:
            return None
        assert idx == list(range(idx[0], idx[-1] + 1)), "parameters of one bucket must be contiguous"
        return self.offsets[idx[0]], self.offsets[idx[-1] + 1]

    def zero_grad(self):
        self.grad.zero_()


class GradReducer:
    """Bucketed, backward-overlapped gradient all-reduce (sum; the mean's 1/world goes into the optimizer)."""

    def __init__(self, flat: FlatParams, group=None):
        self.flat, self.group = flat, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.pending, self.done_hi, self.launched = [], 0, 0
        self._ranges = {}

    def attach(self):
        backbone.GRAD_READY_HOOK = self._on_ready
        return self

    def detach(self):
        backbone.GRAD_READY_HOOK = None

    def start_step(self):
        self.pending, self.done_hi, self.launched = [], 0, 0

    def _on_ready(self, tag: str):
        """tag = 'time_transformers.3' / 'event_transformers.3' ...: all gradients up to and including that encoder are
        final (heads finished before the backbone backward started)."""
        if self.world == 1:
            return
        if tag not in self._ranges:
            self._ranges[tag] = self.flat.range_of(lambda n: tag + "." in n)
        r = self._ranges[tag]
        if r is None:
            return
        self._launch(r[1])

    def _launch(self, hi):
        if hi <= self.done_hi:
            return
        self.pending.append(dist.all_reduce(self.flat.grad[self.done_hi:hi], op=dist.ReduceOp.SUM, group=self.group,
                                            async_op=True))
        self.done_hi = hi
        self.launched += 1

    def finish(self):
        """Reduce whatever is left (embeddings) and make the current stream wait for every bucket."""
        if self.world > 1:
            self._launch(self.flat.numel)
            for w in self.pending:
                w.wait()
        self.pending = []
        return 1.0 / self.world


class FusedAdamW:
    """AdamW over FlatParams.  groups: list of (name_predicate, lr_scale, weight_decay) evaluated in order; parameters
    matching no predicate use (1.0, weight_decay).  Mirrors training_duett/trainer.py:77-125 (_make_param_groups)."""

    def __init__(self, flat: FlatParams, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1, groups=None,
                 max_grad_norm=None):
        self.flat, self.lr, self.betas, self.eps, self.wd = flat, lr, betas, eps, weight_decay
        self.m = torch.zeros_like(flat.data)
        self.v = torch.zeros_like(flat.data)
        self.step_count = 0
        self.max_grad_norm = max_grad_norm
        self._ss = torch.zeros(1, device=flat.data.device)
        self._clip = torch.ones(1, device=flat.data.device)
        # device-resident step counter and LR multiplier: a captured CUDA graph of step() stays valid as they advance
        self.step_dev = torch.zeros(1, device=flat.data.device, dtype=torch.int32)
        self.lr_scale_dev = torch.ones(1, device=flat.data.device)
        # contiguous runs of identical (lr_scale, wd)
        cfg = []
        for n in flat.names:
            c = (1.0, weight_decay)
            for pred, s, wd in (groups or []):
                if pred(n):
                    c = (s, wd)
                    break
            cfg.append(c)
        self.runs = []
        i = 0
        while i < len(cfg):
            j = i
            while j + 1 < len(cfg) and cfg[j + 1] == cfg[i]:
                j += 1
            self.runs.append((flat.offsets[i], flat.offsets[j + 1], cfg[i][0], cfg[i][1]))
            i = j + 1

    def zero_grad(self, set_to_none=False):
        self.flat.zero_grad()

    def set_lr_scale(self, scale: float):
        """LR-schedule multiplier (warm-up / cosine / inv-sqrt), applied on the device."""
        self.lr_scale_dev.fill_(float(scale))

    def step(self, grad_scale=1.0, lr=None):
        self.step_count += 1
        self.step_dev.add_(1)
        ops.advance_drop_step()      # dropout sites add this device counter to their seeds: fresh masks on a replayed graph
        lr = self.lr if lr is None else lr
        clip = None
        if self.max_grad_norm is not None:
            self._ss.zero_()
            ops.sumsq(self.flat.grad, self._ss)
            # ||grad_scale * g|| = grad_scale * sqrt(ss): clip on the scaled norm
            ops.clip_factor(self._ss, self.max_grad_norm / max(grad_scale, 1e-30), self._clip)
            clip = self._clip
        f = self.flat
        for lo, hi, s, wd in self.runs:
            ops.adamw(f.data[lo:hi], f.grad[lo:hi], self.m[lo:hi], self.v[lo:hi], lr * s, self.betas, self.eps, wd,
                      self.step_count, grad_scale_dev=clip, grad_scale=grad_scale, step_dev=self.step_dev,
                      lr_scale_dev=self.lr_scale_dev)
