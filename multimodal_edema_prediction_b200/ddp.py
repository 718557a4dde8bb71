"""Data-parallel plumbing for the DuETT path: flat fp32 parameter/gradient storage, a bucketed gradient all-reduce
that is launched while backward is still running, and the fused AdamW step over the flat buffers.

The reference shards only by data (HF accelerate / Lightning DDP: replicated weights, bucketed gradient all-reduce
overlapped with backward, per-rank BatchNorm statistics — training_duett/trainer.py:217-218,418-419,
duett/train_duett_ssl.py:188-195).  Same algorithm here, driven explicitly:

  * FlatParams lays every trainable parameter out in ONE fp32 buffer in *reverse backward-completion order* (heads first,
    then time/event encoders from the last layer to the first, embeddings last) and makes p.data / p.grad views of it.
    The backbone's kernels accumulate straight into those .grad views (functional.grad_sink), so there is no gradient
    copy and every bucket is a contiguous slice.  With a process group, rank 0's parameters (and the module's buffers:
    BatchNorm running statistics) are broadcast at construction, like DDP does, so replicas start identical whatever
    each rank's seed or checkpoint was.
  * GradReducer receives "these parameters are final" notifications from backbone.encoder_bwd (one per weight matrix
    group: FFN-out, FFN-in, attention) and issues torch.distributed.all_reduce(async_op=True) on that slice — NCCL over
    NVLink 5 / NVSwitch on its own stream, overlapped with the remaining backward kernels.  Everything that was not
    announced (heads, perceiver, embeddings) is reduced in finish(), i.e. after backward has returned, so no assumption
    is made about the order in which autograd finishes those.  The 1/world scaling is folded into the optimizer kernel.
  * FusedAdamW: dx_adamw over contiguous ranges (one launch per LR group), optional global-norm clipping
    (dx_sumsq + dx_clip_factor, no host sync), per-group LR multipliers resident on the device (schedules without
    re-capturing a CUDA graph), bf16 weight shadows written by the same launch (the GEMMs' operands: no cast kernels).
"""
from __future__ import annotations

import math
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


def unused_parameter_names(module: torch.nn.Module):
    """Names of trainable parameters that the module's forward never touches in its current mode, i.e. whose .grad stays
    None under torch autograd — torch.optim.AdamW skips those (no weight decay either), so FusedAdamW must too:
      Model(pretrain=True)  : head.*                                   (duett/duett.py:304-316 returns before the head)
      Model(pretrain=False) : pretrain_*_proj.*, predict_events*_proj.* (duett/duett.py:318)
      DuettFeatureExtractor inside Student/TeacherModel: all of the above (only encode() is called,
                              models/main_architecture_duett.py:1224-1227)."""
    from .duett.duett import Model
    ssl_heads = ("pretrain_value_proj.", "pretrain_presence_proj.", "predict_events_proj.", "predict_events_presence_proj.")
    out = set()
    for mname, m in module.named_modules():
        if not isinstance(m, Model):
            continue
        pre = mname + "." if mname else ""
        backbone_only = mname != ""                       # a Model held by another module is used through encode() only
        for n, p in m.named_parameters():
            if not p.requires_grad:
                continue
            is_ssl, is_head = n.startswith(ssl_heads), n.startswith("head.")
            if backbone_only and (is_ssl or is_head):
                out.add(pre + n)
            elif not backbone_only and ((m.pretrain and is_head) or (not m.pretrain and is_ssl)):
                out.add(pre + n)
    return out


class FlatParams:
    def __init__(self, module: torch.nn.Module, device=None, group=None, broadcast=True):
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        named.sort(key=lambda np_: _order_key(np_[0]))        # stable: keeps registration order inside a class
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        self.module = module
        device = device or self.params[0].device
        # every slice starts on a 512 B boundary of the f32 buffers = 256 B of the bf16 shadow: the shadow views are TMA
        # operands of the tcgen05 GEMMs, and a weight matrix whose 128 B rows straddle two L2 lines costs twice the
        # L2 -> SM requests (measured: +20 % on the HBM-bound N = dim GEMMs with 16 B-aligned slices)
        sizes = [((p.numel() + 127) // 128) * 128 for p in self.params]
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
        self.unused = unused_parameter_names(module)
        self.shadow = None            # bf16 copy of `data` (enable_shadow)
        if broadcast and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            self.broadcast_from_rank0(group)

    def broadcast_from_rank0(self, group=None):
        """DDP's construction-time broadcast: parameters (one flat tensor) and every floating-point / integer buffer
        (BatchNorm running statistics, num_batches_tracked) of the wrapped module."""
        dist.broadcast(self.data, 0, group=group)
        for b in self.module.buffers():
            if b.numel() and b.device == self.data.device:
                dist.broadcast(b, 0, group=group)
        self.sync_shadow()

    # ---- bf16 weight shadows ---------------------------------------------------------------------------------------
    def enable_shadow(self):
        """bf16 mirror of the flat parameter buffer.  FusedAdamW refreshes it inside the optimizer kernel; the backbone uses
        the views as tcgen05 operands instead of casting the fp32 weights every step.  A view is only trusted while the
        parameter's torch version counter is the one recorded at its last refresh (load_state_dict / any torch in-place op
        bumps it -> the backbone falls back to a cast and the next sync re-arms it)."""
        if self.shadow is None:
            self.shadow = torch.empty(self.numel, device=self.data.device, dtype=torch.bfloat16)
            for p, off in zip(self.params, self.offsets):
                if p.dim() >= 2:
                    p._dx_shadow = self.shadow[off:off + p.numel()].view(p.shape)
            self.sync_shadow()
        return self

    def sync_shadow(self):
        if self.shadow is None:
            return
        ops.cast_into(self.data, self.shadow)
        for p in self.params:
            if p.dim() >= 2:
                p._dx_shadow_version = p._version

    def range_of(self, pred):
        """[lo, hi) covering every parameter whose name satisfies pred (they are contiguous by construction)."""
        idx = [i for i, n in enumerate(self.names) if pred(n)]
        if not idx:
            return None
        assert idx == list(range(idx[0], idx[-1] + 1)), "parameters of one bucket must be contiguous"
        return self.offsets[idx[0]], self.offsets[idx[-1] + 1]

    def zero_grad(self):
        self.grad.zero_()


def shadow_of(param: torch.Tensor):
    """The bf16 shadow view of a flat-buffer parameter if it is current, else None."""
    s = getattr(param, "_dx_shadow", None)
    if s is not None and getattr(param, "_dx_shadow_version", -1) == param._version:
        return s
    return None


class GradReducer:
    """Bucketed, backward-overlapped gradient all-reduce (sum; the mean's 1/world goes into the optimizer)."""

    def __init__(self, flat: FlatParams, group=None, bucket_cap_mb: float = 0.0, trace: bool = False):
        self.flat, self.group = flat, group
        # trace: per-bucket CUDA events (eager steps only) -> self.timeline(): when each bucket became ready on the compute
        # stream and when its all-reduce started / finished, relative to start_step() — the evidence for "overlapped"
        self.trace, self._ev, self._t0, self._comm = trace, [], None, None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.pending, self.covered, self.launched = [], [], 0
        self.bucket_cap = int(bucket_cap_mb * (1 << 20) / 4)       # elements; 0 = one bucket per notification
        self._ranges = {}
        self._held = None                                         # small announced range waiting to be merged

    def attach(self, sm_reserve=None):
        """sm_reserve: SMs the persistent GEMM grids leave free for the NCCL kernels that run concurrently with backward
        (default: DX_GEMM_SM_RESERVE or 0; pair it with NCCL_MAX_CTAS so that NCCL stays within them)."""
        if self._on_ready not in backbone.GRAD_READY_HOOKS:
            backbone.GRAD_READY_HOOKS.append(self._on_ready)
        if self.world > 1 and sm_reserve is not None:
            ops.gemm_reserve_sms(sm_reserve)
        return self

    def detach(self):
        if self._on_ready in backbone.GRAD_READY_HOOKS:
            backbone.GRAD_READY_HOOKS.remove(self._on_ready)

    def start_step(self):
        self.pending, self.covered, self.launched, self._held = [], [], 0, None
        if self.trace:
            self._ev = []
            self._t0 = torch.cuda.Event(enable_timing=True)
            self._t0.record()
            if self._comm is None:
                self._comm = torch.cuda.Stream()

    def _on_ready(self, prefix: str, keys=None):
        """prefix = 'time_transformers.3' ...; keys = the parameter names (ENC_KEYS subset) under it that are final now
        (None = all of that encoder's).  Only these announced slices are reduced during backward."""
        if self.world == 1:
            return
        tag = (prefix, None if keys is None else tuple(keys))
        if tag not in self._ranges:
            if keys is None:
                self._ranges[tag] = self.flat.range_of(lambda n: n.startswith(prefix + ".") or ("." + prefix + ".") in n)
            else:
                names = {prefix + "." + k for k in keys}
                self._ranges[tag] = self.flat.range_of(lambda n: any(n == x or n.endswith("." + x) for x in names))
        r = self._ranges[tag]
        if r is None:
            return
        if self._held is not None:           # merge with a held neighbour (announcements arrive in descending address order)
            lo, hi = self._held
            if r[1] == lo:
                r = (r[0], hi)
            elif hi == r[0]:
                r = (lo, r[1])
            else:
                self._launch(lo, hi)
            self._held = None
        if self.bucket_cap and r[1] - r[0] < self.bucket_cap:
            self._held = r
            return
        self._launch(*r)

    def _launch(self, lo, hi):
        if hi <= lo:
            return
        if self.trace:
            main = torch.cuda.current_stream()
            ready, start, end = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            ready.record(main)
            self._comm.wait_stream(main)
            with torch.cuda.stream(self._comm):
                start.record()
                dist.all_reduce(self.flat.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group)   # stream-ordered on _comm
                end.record()
            self._ev.append((lo, hi, ready, start, end))
            self.pending.append(None)
        else:
            self.pending.append(dist.all_reduce(self.flat.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.covered.append((lo, hi))
        self.launched += 1

    def _unused_ranges(self):
        if not hasattr(self, "_unused_cache"):
            f, out, i = self.flat, [], 0
            while i < len(f.names):
                if f.names[i] in f.unused:
                    j = i
                    while j + 1 < len(f.names) and f.names[j + 1] in f.unused:
                        j += 1
                    out.append((f.offsets[i], f.offsets[j + 1]))
                    i = j + 1
                else:
                    i += 1
            self._unused_cache = out
        return self._unused_cache

    def timeline(self):
        """[(MB, ready_ms, start_ms, end_ms)] of the last traced step (call after a device synchronisation)."""
        return [((hi - lo) * 4 / 2 ** 20, self._t0.elapsed_time(r), self._t0.elapsed_time(s), self._t0.elapsed_time(e))
                for lo, hi, r, s, e in self._ev]

    def finish(self):
        """Reduce everything that was not announced during backward (heads, perceiver, embeddings — backward has returned,
        so all of it is final) and make the current stream wait for every bucket."""
        if self.world > 1:
            if self._held is not None:
                self._launch(*self._held)
                self._held = None
            # parameters the active mode never touches (FlatParams.unused: e.g. the SSL projection heads in a supervised run —
            # 538 MB of the stress shape's gradients) carry all-zero gradients on every rank: nothing to reduce
            skip = list(self._unused_ranges())
            pos = 0
            for lo, hi in sorted(self.covered + skip):
                assert lo >= pos, "overlapping all-reduce buckets"
                self._launch(pos, lo)
                pos = hi
            self._launch(pos, self.flat.numel)
            self.covered = [c for c in self.covered if c not in skip]
            if self.trace:
                torch.cuda.current_stream().wait_stream(self._comm)
            for w in self.pending:
                if w is not None:
                    w.wait()
        self.pending = []
        return 1.0 / self.world


class FusedAdamW:
    """AdamW over FlatParams.  groups: list of (name_predicate, lr_scale, weight_decay) evaluated in order; parameters
    matching no predicate use (1.0, weight_decay).  Mirrors training_duett/trainer.py:77-125 (_make_param_groups).
    Parameters the active mode never uses (FlatParams.unused) are skipped entirely, like torch.optim.AdamW skips
    parameters whose .grad is None (no weight decay on an unused head)."""

    def __init__(self, flat: FlatParams, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1, groups=None,
                 max_grad_norm=None, skip_unused=True, group_names=None):
        self.flat, self.lr, self.betas, self.eps, self.wd = flat, lr, betas, eps, weight_decay
        self.m = torch.zeros_like(flat.data)
        self.v = torch.zeros_like(flat.data)
        self.step_count = 0
        self.max_grad_norm = max_grad_norm
        self._ss = torch.zeros(1, device=flat.data.device)
        self._clip = torch.ones(1, device=flat.data.device)
        # device-resident step counter and LR multipliers: a captured CUDA graph of step() stays valid as they advance
        self.step_dev = torch.zeros(1, device=flat.data.device, dtype=torch.int32)
        # contiguous runs of identical (group index, lr_scale, wd); group index -1 = skipped (unused in this mode)
        cfg = []
        for n in flat.names:
            c = (len(groups or []), 1.0, weight_decay)
            for gi, (pred, s, wd) in enumerate(groups or []):
                if pred(n):
                    c = (gi, s, wd)
                    break
            if skip_unused and n in flat.unused:
                c = (-1, 0.0, 0.0)
            cfg.append(c)
        self.runs = []
        i = 0
        while i < len(cfg):
            j = i
            while j + 1 < len(cfg) and cfg[j + 1] == cfg[i]:
                j += 1
            if cfg[i][0] >= 0:
                self.runs.append((flat.offsets[i], flat.offsets[j + 1], cfg[i][1], cfg[i][2], cfg[i][0]))
            i = j + 1
        self.n_groups = len(groups or []) + 1
        self.group_names = list(group_names) if group_names else [f"group{i}" for i in range(self.n_groups - 1)]
        self.group_names = (self.group_names + ["rest"])[:self.n_groups]
        self.group_lr = [lr * next((s for (_, _, s, _, gi) in self.runs if gi == g), 1.0) for g in range(self.n_groups)]
        self.lr_scale_dev = torch.ones(self.n_groups, device=flat.data.device)     # per-group schedule multiplier
        self._sched = None

    # ---- the reference's optimiser recipes -------------------------------------------------------------------------
    @classmethod
    def from_trainer_args(cls, flat: FlatParams, args, total_steps=None, **kw):
        """training_duett/trainer.py:77-125: LR groups backbone ('duett.' / 'cxr.' prefixes, lr x backbone_lr_mult),
        pathology queries (names ending '_queries', lr x query_lr_mult, default 0.2), correction head + beta
        (lr x correction_lr_mult, default 1.0), rest (lr); weight decay args.weight_decay for all
        (trainer.py:383,902).  With total_steps the reference's schedule is attached (advance it with sched_step()):
        LinearLR(1e-4 -> 1, warmup_steps) then CosineAnnealingLR(T_max = total - warmup, eta_min = lr*min_lr_ratio —
        the SAME absolute floor for every group, as in the reference)."""
        g = lambda name, default: float(getattr(args, name, default))
        groups = [
            (lambda n: n.startswith(("duett.", "cxr.")), g("backbone_lr_mult", 0.2), g("weight_decay", 0.05)),
            (lambda n: not n.startswith(("duett.", "cxr.")) and ("correction_head" in n or n.endswith(".beta") or n == "beta"),
             g("correction_lr_mult", 1.0), g("weight_decay", 0.05)),
            (lambda n: n.endswith("_queries"), g("query_lr_mult", 0.2), g("weight_decay", 0.05)),
        ]
        opt = cls(flat, lr=float(args.lr), weight_decay=g("weight_decay", 0.05), groups=groups,
                  group_names=["backbone", "correction_head", "pathology_queries"], **kw)
        if total_steps is not None:
            opt.set_schedule(WarmupCosine(int(getattr(args, "warmup_steps", 300)), int(total_steps),
                                          float(args.lr) * g("min_lr_ratio", 0.01)))
        return opt

    @classmethod
    def for_ssl(cls, flat: FlatParams, lr=3e-4, weight_decay=0.1, warmup_steps=2000, **kw):
        """duett/duett.py:325-327 + duett/train_duett_ssl.py:27-50,191: AdamW(lr, wd), WarmUpCallback(steps) with inverse
        square-root decay, gradient_clip_val=1.0."""
        kw.setdefault("max_grad_norm", 1.0)
        opt = cls(flat, lr=lr, weight_decay=weight_decay, **kw)
        opt.set_schedule(WarmupInvSqrt(warmup_steps))
        return opt

    def set_schedule(self, sched):
        self._sched = sched
        self._sched_t = 0
        self._apply_schedule()

    def _apply_schedule(self):
        f = [self._sched.factor(self._sched_t, self.group_lr[g]) for g in range(self.n_groups)]
        self.set_lr_scales(f)

    def sched_step(self):
        """Advance the attached schedule by one optimisation step (host arithmetic, one small async H2D copy)."""
        self._sched_t += 1
        self._apply_schedule()

    def current_lrs(self):
        return {n: self.group_lr[g] * float(s) for g, (n, s) in enumerate(zip(self.group_names, self.lr_scale_dev.tolist()))}

    def zero_grad(self, set_to_none=False):
        self.flat.zero_grad()

    def set_lr_scale(self, scale: float):
        """One LR-schedule multiplier for every group, applied on the device."""
        self.lr_scale_dev.fill_(float(scale))

    def set_lr_scales(self, scales):
        """Per-group multipliers (same order as the `groups` argument, 'rest' last)."""
        assert len(scales) == self.n_groups
        t = torch.tensor([float(s) for s in scales], dtype=torch.float32)
        if self.lr_scale_dev.is_cuda:
            t = t.pin_memory()
        self.lr_scale_dev.copy_(t, non_blocking=True)

    def step(self, grad_scale=1.0, lr=None):
        self.step_count += 1
        self.step_dev.add_(1)
        ops.advance_drop_step()      # dropout sites add this device counter to their seeds: fresh masks on a replayed graph
        clip = None
        if self.max_grad_norm is not None:
            self._ss.zero_()
            ops.sumsq(self.flat.grad, self._ss)
            # ||grad_scale * g|| = grad_scale * sqrt(ss): clip on the scaled norm
            ops.clip_factor(self._ss, self.max_grad_norm / max(grad_scale, 1e-30), self._clip)
            clip = self._clip
        f = self.flat
        base = self.lr if lr is None else lr
        for lo, hi, s, wd, gi in self.runs:
            ops.adamw(f.data[lo:hi], f.grad[lo:hi], self.m[lo:hi], self.v[lo:hi], base * s, self.betas, self.eps, wd,
                      self.step_count, grad_scale_dev=clip, grad_scale=grad_scale, step_dev=self.step_dev,
                      lr_scale_dev=self.lr_scale_dev[gi:gi + 1], shadow=None if f.shadow is None else f.shadow[lo:hi])


class WarmupCosine:
    """torch SequentialLR([LinearLR(1e-4 -> 1, warmup), CosineAnnealingLR(T_max, eta_min)], milestones=[warmup]) in closed
    form (training_duett/trainer.py:119-125); eta_min is an absolute LR, so the factor depends on the group's base LR."""

    def __init__(self, warmup_steps, total_steps, eta_min, start_factor=1e-4):
        self.warmup = max(int(warmup_steps), 1)
        self.t_max = max(int(total_steps) - self.warmup, 1)
        self.eta_min, self.start = float(eta_min), float(start_factor)

    def factor(self, t, base_lr):
        if t < self.warmup:
            return self.start + (1.0 - self.start) * t / self.warmup
        tc = t - self.warmup
        lr = self.eta_min + (base_lr - self.eta_min) * (1.0 + math.cos(math.pi * tc / self.t_max)) / 2.0
        return lr / base_lr


class WarmupInvSqrt:
    """duett/train_duett_ssl.py:27-50 (WarmUpCallback): lr = s/steps*base while s < steps, then base*sqrt(decay/(s-steps+decay))."""

    def __init__(self, steps=2000, invsqrt=True, decay=None):
        self.steps, self.invsqrt, self.decay = int(steps), invsqrt, (decay or steps)

    def factor(self, t, base_lr):
        if t < self.steps:
            return t / self.steps
        if self.invsqrt:
            return (self.decay / (t - self.steps + self.decay)) ** 0.5
        return 1.0
