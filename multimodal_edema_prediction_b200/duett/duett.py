"""B200-native drop-in for the reference's duett/duett.py (`Model`, `pretrain_model`, `fine_tune_model`).

Same constructor signature, attribute names, state-dict keys, batch format and Lightning step surface as the reference
(duett/duett.py:41-56,159-237,239-323,325-495); the arithmetic runs in hand-written sm_100a kernels through
libduett_b200.so (backbone.py / functional.py / ops.py).  There is no CPU path: tensors must live on a B200.

Not inherited from the reference (documented divergences):
  * `forward` does not overwrite the caller's xs_feats count columns in place (duett/duett.py:252);
  * dropout (transformer_dropout on attention probabilities and FFN hidden, nn.Dropout in the MLP heads) draws its masks
    from the kernels' own counter-based generator (ops.drop_seed / dx_dropout), not from torch's Philox stream: same
    distribution and scaling, different bits — the reference's masks are not a portable contract;
  * Lightning is not in the image, so `Model` is an nn.Module that duck-types the LightningModule hooks it needs
    (`device`, `log`, `training_step`, `configure_optimizers`, `load_from_checkpoint`, `on_load_checkpoint`, `freeze`).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops, state_keys
from ..backbone import DuettEncodeFn, ENC_KEYS, TimeEmbedAssembleFn
from ..functional import (BatchNorm2dFn, BCELogitsFn, GatherVecFn, MaskedMseBceFn, cast, dropout, linear)


# ------------------------------------------------------------------------------------------------------------------
# parameter holders with the reference's module structure (so state-dict keys line up); forward = CUDA kernels
# ------------------------------------------------------------------------------------------------------------------
class BatchNormLastDim(nn.Module):
    """duett/duett.py:11-22.  Statistics over all leading dims; running buffers updated by the kernel."""

    def __init__(self, d, **kwargs):
        super().__init__()
        self.batch_norm = nn.BatchNorm1d(d, **kwargs)

    def forward(self, x):
        if x.ndim not in (2, 3):
            raise NotImplementedError("BatchNormLastDim not implemented for ndim > 3 yet")
        bn = self.batch_norm
        x2 = x.reshape(-1, x.shape[-1]).float()
        y = BatchNorm2dFn.apply(x2, bn.weight, bn.bias, bn.running_mean, bn.running_var, self.training)
        if self.training:
            bn.num_batches_tracked += 1
        return y.reshape(x.shape)


class DxSequential(nn.Sequential):
    """nn.Sequential of Linear / activation / Dropout / BatchNormLastDim holders, executed with fused kernels:
    a Linear followed by ReLU/Tanh/GELU becomes one GEMM with the activation in its epilogue."""

    _ACT = {nn.ReLU: ops.ACT_RELU, nn.Tanh: ops.ACT_TANH, nn.GELU: ops.ACT_GELU}

    def forward(self, x):
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Linear):
                act = ops.ACT_NONE
                if i + 1 < len(mods) and type(mods[i + 1]) in self._ACT:
                    act = self._ACT[type(mods[i + 1])]
                    i += 1
                x = linear(x if x.dtype == torch.float32 else cast(x, torch.float32), m.weight, m.bias, act)
            elif isinstance(m, nn.Dropout):
                x = dropout(x, m.p, self.training)
            elif isinstance(m, (BatchNormLastDim, nn.LayerNorm)) or hasattr(m, "dx_forward"):
                x = m(x)
            else:
                raise NotImplementedError(f"DxSequential: unsupported layer {type(m).__name__}")
            i += 1
        return x


def simple_mlp(d_in, d_out, n_hidden, d_hidden, final_activation=False, input_batch_norm=False,
               hidden_batch_norm=False, dropout=0., activation=nn.ReLU):
    """Same layer list (hence the same state-dict indices) as duett/duett.py:24-39."""
    layers = [BatchNormLastDim(d_in)] if input_batch_norm else []
    if n_hidden == 0:
        layers += [nn.Linear(d_in, d_out)]
    else:
        layers += [nn.Linear(d_in, d_hidden), activation(), nn.Dropout(dropout)]
        for _ in range(n_hidden - 1):
            layers += ([BatchNormLastDim(d_hidden)] if hidden_batch_norm else []) + \
                      [nn.Linear(d_hidden, d_hidden), activation(), nn.Dropout(dropout)]
        layers += ([BatchNormLastDim(d_hidden)] if hidden_batch_norm else []) + [nn.Linear(d_hidden, d_out)]
    if final_activation:
        layers.append(activation())
    return DxSequential(*layers)


class StackedEmbeddingMLPs(nn.Module):
    """The V per-variable simple_mlp(2, d, 1, 64, hidden_batch_norm=True) of duett/duett.py:84-86 as stacked tensors."""

    def __init__(self, V, d, d_hidden):
        super().__init__()
        if d_hidden != 64:
            raise NotImplementedError("d_hidden_mlp_embedding must be 64 (kernel tile width)")
        k0, k4 = 1 / math.sqrt(2), 1 / math.sqrt(d_hidden)
        self.w0 = nn.Parameter((torch.rand(V, d_hidden, 2) * 2 - 1) * k0)
        self.b0 = nn.Parameter((torch.rand(V, d_hidden) * 2 - 1) * k0)
        self.bn_w = nn.Parameter(torch.ones(V, d_hidden))
        self.bn_b = nn.Parameter(torch.zeros(V, d_hidden))
        self.register_buffer("bn_rm", torch.zeros(V, d_hidden))
        self.register_buffer("bn_rv", torch.ones(V, d_hidden))
        self.register_buffer("bn_nbt", torch.zeros(V, dtype=torch.long))
        self.w4 = nn.Parameter((torch.rand(V, d, d_hidden) * 2 - 1) * k4)
        self.b4 = nn.Parameter((torch.rand(V, d) * 2 - 1) * k4)

    def __len__(self):
        return self.w0.shape[0]


class AxisEncoder(nn.Module):
    """Parameters of one x_transformers.Encoder(dim, depth=1, heads, pre_norm, use_scalenorm, attn_dim_head=d//heads,
    ff_mult=d_ff/dim) — duett/duett.py:95-105.  q/k/v are stored fused as wqkv [3d, dim]."""

    def __init__(self, dim, heads, dim_head, ff_mult, dropout=0.0):
        super().__init__()
        inner = heads * dim_head
        ff_inner = int(dim * ff_mult)            # x_transformers' float expression, kept verbatim (SURVEY §8 note)
        self.dim, self.heads, self.inner, self.ff_inner, self.dropout = dim, heads, inner, ff_inner, dropout
        u = lambda o, i: (torch.rand(o, i) * 2 - 1) / math.sqrt(i)
        self.g_attn = nn.Parameter(torch.ones(1))
        self.wqkv = nn.Parameter(torch.cat([u(inner, dim) for _ in range(3)], 0))
        self.wo = nn.Parameter(u(dim, inner))
        self.g_ff = nn.Parameter(torch.ones(1))
        self.w1 = nn.Parameter(u(ff_inner, dim))
        self.b1 = nn.Parameter((torch.rand(ff_inner) * 2 - 1) / math.sqrt(dim))
        self.w2 = nn.Parameter(u(dim, ff_inner))
        self.b2 = nn.Parameter((torch.rand(dim) * 2 - 1) / math.sqrt(ff_inner))
        self.g_final = nn.Parameter(torch.ones(1))


class _Metric:
    """Minimal AUROC / AveragePrecision accumulator (torchmetrics is not in the image; sklearn scores on the host)."""

    def __init__(self, kind):
        self.kind, self.p, self.y = kind, [], []

    def update(self, preds, target):
        self.p.append(preds.detach().float()); self.y.append(target.detach())     # stays on the device: no per-step sync

    def reset(self):
        self.p, self.y = [], []

    def compute(self):
        from sklearn.metrics import average_precision_score, roc_auc_score
        p, y = torch.cat(self.p).cpu().numpy(), torch.cat(self.y).cpu().numpy()
        return torch.tensor(roc_auc_score(y, p) if self.kind == "auroc" else average_precision_score(y, p))


class _Staging:
    """Pinned, double-buffered host staging for Model._upload.  One pair of flat buffers per (slot, dtype), sized for the
    largest batch seen so far and viewed at the requested shape — ragged batches (a different pad length or a short last
    batch every step) therefore do not accumulate one pinned allocation per distinct shape.  acquire() hands out the buffer
    whose previous copy has completed (waits on its event if necessary); the caller records an event after enqueueing the
    H2D copy and passes it to release()."""

    def __init__(self, pin=True):
        self.pin, self.slots = pin, {}

    def acquire(self, slot, shape, dtype):
        n = 1
        for d in shape:
            n *= int(d)
        s = self.slots.setdefault((slot, dtype), {"bufs": [None, None], "evts": [None, None], "i": 0})
        i = s["i"]
        s["i"] = 1 - i
        if s["evts"][i] is not None:
            s["evts"][i].synchronize()          # the copy that last used this buffer has completed
            s["evts"][i] = None
        if s["bufs"][i] is None or s["bufs"][i].numel() < n:
            s["bufs"][i] = torch.empty(max(n, 1), dtype=dtype, pin_memory=self.pin)

        def release(event):
            s["evts"][i] = event

        return s["bufs"][i][:n].view(shape), release

    def pinned_bytes(self):
        return sum(b.numel() * b.element_size() for s in self.slots.values() for b in s["bufs"] if b is not None)

    # a copied / pickled Model (copy.deepcopy for an EMA replica, torch.save(model)) starts with empty staging: the buffers
    # are scratch and CUDA events cannot be copied
    def __deepcopy__(self, memo):
        return _Staging(self.pin)

    def __reduce__(self):
        return (_Staging, (self.pin,))


def pretrain_model(d_static_num, d_time_series_num, d_target, **kwargs):
    return Model(d_static_num, d_time_series_num, d_target, **kwargs)


def fine_tune_model(ckpt_path, **kwargs):
    return Model.load_from_checkpoint(ckpt_path, pretrain=False, aug_noise=0., aug_mask=0.5, transformer_dropout=0.5,
                                      lr=1.e-4, weight_decay=1.e-5, fusion_method='rep_token', **kwargs)


class Model(nn.Module):
    def __init__(self, d_static_num, d_time_series_num, d_target, lr=3.e-4, weight_decay=1.e-1, glu=False,
                 scalenorm=True, n_hidden_mlp_embedding=1, d_hidden_mlp_embedding=64, d_embedding=24, d_feedforward=512,
                 max_len=48, n_transformer_head=2, n_duett_layers=2, d_hidden_tab_encoder=128, n_hidden_tab_encoder=1,
                 norm_first=True, fusion_method='masked_embed', n_hidden_head=1, d_hidden_head=64, aug_noise=0.,
                 aug_mask=0., pretrain=True, pretrain_masked_steps=1, pretrain_n_hidden=0, pretrain_d_hidden=64,
                 pretrain_dropout=0.5, pretrain_value=True, pretrain_presence=True, pretrain_presence_weight=0.2,
                 predict_events=True, transformer_dropout=0., pos_frac=None, freeze_encoder=False, seed=42,
                 save_representation=None, masked_transform_timesteps=32, precision="auto", **kwargs):
        super().__init__()
        if glu or not scalenorm or not norm_first or n_hidden_mlp_embedding != 1:
            raise NotImplementedError("B200 path covers the reference's configuration: pre-norm ScaleNorm, no GLU, "
                                      "1-hidden-layer embedding MLPs")
        if d_embedding % 8 or d_embedding % n_transformer_head:
            raise ValueError("d_embedding must be a multiple of 8 and of n_transformer_head")
        if int(pretrain_masked_steps) < 1:
            raise ValueError("pretrain_masked_steps must be >= 1")
        self.lr, self.weight_decay = lr, weight_decay
        self.d_time_series_num, self.d_target, self.d_embedding = d_time_series_num, d_target, d_embedding
        self.max_len, self.pretrain = max_len, pretrain
        self.pretrain_masked_steps, self.pretrain_dropout = pretrain_masked_steps, pretrain_dropout
        self.freeze_encoder = freeze_encoder
        self.set_pos_frac(pos_frac)
        self.rng = np.random.default_rng(seed)
        self.aug_noise, self.aug_mask, self.fusion_method = aug_noise, aug_mask, fusion_method
        self.pretrain_presence, self.pretrain_presence_weight = pretrain_presence, pretrain_presence_weight
        self.predict_events, self.masked_transform_timesteps = predict_events, masked_transform_timesteps
        self.pretrain_value, self.save_representation = pretrain_value, save_representation
        self.n_transformer_head, self.n_duett_layers = n_transformer_head, n_duett_layers
        self.transformer_dropout = transformer_dropout
        # "auto" (follow torch.autocast), "bf16" (tcgen05 kind::f16), "tf32" (fp32 storage, tcgen05 kind::tf32 contractions — the
        # reference's SSL / fine-tune precision, duett/duett.py:9) or "fp32" (exact FFMA products: the 1e-3 parity mode)
        if precision not in ("auto", "bf16", "tf32", "fp32"):
            raise ValueError(f"precision must be auto / bf16 / tf32 / fp32, got {precision!r}")
        self.precision = precision
        self.final_norm = True
        self.register_buffer("MASKED_EMBEDDING_KEY", torch.tensor(0))
        self.register_buffer("REPRESENTATION_EMBEDDING_KEY", torch.tensor(1))

        self.special_embeddings = nn.Embedding(8, d_embedding)
        self.embedding_layers = StackedEmbeddingMLPs(d_time_series_num, d_embedding, d_hidden_mlp_embedding)
        self.n_obs_embedding = nn.Embedding(16, 1)
        if d_feedforward is None:
            d_feedforward = d_embedding * 4
        et_dim = d_embedding * (masked_transform_timesteps + 1)
        tt_dim = d_embedding * (d_time_series_num + 1)
        dh = d_embedding // n_transformer_head
        self.event_transformers = nn.ModuleList([AxisEncoder(et_dim, n_transformer_head, dh, d_feedforward / et_dim,
                                                             transformer_dropout) for _ in range(n_duett_layers)])
        self.full_event_embedding = nn.Embedding(d_time_series_num + 1, et_dim)
        self.time_transformers = nn.ModuleList([AxisEncoder(tt_dim, n_transformer_head, dh, d_feedforward / tt_dim,
                                                            transformer_dropout) for _ in range(n_duett_layers)])
        self.full_time_embedding = self.cve(batch_norm=True, d_embedding=tt_dim)
        self.full_rep_embedding = nn.Embedding(tt_dim, 1)

        d_representation = d_embedding * (d_time_series_num + 1)
        self.head = simple_mlp(d_representation, d_target, n_hidden_head, d_hidden_head, hidden_batch_norm=True,
                               final_activation=False, activation=nn.ReLU)
        self.pretrain_value_proj = simple_mlp(d_representation, d_time_series_num, pretrain_n_hidden, pretrain_d_hidden,
                                              hidden_batch_norm=True)
        if self.pretrain_presence:
            self.pretrain_presence_proj = simple_mlp(d_representation, d_time_series_num, pretrain_n_hidden,
                                                     pretrain_d_hidden, hidden_batch_norm=True)
        if self.predict_events:
            self.predict_events_proj = simple_mlp(et_dim, masked_transform_timesteps, pretrain_n_hidden,
                                                  pretrain_d_hidden, hidden_batch_norm=True)
            if self.pretrain_presence:
                self.predict_events_presence_proj = simple_mlp(et_dim, masked_transform_timesteps, pretrain_n_hidden,
                                                               pretrain_d_hidden, hidden_batch_norm=True)
        self.tab_encoder = simple_mlp(d_static_num, d_embedding, n_hidden_tab_encoder, d_hidden_tab_encoder,
                                      hidden_batch_norm=True)

        self.train_auroc, self.val_auroc, self.test_auroc = _Metric("auroc"), _Metric("auroc"), _Metric("auroc")
        self.train_ap, self.val_auprc, self.test_auprc = _Metric("ap"), _Metric("ap"), _Metric("ap")
        self.current_epoch = 0
        self._register_state_dict_hook(lambda mod, sd, prefix, meta: state_keys.to_reference(sd, prefix))
        self._register_load_state_dict_pre_hook(
            lambda sd, prefix, *a: state_keys.from_reference(sd, prefix))

    # ---- LightningModule duck-typing -----------------------------------------------------------------------------
    @property
    def device(self):
        return self.special_embeddings.weight.device

    def log(self, *a, **k):
        pass

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, strict=True, map_location="cpu", **kwargs):
        """Reads a Lightning .ckpt ({'state_dict': ...}) — models/main_architecture_duett.py:106-118."""
        ckpt = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
        model = cls(**kwargs)
        model.on_load_checkpoint(ckpt)
        model.load_state_dict(ckpt["state_dict"], strict=strict)
        return model

    def on_load_checkpoint(self, checkpoint):
        """Tolerant loading (duett/duett.py:459-487): missing keys keep their init, reshaped head* keys are skipped,
        unknown keys are dropped, optimizer state is discarded when anything changed."""
        print('Loading from checkpoint')
        state_dict = checkpoint["state_dict"]
        model_state_dict = self.state_dict()
        is_changed = False
        # The tolerant logic below backfills missing keys and drops unknown ones.  For the axis encoders that would turn a
        # key-layout mismatch into SILENTLY re-initialised transformers: the keys inside {event,time}_transformers.{l} belong
        # to x_transformers, which the reference does not pin (state_keys.ENC_MAP / QKV hold the 1.x/2.x layout).  Refuse.
        import re
        enc = re.compile(r"^(event|time)_transformers\.\d+\.")
        missing = sorted(k for k in model_state_dict if enc.match(k) and k not in state_dict)
        unknown = sorted(k for k in state_dict if enc.match(k) and k not in model_state_dict)
        if missing or unknown:
            raise RuntimeError(
                "checkpoint and model disagree on the x_transformers key layout of the axis encoders; loading would leave "
                f"transformer weights randomly initialised.\n  expected but absent ({len(missing)}): {missing[:8]}\n  present but "
                f"unmapped ({len(unknown)}): {unknown[:8]}\nAdd the checkpoint's layout to state_keys.ENC_MAP / state_keys.QKV.")
        for k in model_state_dict:
            if k not in state_dict:
                state_dict[k] = model_state_dict[k]
                is_changed = True
        for k in list(state_dict):
            if k in model_state_dict:
                if k.startswith('head') and state_dict[k].shape != model_state_dict[k].shape:
                    print(f"Skip loading parameter: {k}, required shape: {model_state_dict[k].shape}, "
                          f"loaded shape: {state_dict[k].shape}")
                    state_dict[k] = model_state_dict[k]
                    is_changed = True
            else:
                print(f"Dropping parameter {k}")
                del state_dict[k]
                is_changed = True
        if is_changed:
            checkpoint.pop("optimizer_states", None)
        if self.freeze_encoder:
            self.freeze()

    def freeze(self):
        print('Freezing')
        for n, w in self.named_parameters():
            if "head" not in n:
                w.requires_grad = False
            else:
                print("Skip freezing:", n)

    def set_pos_frac(self, pos_frac):
        if type(pos_frac) == list:
            pos_frac = torch.tensor(pos_frac, device=torch.device('cuda'))
        self.pos_frac = pos_frac
        if pos_frac is not None:
            self.pos_weight = 1 / (2 * pos_frac)
            self.neg_weight = 1 / (2 * (1 - pos_frac))

    def cve(self, d_embedding=None, batch_norm=False):
        if d_embedding is None:
            d_embedding = self.d_embedding
        d_hidden = int(np.sqrt(d_embedding))
        if batch_norm:
            return DxSequential(nn.Linear(1, d_hidden), nn.Tanh(), BatchNormLastDim(d_hidden),
                                nn.Linear(d_hidden, d_embedding))
        return DxSequential(nn.Linear(1, d_hidden), nn.Tanh(), nn.Linear(d_hidden, d_embedding))

    def configure_optimizers(self):
        return [torch.optim.AdamW(self.parameters(), lr=self.lr, weight_decay=self.weight_decay)]

    # ---- host-side batch assembly (kept on the host, same semantics as duett/duett.py:159-237) -------------------
    def feats_to_input(self, x, batch_size, limits=None):
        xs_ts, xs_static, times = x
        xs_ts, times = list(xs_ts), list(times)
        if len(xs_ts) and not torch.is_tensor(xs_ts[0]) and hasattr(xs_ts[0], "slot"):
            # raw event rows from MIMICDataset.__getitem__ (host-only, DataLoader-worker safe): one dx_bin_events launch bins
            # the whole batch on the device (duett/mimic_dataset.py:33-46 semantics), here in the main process
            from .mimic_dataset import bin_stay_rows
            xs_ts = list(bin_stay_rows(xs_ts, self.device).unbind(0))
        augment = self.training and (self.aug_noise > 0 or self.aug_mask > 0) and not self.pretrain
        if not augment and all(f.shape[0] <= self.max_len for f in xs_ts):
            # plain batch (evaluation, SSL, KD student): no per-sample work on the host - the samples are stacked into
            # the pinned staging buffer as they are and the all-zero mask column is appended on the device
            n_timesteps = [len(ts) for ts in times]
            pad_to = int(np.max(n_timesteps))
            dev = self.device
            if any(n != pad_to for n in n_timesteps):
                xs_ts = [F.pad(t, (0, 0, 0, pad_to - t.shape[0])) for t in xs_ts]
                times = [F.pad(t, (0, pad_to - t.shape[0])) for t in times]
            raw = self._upload(xs_ts, "ts_raw", dev)
            xs_ts = torch.cat((raw, raw.new_zeros(raw.shape[0], raw.shape[1], 1)), dim=2)
            return self._upload(list(xs_static), "static", dev), xs_ts, self._upload(times, "times", dev), n_timesteps
        for i, f in enumerate(xs_ts):
            n_vars = f.shape[1] // 2
            if f.shape[0] > self.max_len:
                f = f[-self.max_len:]
                times[i] = times[i][-self.max_len:]
            if self.training and self.aug_noise > 0 and not self.pretrain:
                f = f.clone()
                f[:, :n_vars] += self.aug_noise * torch.randn_like(f[:, :n_vars]) * f[:, n_vars:]
            f = torch.cat((f, torch.zeros_like(f[:, :1])), dim=1)
            if self.training and self.aug_mask > 0 and not self.pretrain:
                mask = torch.rand(f.shape[0]) < self.aug_mask
                f[mask, :] = 0.
                f[mask, -1] = 1.
            xs_ts[i] = f
        n_timesteps = [len(ts) for ts in times]
        pad_to = int(np.max(n_timesteps))
        dev = self.device
        ragged = any(n != pad_to for n in n_timesteps)
        if ragged:
            xs_ts = [F.pad(t, (0, 0, 0, pad_to - t.shape[0])) for t in xs_ts]
            times = [F.pad(t, (0, pad_to - t.shape[0])) for t in times]
        xs_ts = self._upload(xs_ts, "ts", dev)
        xs_times = self._upload(times, "times", dev)
        xs_static = self._upload(list(xs_static), "static", dev)
        if self.training and self.aug_noise > 0 and not self.pretrain:
            xs_static = xs_static + self.aug_noise * torch.randn_like(xs_static)
        return xs_static, xs_ts, xs_times, n_timesteps

    def _upload(self, tensors, slot, dev):
        """Stack per-sample tensors and move them to the model's device with ONE copy.  Host inputs are stacked into a
        pinned, double-buffered staging area (_Staging, so the H2D copy is asynchronous; engine._move_lists leaves the
        per-sample tuples on the host for exactly this); inputs already on the device (device-binned StayRows, callers that
        moved their samples themselves) are stacked there."""
        t0 = tensors[0]
        if t0.is_cuda or dev.type != "cuda":
            if any(t.is_cuda != t0.is_cuda for t in tensors):          # device-binned x_ts next to host static / times
                tensors = [t.to(dev, non_blocking=True) for t in tensors]
            return torch.stack(tensors).to(dev, non_blocking=True)
        shape = (len(tensors),) + tuple(t0.shape)
        st = self.__dict__.setdefault("_staging", _Staging(pin=True))
        buf, release = st.acquire(slot, shape, t0.dtype)
        torch.stack(tensors, out=buf)
        out = buf.to(dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        release(ev)
        return out

    def pretrain_prep_batch(self, x, batch_size):
        """SSL masking with the host numpy RNG, byte-identical to duett/duett.py:189-237 (same draw order: per sample the
        timestep(s) then one variable, then the [B,V] variable-dropout matrix); index work vectorised on the host.
        pretrain_masked_steps = k > 1 (:199-203): k timesteps per sample in ONE rng.choice call, with replacement; targets
        become [B,k,V] in draw order.  Every sample then needs >= max(2, k) timesteps — for shorter ones the reference's
        targets are ragged and its torch.stack raises."""
        xs_static, xs_ts, xs_times, n_timesteps = self.feats_to_input(x, batch_size)
        V = (xs_ts.shape[2] - 1) // 2
        B = xs_ts.shape[0]
        k = int(self.pretrain_masked_steps)
        steps, evs = [], []
        for n in n_timesteps:
            if k > 1:
                if n < max(2, k):
                    raise ValueError(f"pretrain_masked_steps={k} needs at least {max(2, k)} timesteps in every sample "
                                     f"(got {n}): the reference's ragged targets cannot be stacked")
                steps.append([int(i) for i in self.rng.choice(np.arange(n), size=k)])
            else:
                steps.append(n if n < 2 else int(self.rng.choice(np.arange(0, n))))
            if self.predict_events:
                evs.append(int(self.rng.choice(np.arange(0, self.d_time_series_num))))
        dev = xs_ts.device
        if k == 1 and any(st >= xs_ts.shape[1] for st in steps):
            raise IndexError("pretrain_prep_batch: a sample has fewer than 2 timesteps")      # the reference's xs_ts[i][step] raises
        # the draws go up as ONE small pinned index block ([2,B] int32 + [B,V] uint8) and one kernel does the selection
        # (duett/duett.py:198-233: targets, masked copy, event column, variable dropout) — SURVEY §8f-2
        keep = None
        if self.pretrain_dropout > 0:
            keep = np.ascontiguousarray(self.rng.random((batch_size, V)) > self.pretrain_dropout).view(np.uint8)
        # one index block: row 0 = the masked variable, rows 1..k = the masked timestep(s)
        steps_kb = [steps] if k == 1 else [list(col) for col in zip(*steps)]
        idx = torch.tensor([evs if self.predict_events else [0] * B] + steps_kb, dtype=torch.int32)
        if dev.type == "cuda":
            idx = idx.pin_memory().to(dev, non_blocking=True)
            keep_d = None if keep is None else torch.from_numpy(keep).pin_memory().to(dev, non_blocking=True)
        else:
            keep_d = None if keep is None else torch.from_numpy(keep)
        step_d = idx[1] if k == 1 else idx[1:].t().contiguous()                  # [B] or [B,k]
        x_c, y_ts, y_ts_masks, y_events, y_events_mask = ops.ssl_mask(
            xs_ts.float().contiguous(), step_d, idx[0] if self.predict_events else None, keep_d)
        if not self.predict_events:
            y_events, y_events_mask = [], []
        return (xs_static, x_c, xs_times, n_timesteps), y_ts, y_ts_masks, y_events, y_events_mask

    # ---- forward ---------------------------------------------------------------------------------------------------
    def _act_dtype(self):
        if self.precision == "bf16":
            return torch.bfloat16
        if self.precision in ("fp32", "tf32"):
            return torch.float32
        if torch.is_autocast_enabled() and torch.get_autocast_gpu_dtype() == torch.bfloat16:
            return torch.bfloat16
        return torch.float32

    def _spec_and_params(self, at):
        names, params = [], []

        def add(n, p):
            names.append(n); params.append(p)

        add("special_embeddings.weight", self.special_embeddings.weight)
        e = self.embedding_layers
        for n in ("w0", "b0", "bn_w", "bn_b", "w4", "b4"):
            add("emb." + n, getattr(e, n))
        add("n_obs_embedding.weight", self.n_obs_embedding.weight)
        add("full_event_embedding.weight", self.full_event_embedding.weight)
        for kind, lst in (("event", self.event_transformers), ("time", self.time_transformers)):
            for l, enc in enumerate(lst):
                for k in ENC_KEYS:
                    add(f"{kind}_transformers.{l}.{k}", getattr(enc, k))
        spec = dict(names=names, d=self.d_embedding, V=self.d_time_series_num, T=None, n_layers=self.n_duett_layers,
                    heads=self.n_transformer_head, act_dtype=at, training=self.training, final_norm=self.final_norm,
                    emb_rm=e.bn_rm, emb_rv=e.bn_rv, dropout=float(self.transformer_dropout), tf32=self.precision == "tf32")
        return spec, params

    def encode(self, x):
        """DuettFeatureExtractor.encode (models/main_architecture_duett.py:31-94) -> transformed [B,T+1,(V+1)d]."""
        xs_static, xs_feats, xs_times, _ = x
        ops.require_device(xs_feats)
        at = self._act_dtype()
        B, T, _ = xs_feats.shape
        with torch.autocast("cuda", enabled=False):
            xs_feats = xs_feats.float().contiguous()
            tab = self.tab_encoder(xs_static.float().contiguous())                       # [B,d] f32
            fe = self.full_time_embedding
            hid = linear(xs_times.float().reshape(B * T, 1).contiguous(), fe[0].weight, fe[0].bias, ops.ACT_TANH)
            hid = fe[2](hid)
            te = TimeEmbedAssembleFn.apply(hid, fe[3].weight, fe[3].bias, self.full_rep_embedding.weight, B, T, at)
            spec, params = self._spec_and_params(at)
            spec["T"] = T
            if T + 1 != self.event_transformers[0].dim // self.d_embedding:
                raise ValueError(f"batch has {T} timesteps but the model was built for "
                                 f"{self.event_transformers[0].dim // self.d_embedding - 1} (masked_transform_timesteps)")
            out = DuettEncodeFn.apply(spec, xs_feats, tab, te, *params)
            if self.training:
                self.embedding_layers.bn_nbt += 1
        return out

    def _row(self, transformed, idx):
        """transformed[b, idx[b], :] as f32 (idx: python int or LongTensor [B])."""
        B, T1, Ep = transformed.shape
        if isinstance(idx, int):
            idx = torch.full((B,), idx % T1, device=transformed.device, dtype=torch.int64)
        off = (torch.arange(B, device=transformed.device, dtype=torch.int64) * T1 + idx) * Ep
        return GatherVecFn.apply(transformed, off, Ep)

    def _masked_rows(self, transformed, xs_feats):
        """pretrain_masked_steps = k > 1 (duett/duett.py:287-293): per sample the rows of the DISTINCT masked timesteps in time
        order, zero-padded to k rows -> [B*k, E'] f32.  Index glue on the [B,T] flag matrix; the gather kernel writes a zero
        row for a negative offset (and its backward scatters nothing there)."""
        B, T1, Ep = transformed.shape
        k = int(self.pretrain_masked_steps)
        flag = xs_feats[:, :, -1] > 0                                             # [B,T]
        order = torch.argsort(flag.to(torch.int8), dim=1, descending=True, stable=True)[:, :k]   # masked steps first, in time order
        valid = torch.arange(k, device=flag.device)[None, :] < flag.sum(1, keepdim=True)
        ar = torch.arange(B, device=flag.device, dtype=torch.int64)[:, None]
        off = torch.where(valid, (ar * T1 + order) * Ep, torch.full_like(order, -1))
        return GatherVecFn.apply(transformed, off.reshape(-1).contiguous(), Ep)

    def forward(self, x, pretrain=False, representation=False):
        from ..functional import MeanRowsFn
        xs_static, xs_feats, xs_times, n_timesteps = x
        transformed = self.encode(x)
        B, T1, Ep = transformed.shape
        T, V, d = T1 - 1, self.d_time_series_num, self.d_embedding
        with torch.autocast("cuda", enabled=False):
            if self.fusion_method == 'rep_token':
                z = self._row(transformed, T)
            elif self.fusion_method == 'masked_embed' and self.pretrain_masked_steps > 1:
                z = self._masked_rows(transformed, xs_feats)            # [B*k,E'], zero rows where a draw repeated
            elif self.fusion_method == 'masked_embed':
                step = (xs_feats[:, :, -1] == 1).float().argmax(1)      # index glue on a [B,T] flag matrix
                z = self._row(transformed, step)
            elif self.fusion_method == 'averaging':
                z = MeanRowsFn.apply(transformed, T)
            else:
                raise ValueError(self.fusion_method)
            ks = self.pretrain_masked_steps if self.fusion_method == 'masked_embed' else 1
            if representation:
                return z.view(B, ks, -1) if ks > 1 else z
            if pretrain:
                y_hat_presence = self.pretrain_presence_proj(z).squeeze() if self.pretrain_presence else None
                y_hat_value = self.pretrain_value_proj(z).squeeze(1) if self.pretrain_value else None
                if ks > 1:       # the heads ran on all B*k rows (their BatchNorm sees the zero-padded rows, like the reference's)
                    y_hat_presence = None if y_hat_presence is None else y_hat_presence.reshape(B, ks, -1)
                    y_hat_value = None if y_hat_value is None else y_hat_value.reshape(B, ks, -1)
                y_hat_events, y_hat_events_presence = None, None
                if self.predict_events:
                    var = (xs_feats[:, 0, V:2 * V] == -1).float().argmax(1)             # masked variable per sample
                    ar = torch.arange(B, device=z.device, dtype=torch.int64)
                    off = ((ar[:, None] * T1 + torch.arange(T1, device=z.device)[None, :]) * (V + 1) + var[:, None]) * d
                    z_events = GatherVecFn.apply(transformed, off.reshape(-1), d).reshape(B, T1 * d)
                    y_hat_events = self.predict_events_proj(z_events).squeeze()
                    y_hat_events_presence = self.predict_events_presence_proj(z_events).squeeze() \
                        if self.pretrain_presence else None
                return y_hat_value, y_hat_presence, y_hat_events, y_hat_events_presence
            if ks > 1:
                raise NotImplementedError("the supervised head on pretrain_masked_steps > 1 'masked_embed' features: the "
                                          "reference's [B,k,E'] head input has no defined meaning (fine-tuning uses rep_token)")
            out = self.head(z).squeeze(1)
        if self.save_representation:
            return out, z
        return out

    # ---- Lightning-style steps (duett/duett.py:329-457) ----------------------------------------------------------------
    @staticmethod
    def _value_presence_loss(y_hat, p_hat, y, mask, w):
        """mse(y_hat*m, y*m) + w * bce_with_logits(p_hat, m) with either head absent (pretrain_value / pretrain_presence
        off, duett/duett.py:336-357).  One kernel computes both terms; an absent head is fed zeros so that its term is
        exactly zero (zero prediction against a zero target, or presence weight 0)."""
        if y_hat is None and p_hat is None:
            return torch.zeros((), device=mask.device)
        if y_hat is None:
            z = torch.zeros_like(mask, dtype=torch.float32)
            return MaskedMseBceFn.apply(z, p_hat, z, mask, w)
        if p_hat is None:
            return MaskedMseBceFn.apply(y_hat, torch.zeros_like(mask, dtype=torch.float32), y, mask, 0.0)
        return MaskedMseBceFn.apply(y_hat, p_hat, y, mask, w)

    def _ssl_loss(self, outs, y, mask, y_events, y_events_mask):
        y_hat_value, y_hat_presence, y_hat_events, y_hat_events_presence = outs
        w = self.pretrain_presence_weight
        if y.dim() == 3:
            # pretrain_masked_steps = k > 1: the reference averages the k per-step losses (duett/duett.py:338-349); every step
            # has B*V terms, so that is the mean over all [B,k,V] terms = the same kernel on B*k rows
            if any(t is not None and t.shape != y.shape for t in (y_hat_value, y_hat_presence)):
                raise ValueError("pretrain_masked_steps > 1 needs fusion_method='masked_embed' (the reference's per-step "
                                 "y_hat[:, i] indexing has no meaning for a [B,V] prediction)")
            V = y.shape[-1]
            y, mask = y.reshape(-1, V), mask.reshape(-1, V)
            y_hat_value = None if y_hat_value is None else y_hat_value.reshape(-1, V)
            y_hat_presence = None if y_hat_presence is None else y_hat_presence.reshape(-1, V)
        loss = self._value_presence_loss(y_hat_value, y_hat_presence, y, mask, w)
        if self.predict_events:
            # the event-value prediction is always computed (duett/duett.py:314) but only scored with pretrain_value (:351)
            loss = loss + self._value_presence_loss(y_hat_events if self.pretrain_value else None, y_hat_events_presence,
                                                    y_events, y_events_mask, w)
        return loss

    def _supervised_loss(self, y_hat, y):
        if self.pos_frac is not None:
            loss = BCELogitsFn.apply(y_hat, y.float(), float(self.pos_weight), float(self.neg_weight))
        else:
            loss = BCELogitsFn.apply(y_hat, y.float(), 1.0, 1.0)
        return loss.double()       # the reference's labels are float64, so its loss is float64 (duett/duett.py:331)

    def training_step(self, batch, batch_idx):
        x, y = batch
        y = torch.tensor(y, dtype=torch.float64, device=self.device)
        batch_size = y.shape[0]
        if self.pretrain:
            x_pretrain, y, mask, y_events, y_events_mask = self.pretrain_prep_batch(x, batch_size)
            outs = self.forward(x_pretrain, pretrain=True)
            loss = self._ssl_loss(outs, y, mask, y_events, y_events_mask)
        else:
            y_hat = self.forward(self.feats_to_input(x, batch_size))
            loss = self._supervised_loss(y_hat, y)
            self.train_auroc.update(y_hat, y.to(int))
            self.train_ap.update(y_hat, y.to(int))
        self.log('train_loss', loss, sync_dist=True)
        return loss

    def validation_step(self, batch, batch_idx):
        x, y = batch
        y = torch.tensor(y, dtype=torch.float64, device=self.device)
        batch_size = y.shape[0]
        if self.pretrain:
            x_pretrain, y, mask, y_events, y_events_mask = self.pretrain_prep_batch(x, batch_size)
            outs = self.forward(x_pretrain, pretrain=True)
            loss = self._ssl_loss(outs, y, mask, y_events, y_events_mask)
        else:
            y_hat = self.forward(self.feats_to_input(x, batch_size))
            loss = self._supervised_loss(y_hat, y)
            self.val_auroc.update(y_hat, y.to(int))
            self.val_auprc.update(y_hat, y.to(int))
        self.log('val_loss', loss, on_epoch=True, sync_dist=True, prog_bar=True, rank_zero_only=True)
        return loss

    def test_step(self, batch, batch_idx):
        x, y = batch
        y = torch.tensor(y, dtype=torch.float64, device=self.device)
        batch_size = y.shape[0]
        out = self.forward(self.feats_to_input(x, batch_size))
        y_hat = out[0] if self.save_representation else out
        loss = self._supervised_loss(y_hat, y)
        self.test_auroc.update(y_hat, y.to(int))
        self.test_auprc.update(y_hat, y.to(int))
        self.log('test_loss', loss, on_epoch=True, sync_dist=True, rank_zero_only=True)
        return loss, self.test_auroc, self.test_auprc

    def on_train_epoch_end(self):
        if not self.pretrain:                  # duett/duett.py:420-423
            self.log('train_auroc', self.train_auroc, sync_dist=True, rank_zero_only=True)
            self.log('train_ap', self.train_ap, sync_dist=True, rank_zero_only=True)

    def on_validation_epoch_end(self):
        if not self.pretrain:
            print(f'[epoch {self.current_epoch:>3d}] val_auroc={self.val_auroc.compute():.4f}  '
                  f'val_auprc={self.val_auprc.compute():.4f}')
