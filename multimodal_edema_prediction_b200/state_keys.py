"""State-dict key mapping between this package's fused parameter storage and the reference's checkpoint layout.

Internally the V per-variable embedding MLPs are stacked tensors and q/k/v projections are one fused [3d, dim]
matrix.  Checkpoints keep the reference's keys (probed from duett/duett.py + x_transformers):

    embedding_layers.{i}.0.{weight,bias}                    <- embedding_layers.w0[i], .b0[i]
    embedding_layers.{i}.3.batch_norm.{weight,bias,running_mean,running_var,num_batches_tracked}
    embedding_layers.{i}.4.{weight,bias}                    <- embedding_layers.w4[i], .b4[i]
    {event,time}_transformers.{l}.layers.0.0.0.g            <- g_attn
    {event,time}_transformers.{l}.layers.0.1.to_{q,k,v}.weight <- wqkv[0:d], [d:2d], [2d:3d]
    {event,time}_transformers.{l}.layers.0.1.to_out.weight  <- wo
    {event,time}_transformers.{l}.layers.1.0.0.g            <- g_ff
    {event,time}_transformers.{l}.layers.1.1.ff.0.0.{weight,bias} <- w1, b1
    {event,time}_transformers.{l}.layers.1.1.ff.2.{weight,bias}   <- w2, b2
    {event,time}_transformers.{l}.final_norm.g              <- g_final

The x_transformers part of the table is the library's 1.x/2.x layout (unpinned by the reference — SURVEY §8c); it lives
in ENC_MAP so another vintage only needs a different table.
"""
from __future__ import annotations

import re

import torch

ENC_MAP = {
    "g_attn": "layers.0.0.0.g",
    "wo": "layers.0.1.to_out.weight",
    "g_ff": "layers.1.0.0.g",
    "w1": "layers.1.1.ff.0.0.weight",
    "b1": "layers.1.1.ff.0.0.bias",
    "w2": "layers.1.1.ff.2.weight",
    "b2": "layers.1.1.ff.2.bias",
    "g_final": "final_norm.g",
}
QKV = ("layers.0.1.to_q.weight", "layers.0.1.to_k.weight", "layers.0.1.to_v.weight")
EMB_MAP = {
    "w0": "0.weight", "b0": "0.bias", "bn_w": "3.batch_norm.weight", "bn_b": "3.batch_norm.bias",
    "bn_rm": "3.batch_norm.running_mean", "bn_rv": "3.batch_norm.running_var",
    "bn_nbt": "3.batch_norm.num_batches_tracked", "w4": "4.weight", "b4": "4.bias",
}
_ENC_RE = re.compile(r"^(event|time)_transformers\.(\d+)\.(\w+)$")


def to_reference(sd: dict, prefix: str = "") -> None:
    """In-place: internal keys -> reference keys (state_dict post-hook)."""
    for k in [k for k in sd if k.startswith(prefix)]:
        rel = k[len(prefix):]
        if rel.startswith("embedding_layers.") and rel.split(".", 1)[1] in EMB_MAP:
            t = sd.pop(k)
            suffix = EMB_MAP[rel.split(".", 1)[1]]
            for i in range(t.shape[0]):
                sd[f"{prefix}embedding_layers.{i}.{suffix}"] = t[i]
            continue
        m = _ENC_RE.match(rel)
        if m:
            kind, l, name = m.groups()
            base = f"{prefix}{kind}_transformers.{l}."
            t = sd.pop(k)
            if name == "wqkv":
                d = t.shape[0] // 3
                for j, q in enumerate(QKV):
                    sd[base + q] = t[j * d:(j + 1) * d]
            elif name in ENC_MAP:
                sd[base + ENC_MAP[name]] = t
            else:
                sd[k] = t


def from_reference(sd: dict, prefix: str = "") -> None:
    """In-place: reference keys -> internal keys (load_state_dict pre-hook). Internal keys pass through untouched."""
    emb: dict = {}
    qkv: dict = {}
    inv_enc = {v: k for k, v in ENC_MAP.items()}
    inv_emb = {v: k for k, v in EMB_MAP.items()}
    for k in [k for k in sd if k.startswith(prefix)]:
        rel = k[len(prefix):]
        m = re.match(r"^embedding_layers\.(\d+)\.(.+)$", rel)
        if m and m.group(2) in inv_emb:
            emb.setdefault(inv_emb[m.group(2)], {})[int(m.group(1))] = sd.pop(k)
            continue
        m = re.match(r"^(event|time)_transformers\.(\d+)\.(.+)$", rel)
        if m:
            kind, l, tail = m.groups()
            base = f"{prefix}{kind}_transformers.{l}."
            if tail in QKV:
                qkv.setdefault(base, {})[QKV.index(tail)] = sd.pop(k)
            elif tail in inv_enc:
                sd[base + inv_enc[tail]] = sd.pop(k)
    for name, parts in emb.items():
        sd[f"{prefix}embedding_layers.{name}"] = torch.stack([parts[i] for i in range(len(parts))])
    for base, parts in qkv.items():
        if len(parts) == 3:
            sd[base + "wqkv"] = torch.cat([parts[0], parts[1], parts[2]], dim=0)
