"""Seeded synthetic inputs in the reference's collate format: per sample `obs ~ Bernoulli(density)` over
[T,V], z-scored values N(0,1)*obs, observation counts obs*U{1,2,3}, static features [N(0,1) age, one-hot-ish
Bernoulli(0.2)], `bin_ends = arange(1,T+1)/24` (data_processing.py:343), labels Bernoulli(0.3).  Used by bench.py and the
tools; there are no datasets in the build environment."""
from __future__ import annotations

import torch


def synth_batch(d_static_num: int, V: int, T: int, B: int, seed: int, density: float = 0.2) -> dict:
    """{"x_ts": B x [T,2V], "x_static": B x [S], "bin_ends": B x [T], "y": [B]} on the host (tuples of per-sample tensors,
    duett/mimic_dataset.py:83,93-95)."""
    g = torch.Generator().manual_seed(seed)
    obs = (torch.rand(B, T, V, generator=g) < density).float()
    vals = torch.randn(B, T, V, generator=g) * obs
    cnts = obs * torch.randint(1, 4, (B, T, V), generator=g).float()
    x_ts = torch.cat((vals, cnts), dim=2)
    x_static = torch.cat((torch.randn(B, 1, generator=g), (torch.rand(B, d_static_num - 1, generator=g) < 0.2).float()), dim=1)
    bin_ends = (torch.arange(1, T + 1).float() / 24.0)[None].expand(B, -1).contiguous()
    y = (torch.rand(B, generator=g) < 0.3).float()
    return {"x_ts": tuple(x_ts), "x_static": tuple(x_static), "bin_ends": tuple(bin_ends), "y": y}
