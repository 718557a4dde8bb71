"""B200-native drop-in for the reference's models/main_architecture_duett.py (live classes only):

    DuettFeatureExtractor, load_duett_backbone, PatchDualPathologyPerceiver, _PerceiverBlock, TeacherModel, StudentModel

Same constructor signatures, attribute names (`.duett .cxr .perceiver .img_proj .head`), return types and state-dict
keys as the reference (models/main_architecture_duett.py:26-123, 536-654, 745-774, 993-1235).  CXREncoder (frozen
RAD-DINO ViT) and LocalTrajectoryEncoder are outside the hot path: TeacherModel accepts any `cxr_encoder`
module returning `(cls [B,768], patches [B,N,768])`.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from ..duett.duett import DxSequential, Model as DuettBase
from ..functional import AttentionFn, FusionLogitsFn, MeanRowsFn, cast, dropout, layer_norm, linear


class DuettFeatureExtractor(DuettBase):
    @property
    def d_representation(self) -> int:
        return self.d_embedding * (self.d_time_series_num + 1)
    # encode(x) is inherited: DuettBase.encode IS the fused backbone (one implementation serves both classes)


def load_duett_backbone(ckpt_path: str, d_static_num: int, d_time_series_num: int, n_timesteps: int,
                        freeze: bool = False, aug_noise: float = 0.0, aug_mask: float = 0.0,
                        transformer_dropout: float = 0.0, **model_kwargs) -> DuettFeatureExtractor:
    """models/main_architecture_duett.py:98-123.  The reference's checkpoints carry no hyper-parameters, so anything
    beyond the defaults (d_embedding, n_duett_layers, ...) must be supplied through **model_kwargs."""
    model = DuettFeatureExtractor.load_from_checkpoint(
        ckpt_path, pretrain=False, d_static_num=d_static_num, d_time_series_num=d_time_series_num, d_target=1,
        masked_transform_timesteps=n_timesteps, max_len=n_timesteps, aug_noise=aug_noise, aug_mask=aug_mask,
        transformer_dropout=transformer_dropout, strict=False, **model_kwargs)
    if freeze:
        for p in model.parameters():
            p.requires_grad = False
        model.eval()
    return model


class _MHA(nn.Module):
    """Parameter layout of nn.MultiheadAttention(d, n_heads, batch_first=True): in_proj_weight [3d,d], in_proj_bias [3d],
    out_proj.{weight,bias}; forward = projection GEMMs + fused softmax-attention kernel."""

    def __init__(self, d, n_heads, dropout):
        super().__init__()
        self.embed_dim, self.num_heads, self.dropout = d, n_heads, dropout
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = nn.Linear(d, d)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)

    def forward(self, q_in, kv_in, residual=None, same_kv=False, need_weights=False):
        """need_weights=True also returns the head-averaged attention probabilities [B,Sq,Sk] f32
        (nn.MultiheadAttention(need_weights=True, average_attn_weights=True))."""
        d = self.embed_dim
        W, b = self.in_proj_weight, self.in_proj_bias
        if same_kv and q_in is kv_in:
            qkv = linear(q_in, W, b)
            q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
        else:
            q = linear(q_in, W[:d], b[:d])
            kv = linear(kv_in, W[d:], b[d:])
            k, v = kv[..., :d], kv[..., d:]
        drop_p = float(self.dropout) if self.training else 0.0
        if not need_weights:
            o = AttentionFn.apply(q, k, v, self.num_heads, drop_p)
            return linear(o, self.out_proj.weight, self.out_proj.bias, res=residual)
        if drop_p > 0:
            # torch returns the *dropped* probabilities here; the reference asks for maps at inference only
            raise NotImplementedError("attention maps with attention dropout active: call .eval() first "
                                      "(return_attn=True is the reference's inference-time visualisation path)")
        o, lse = AttentionFn.apply(q, k, v, self.num_heads, 0.0, True)
        with torch.no_grad():
            attn_w = ops.attn_probs_mean(q.detach(), k.detach(), lse, self.num_heads)
        return linear(o, self.out_proj.weight, self.out_proj.bias, res=residual), attn_w


class _PerceiverBlock(nn.Module):
    """models/main_architecture_duett.py:745-774: LN(q), LN(kv) -> MHA -> +res -> LN -> FF(d,4d,GELU) -> +res."""
    _DEBUG_NORMS: bool = False

    def __init__(self, d: int, n_heads: int, dropout: float):
        super().__init__()
        self.norm_q = nn.LayerNorm(d)
        self.norm_kv = nn.LayerNorm(d)
        self.attn = _MHA(d, n_heads, dropout)
        self.norm_ff = nn.LayerNorm(d)
        self.ff = nn.Sequential(nn.Linear(d, d * 4), nn.GELU(), nn.Dropout(dropout), nn.Linear(d * 4, d),
                                nn.Dropout(dropout))
        self.dropout = dropout

    def forward(self, latents, kv, return_attn: bool = False):
        q = layer_norm(latents, self.norm_q.weight, self.norm_q.bias)
        k = layer_norm(kv, self.norm_kv.weight, self.norm_kv.bias)
        attn_w = None
        if return_attn:
            latents, attn_w = self.attn(q, k, residual=latents, need_weights=True)
        else:
            latents = self.attn(q, k, residual=latents)
        f = layer_norm(latents, self.norm_ff.weight, self.norm_ff.bias)
        h = linear(f, self.ff[0].weight, self.ff[0].bias, ops.ACT_GELU)
        h = dropout(h, self.ff[2].p, self.training, "perceiver.ff")
        if self.ff[4].p > 0 and self.training:      # Dropout after the second Linear sits before the residual add
            latents = latents + dropout(linear(h, self.ff[3].weight, self.ff[3].bias), self.ff[4].p, True, "perceiver.ff_out")
        else:
            latents = linear(h, self.ff[3].weight, self.ff[3].bias, res=latents)
        return (latents, attn_w) if return_attn else latents


class PatchDualPathologyPerceiver(nn.Module):
    """models/main_architecture_duett.py:536-654."""

    def __init__(self, n_pathologies: int, d_ts: int, d_latent: int = 256, n_heads: int = 4, dropout: float = 0.1,
                 head_hidden: int = 64, head_dropout: float = 0.1):
        super().__init__()
        self.n_pathologies, self.d_latent, self.d_ts = n_pathologies, d_latent, d_ts
        self.shared_queries = nn.Parameter(torch.randn(n_pathologies, d_latent) * 0.02)
        self.ts_proj = nn.Linear(d_ts, d_latent)
        self.img_cross = _PerceiverBlock(d_latent, n_heads, dropout)
        self.img_self = _PerceiverBlock(d_latent, n_heads, dropout)
        self.ts_cross = _PerceiverBlock(d_latent, n_heads, dropout)
        self.ts_self = _PerceiverBlock(d_latent, n_heads, dropout)

        def _mk_head():
            return DxSequential(nn.Linear(d_latent, head_hidden), nn.GELU(), nn.Dropout(head_dropout),
                                nn.Linear(head_hidden, 1))

        self.image_head = _mk_head()
        self.temporal_head = _mk_head()
        self.correction_head = nn.Sequential(nn.LayerNorm(d_latent), nn.Linear(d_latent, head_hidden), nn.GELU(),
                                             nn.Dropout(head_dropout), nn.Linear(head_hidden, 1, bias=False))
        nn.init.zeros_(self.correction_head[-1].weight)
        self.head_dropout = head_dropout
        self.beta = nn.Parameter(torch.ones(n_pathologies))
        self.image_label_bias = nn.Parameter(torch.zeros(n_pathologies))
        self.temporal_label_bias = nn.Parameter(torch.zeros(n_pathologies))

    def forward(self, ts_tokens, img_patches_proj, return_attn=False, ts_ablation="hourly_only"):
        if ts_tokens.ndim != 3:
            raise ValueError(f"ts_tokens must be [B, T+1, d_ts], got {tuple(ts_tokens.shape)}")
        if ts_ablation not in ("full", "hourly_only", "rep_only"):
            raise ValueError(f"unknown ts_ablation={ts_ablation!r}; expected one of "
                             "{'full', 'hourly_only', 'rep_only'}")
        B = ts_tokens.size(0)
        at = ts_tokens.dtype
        with torch.autocast("cuda", enabled=False):
            q0 = cast(self.shared_queries, at).unsqueeze(0).expand(B, -1, -1).contiguous()
            ts_all = linear(ts_tokens, self.ts_proj.weight, self.ts_proj.bias)          # [B,T+1,d_latent]
            if ts_ablation == "full":
                ts_kv = ts_all
            elif ts_ablation == "hourly_only":
                ts_kv = ts_all[:, :-1]
            else:
                ts_kv = ts_all[:, -1:]
            ts_kv = ts_kv.contiguous()
            img_kv = img_patches_proj if img_patches_proj.dtype == at else cast(img_patches_proj, at)
            img_attn = ts_attn = None
            if return_attn:        # [B,K,N_patches] / [B,K,T or T+1 or 1] f32, head-averaged
                I, img_attn = self.img_cross(q0, img_kv, return_attn=True)
            else:
                I = self.img_cross(q0, img_kv)
            I = self.img_self(I, I)
            if return_attn:
                T_tok, ts_attn = self.ts_cross(q0, ts_kv, return_attn=True)
            else:
                T_tok = self.ts_cross(q0, ts_kv)
            T_tok = self.ts_self(T_tok, T_tok)
            If, Tf = cast(I, torch.float32), cast(T_tok, torch.float32)
            hi = self.image_head(If).squeeze(-1)
            ht = self.temporal_head(Tf).squeeze(-1)
            ch = self.correction_head
            c = layer_norm(Tf, ch[0].weight, ch[0].bias)
            c = linear(c, ch[1].weight, ch[1].bias, ops.ACT_GELU)
            c = dropout(c, ch[3].p, self.training, "correction_head")
            ts_correction = linear(c, ch[4].weight, None).squeeze(-1)
            img_logits, ts_logits, scaled_correction, fusion_logits = FusionLogitsFn.apply(
                hi, ht, ts_correction, self.image_label_bias, self.temporal_label_bias, self.beta)
        out = {"img_logits": img_logits, "ts_logits": ts_logits, "fusion_logits": fusion_logits, "img_tokens": I,
               "ts_tokens": T_tok, "fusion_tokens": T_tok, "ts_correction": ts_correction,
               "scaled_correction": scaled_correction}
        if return_attn:
            out["img_attn"], out["ts_attn"] = img_attn, ts_attn
        return out


class TeacherModel(nn.Module):
    """models/main_architecture_duett.py:993-1197 — the live `patch_dual_pathology_mode` branch.  The legacy branches
    (TemporalPerceiver / PathologyPerceiver / DualPathologyPerceiver) are commented out or unreachable in the reference
    snapshot and raise here."""

    def __init__(self, duett_backbone: DuettFeatureExtractor, cxr_encoder, perceiver, head_hidden: int = 128,
                 head_dropout: float = 0.1, cxr_return_patches: bool = True, d_img: int = 768, use_aux_cxr: bool = True,
                 aux_head_hidden: int = 128, pathology_mode: bool = False, dual_pathology_mode: bool = False,
                 patch_dual_pathology_mode: bool = False, pretrained_cxr_head_ckpt: Optional[str] = None,
                 pathology_labels: Optional[tuple] = None):
        super().__init__()
        n_modes = sum([pathology_mode, dual_pathology_mode, patch_dual_pathology_mode])
        if n_modes > 1:
            raise ValueError("at most one of pathology_mode / dual_pathology_mode / patch_dual_pathology_mode may be True")
        if not patch_dual_pathology_mode:
            raise NotImplementedError("only patch_dual_pathology_mode=True (the reference's live teacher) is implemented")
        self.duett, self.cxr, self.perceiver = duett_backbone, cxr_encoder, perceiver
        self.cxr_return_patches = cxr_return_patches
        self.pathology_mode, self.dual_pathology_mode = pathology_mode, dual_pathology_mode
        self.patch_dual_pathology_mode = patch_dual_pathology_mode
        self.img_proj = nn.Linear(d_img, perceiver.d_latent)
        self.head, self.aux_cxr_head, self.use_aux_cxr = None, None, False

    def forward(self, x_ts_list, x_static_list, bin_ends_list, pixel_values: torch.Tensor,
                batch_size: Optional[int] = None, return_attn: bool = False):
        if batch_size is None:
            batch_size = pixel_values.shape[0]
        duett_in = self.duett.feats_to_input((x_ts_list, x_static_list, bin_ends_list), batch_size)
        ts_tokens = self.duett.encode(duett_in)                                        # [B,T+1,E']
        img_cls, img_patches = self.cxr(pixel_values)
        at = ts_tokens.dtype
        with torch.autocast("cuda", enabled=False):
            patches = img_patches if img_patches.dtype == at else cast(img_patches.contiguous(), at)
            img_patches_proj = linear(patches, self.img_proj.weight, self.img_proj.bias)
        out = self.perceiver(ts_tokens, img_patches_proj, return_attn=return_attn)
        result = {"main_logit": out["fusion_logits"][:, 0], "img_logits": out["img_logits"],
                  "ts_logits": out["ts_logits"], "fusion_logits": out["fusion_logits"],
                  "ts_correction": out["ts_correction"], "scaled_correction": out["scaled_correction"]}
        if return_attn:                # models/main_architecture_duett.py:1123-1128
            for k in ("img_tokens", "ts_tokens", "fusion_tokens", "img_attn", "ts_attn"):
                result[k] = out[k]
        return result


class StudentModel(nn.Module):
    """DuETT(TS) + MLP head (models/main_architecture_duett.py:1202-1235)."""

    def __init__(self, duett_backbone: DuettFeatureExtractor, pool: str = "mean", head_hidden: int = 128,
                 head_dropout: float = 0.1):
        super().__init__()
        self.duett = duett_backbone
        self.pool = pool
        d_rep = duett_backbone.d_representation
        self.head = DxSequential(nn.Linear(d_rep, head_hidden), nn.GELU(), nn.Dropout(head_dropout),
                                 nn.Linear(head_hidden, 1))

    def forward(self, x_ts_list, x_static_list, bin_ends_list, batch_size: Optional[int] = None) -> torch.Tensor:
        if batch_size is None:
            batch_size = len(x_ts_list)
        duett_in = self.duett.feats_to_input((x_ts_list, x_static_list, bin_ends_list), batch_size)
        ts_tokens = self.duett.encode(duett_in)                                        # [B,T+1,d_rep]
        with torch.autocast("cuda", enabled=False):
            if self.pool == "rep_token":
                feat = self.duett._row(ts_tokens, ts_tokens.shape[1] - 1)
            elif self.pool == "mean":
                feat = MeanRowsFn.apply(ts_tokens, ts_tokens.shape[1] - 1)              # [REP] row excluded
            else:
                raise ValueError(f"unknown pool: {self.pool}")
            return self.head(feat).squeeze(-1)
