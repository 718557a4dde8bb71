"""argtypes/restype declarations for the non-GEMM entry points of libduett_b200.so (include/duett_b200.h)."""
from __future__ import annotations

import ctypes as C

P, I, L, F, U64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64

SIGNATURES = {
    # name: argtypes (the trailing P is always the cudaStream_t)
    "dx_relayout_fwd": [P, P, P, P, P, P, P, I, I, I, I, I, P],
    "dx_relayout_bwd": [P, P, P, P, P, P, I, I, I, I, I, P],
    "dx_colsum": [P, L, I, L, P, I, I, P],
    "dx_axpy": [P, P, L, F, I, I, P],
    "dx_cast": [P, I, P, I, L, P],
    "dx_attn_fwd": [P, L, L, P, L, L, P, L, L, P, L, L, P, I, I, I, I, I, I, F, U64, P, P],
    "dx_attn_bwd": [P, L, L, P, L, L, P, L, L, P, L, L, P, L, L, P, L, L, P, L, L, P, L, L, P, P, I, I, I, I, I, I, F, U64, P, P],
    "dx_attn_probs_mean": [P, L, L, P, L, L, P, P, I, I, I, I, I, I, P],
    "dx_dropout": [P, P, L, F, U64, P, I, P],
    "dx_rowdot_bias": [P, P, P, P, I, I, I, P],
    "dx_rowdot_scale": [P, P, P, P, I, I, I, P],
    "dx_scalenorm_scale": [P, P, F, P, I, P],
    "dx_embed_stats": [P, I, I, I, P, P, P, P, P, P, P, P, I, P],
    "dx_embed_hidden": [P, I, I, I, P, P, P, P, P, P, P, P, I, P],
    "dx_embed_special": [P, I, I, I, I, P, P, P, I, P],
    "dx_embed_special_bwd": [P, I, I, I, I, P, I, P, P, P],
    "dx_embed_bn_reduce": [P, I, I, I, P, P, P, P, P, P, I, P, P],
    "dx_embed_bwd_front": [P, I, I, I, P, P, P, P, P, P, P, I, P, P, P, P, I, P],
    "dx_bn2d_fwd": [P, I, I, P, P, P, P, P, P, P, P, I, P],
    "dx_bn2d_bwd": [P, P, I, I, P, P, P, P, P, P, P, I, P],
    "dx_layernorm_fwd": [P, I, I, P, P, P, P, P, I, P],
    "dx_layernorm_bwd": [P, P, I, I, P, P, P, P, P, P, I, P],
    "dx_kd_loss": [P, P, P, I, F, F, F, F, P, P, P],
    "dx_bce_logits": [P, P, I, F, F, P, P, P],
    "dx_masked_mse_bce": [P, P, P, P, I, F, P, P, P, P],
    "dx_masked_bce_cols": [P, P, P, P, P, I, I, F, P, P, P],
    "dx_aux_residual_kl": [P, P, P, P, I, F, P, P, P],
    "dx_mean_rows": [P, P, I, I, I, L, I, P],
    "dx_mean_rows_bwd": [P, P, I, I, I, L, I, P],
    "dx_gather_vec": [P, P, P, I, I, I, P],
    "dx_scatter_vec": [P, P, P, I, I, I, I, P],
    "dx_adamw": [P, P, P, P, L, F, F, F, F, F, I, P, F, P, P, P, P],
    "dx_sumsq": [P, L, P, P],
    "dx_clip_factor": [P, F, P, P],
    "dx_sum_n": [P, I, P, L, I, P],
    "dx_bin_events": [P, P, P, P, P, P, I, I, I, P, P],
    "dx_ssl_mask": [P, P, P, P, I, I, I, I, P, P, P, P, P, P],
    "dx_binary_auc": [P, P, L, P, P, L, I, P, P],
    "dx_act_bwd": [P, P, P, L, I, I, P],
    "dx_act_fwd": [P, P, L, I, I, P],
    "dx_fusion_logits": [P, P, P, P, P, P, P, P, P, P, I, I, P],
    "dx_fusion_logits_bwd": [P, P, P, P, P, P, P, P, P, P, I, I, P],
    "dx_scale_dev": [P, P, P, L, I, P],
    "dx_sum_div_acc": [P, L, P, P, P],
}


def declare(lib: C.CDLL) -> None:
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)       # raises AttributeError if the library does not export the symbol
        fn.argtypes = argtypes
        fn.restype = C.c_int
