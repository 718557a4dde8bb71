"""ORACLE / TEST INFRASTRUCTURE ONLY — generates tests/golden/g6_teacher_attn.npz by running the REFERENCE'S OWN
TeacherModel.forward(..., return_attn=True) (models/main_architecture_duett.py:1075-1129 -> PatchDualPathologyPerceiver
.forward :595-654 -> _PerceiverBlock.forward :759-774, nn.MultiheadAttention(need_weights=True, average_attn_weights=True))
in eval mode — the inference-time visualisation call of analysis/visualize_pathology.py.

The teacher is the one of fixture g4 (same parameters, same inputs; oracle/make_golden.py), so the fixture stores only the
extra outputs: the head-averaged attention maps of the two cross-attention blocks, the latent tokens and the eval-mode logits.

Run in the authoring container only (needs /root/reference, read-only):   python oracle/make_golden_attn.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DUETT_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "shims"), REF, ROOT, os.path.join(ROOT, "tests")]


def main():
    torch.set_num_threads(4)
    from golden_util import load
    from models.main_architecture_duett import DuettFeatureExtractor, PatchDualPathologyPerceiver, TeacherModel
    torch.set_float32_matmul_precision("highest")
    G = load("g4_teacher")
    kw = dict(d_static_num=3, d_time_series_num=5, d_target=1, d_embedding=8, masked_transform_timesteps=4, max_len=4,
              n_duett_layers=2, d_feedforward=96)
    d_img = 16

    class StubCXR(torch.nn.Module):      # CXR embeddings ride in the pixel_values slot
        d_out = d_img

        def forward(self, pv):
            return pv[:, 0], pv[:, 1:]

    torch.manual_seed(3)
    duett = DuettFeatureExtractor(pretrain=False, **kw)
    perceiver = PatchDualPathologyPerceiver(7, duett.d_representation, d_latent=32, n_heads=4, dropout=0.0, head_hidden=16,
                                            head_dropout=0.0)
    teacher = TeacherModel(duett, StubCXR(), perceiver, patch_dual_pathology_mode=True, d_img=d_img)
    teacher.load_state_dict(G["param"], strict=True)
    teacher.eval()
    I = G["in"]
    with torch.no_grad():
        out = teacher(list(I["x_ts"]), list(I["x_static"]), list(I["bin_ends"]), I["pixel_values"], return_attn=True)
    blobs = {"out/" + k: out[k].detach().cpu().numpy() for k in
             ("main_logit", "img_logits", "ts_logits", "fusion_logits", "ts_correction", "scaled_correction", "img_tokens",
              "ts_tokens", "fusion_tokens", "img_attn", "ts_attn")}
    assert blobs["out/img_attn"].shape == (6, 7, 10) and blobs["out/ts_attn"].shape == (6, 7, 4)
    assert np.allclose(blobs["out/img_attn"].sum(-1), 1.0, atol=1e-5)
    path = os.path.join(ROOT, "tests", "golden", "g6_teacher_attn.npz")
    np.savez_compressed(path, **blobs)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB, {len(blobs)} arrays)")


if __name__ == "__main__":
    main()
