"""ORACLE / TEST INFRASTRUCTURE ONLY — generates tests/golden/g7_ssl_masked_steps.npz by running the REFERENCE'S OWN
Model.training_step (duett/duett.py:328-357) with pretrain_masked_steps = 2: pretrain_prep_batch draws two timesteps per
sample WITH replacement (:202, rng.choice(..., size=k)), forward gathers the distinct masked rows in time order and
zero-pads them to k rows (:287-293), the SSL heads run on all B*k rows (BatchNorm statistics include the zero-padded rows)
and the loss is the mean of the per-step losses (:337-349).  Same fixture layout as oracle/make_golden.py's G3.

Run in the authoring container only (needs /root/reference, read-only):   python oracle/make_golden_masked_steps.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DUETT_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "shims"), REF, ROOT]

from oracle import duett_oracle as O  # noqa: E402
from oracle.make_golden import grads_of, params_of, save, self_dev  # noqa: E402

K_STEPS, SEED = 2, 42


def main():
    torch.set_num_threads(4)
    from duett.duett import Model
    torch.set_float32_matmul_precision("highest")
    cfg = O.DuettConfig(d_static_num=3, d_time_series_num=5, n_timesteps=4, d_embedding=8, n_layers=2, d_feedforward=96)
    kw = dict(d_static_num=cfg.d_static_num, d_time_series_num=cfg.V, d_target=1, d_embedding=cfg.d_embedding,
              masked_transform_timesteps=cfg.T, max_len=cfg.T, n_duett_layers=cfg.n_layers, d_feedforward=cfg.d_feedforward)
    torch.manual_seed(4)
    model = Model(pretrain=True, seed=SEED, pretrain_masked_steps=K_STEPS, **kw)
    model.train()
    batch = O.synth_batch(cfg, B=6, seed=1238, density=0.5)
    blobs = params_of(model)
    x = (batch["x_ts"], batch["x_static"], list(batch["bin_ends"]))
    x_pre, y, mask, y_events, y_events_mask = model.pretrain_prep_batch(x, 6)
    n_masked = (x_pre[1][:, :, -1] > 0).sum(1)
    assert n_masked.min() < K_STEPS <= n_masked.max(), "want both a duplicated draw (zero-padded row) and two distinct steps"
    model.rng = np.random.default_rng(SEED)   # rewind so training_step draws the same masks
    outs = model.forward(tuple(t.clone() if torch.is_tensor(t) else t for t in x_pre), pretrain=True)
    model.load_state_dict({k[len("param/"):]: v for k, v in blobs.items()})
    loss = model.training_step((x, tuple(batch["y"].tolist())), 0)
    model.zero_grad()
    loss.backward()
    blobs.update(grads_of(model))
    blobs.update({"in/x_ts": torch.stack(batch["x_ts"]), "in/x_static": torch.stack(batch["x_static"]),
                  "in/bin_ends": torch.stack(batch["bin_ends"]), "out/xs_ts_clipped": x_pre[1], "out/y": y,
                  "out/mask": mask, "out/y_events": y_events, "out/y_events_mask": y_events_mask,
                  "out/y_hat_value": outs[0], "out/y_hat_presence": outs[1], "out/y_hat_events": outs[2],
                  "out/y_hat_events_presence": outs[3], "out/loss": loss, "out/n_masked": n_masked})
    def _ssl_step():
        model.rng = np.random.default_rng(SEED)
        return model.training_step((x, tuple(batch["y"].tolist())), 0)
    blobs.update(self_dev(model, blobs, _ssl_step))
    save("g7_ssl_masked_steps", blobs)


if __name__ == "__main__":
    main()
