"""Input binning for the DuETT path — the tensor builders of the reference's duett/mimic_dataset.py:33-54,83-95 behind the
same names, with the python row walk replaced by one device launch per batch.

    build_stay_tensor(df_stay, means, stds, n_timesteps, all_vars, all_counts)  -> [T, 2V] f32   (mimic_dataset.py:33-46)
    build_batch_tensors([df_stay, ...], ...)                                    -> [B, T, 2V] f32 (one dx_bin_events launch)
    encode_static(row, age_mean, age_std, onehot_static)                        -> [S] f32        (mimic_dataset.py:49-53)
    collate_into_seqs(batch)                                                    (mimic_dataset.py:93-95)

The frames stay pandas objects owned by the caller (the artifacts, cohort filter, split and statistics of
`prepare_for_*` are pandas data preparation and out of scope); only their numeric columns are handed to the
kernel.  Results are bit-identical to the reference's: float64 arithmetic, one rounding to float32, later rows of a slot
overwrite earlier ones.  There is no CPU path: the tensors are built on the CUDA device and returned there.

DataLoader workers.  The reference feeds MIMICDataset through DataLoader(num_workers=8, pin_memory=True,
persistent_workers=True) (duett/train_duett_ssl.py:137, train_duett_finetune.py:143, training_duett/trainer.py:54); a forked
worker cannot touch CUDA.  So `MIMICDataset.__getitem__` is host-only: it returns the stay's raw event rows as a
`StayRows` record (numpy arrays: picklable, pinnable) in the x_ts slot, `collate_into_seqs` passes them through, and the ONE
`dx_bin_events` launch for the whole batch happens in the main process when `Model.feats_to_input` (or `bin_stay_rows`)
meets them.  The device builders below refuse to run inside a worker with a clear error.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


def _rows(df_stay, n_timesteps, all_vars, all_counts):
    """(slot int32 [R], vals f64 [R,V], cnts f64 [R,V]) of one stay frame, python-index semantics for the slot."""
    slot = df_stay["slot_idx"].to_numpy()
    slot = np.asarray([int(s) for s in slot], dtype=np.int64) if slot.dtype.kind not in "iu" else slot.astype(np.int64)
    if (slot < -n_timesteps).any():
        raise IndexError(f"index {int(slot.min())} is out of bounds for dimension 0 with size {n_timesteps}")
    slot = np.where(slot < 0, slot + n_timesteps, slot)          # x_ts[t] with a negative python index
    vals = df_stay[list(all_vars)].to_numpy(dtype=np.float64)
    cnts = df_stay[list(all_counts)].to_numpy(dtype=np.float64)  # several variables may share one count column
    return slot.astype(np.int32), np.ascontiguousarray(vals), np.ascontiguousarray(cnts)


class StayRows:
    """Host-side event rows of one stay (what build_stay_tensor reads from its frame) + the z-score statistics: the
    x_ts element a DataLoader worker hands to the main process.  `shape` mimics the [T, 2V] tensor it stands for."""
    __slots__ = ("slot", "vals", "cnts", "mu", "sd", "n_timesteps")

    def __init__(self, slot, vals, cnts, mu, sd, n_timesteps):
        self.slot, self.vals, self.cnts, self.mu, self.sd, self.n_timesteps = slot, vals, cnts, mu, sd, int(n_timesteps)

    @property
    def shape(self):
        return (self.n_timesteps, 2 * self.vals.shape[1])

    def __len__(self):
        return self.n_timesteps


def _no_worker(what):
    if torch.utils.data.get_worker_info() is not None:
        raise RuntimeError(f"{what} launches a CUDA kernel and cannot run inside a DataLoader worker process; let "
                           "MIMICDataset.__getitem__ return StayRows (host-only) and bin in the main process "
                           "(Model.feats_to_input / bin_stay_rows do it), or use num_workers=0")


def _bin_parts(parts, mu, sd, n_timesteps, device):
    device = torch.device(device) if device is not None else torch.device("cuda")
    V = mu.shape[0]
    row_start = np.zeros(len(parts) + 1, dtype=np.int64)
    np.cumsum([p[0].shape[0] for p in parts], out=row_start[1:])
    if row_start[-1] == 0:                                        # no rows at all: one inert row keeps the pointers valid
        slot, vals, cnts = np.full(1, -1, np.int32), np.zeros((1, V)), np.zeros((1, V))
    else:
        slot = np.concatenate([p[0] for p in parts])
        vals = np.concatenate([p[1] for p in parts])
        cnts = np.concatenate([p[2] for p in parts])
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)
    return ops.bin_events(up(slot), up(vals), up(cnts), up(row_start), up(mu), up(sd), int(n_timesteps))


def bin_stay_rows(rows, device=None):
    """[B, T, 2V] float32 on the device from B StayRows records (one dx_bin_events launch, main process only)."""
    _no_worker("bin_stay_rows")
    r0 = rows[0]
    return _bin_parts([(r.slot, r.vals, r.cnts) for r in rows], r0.mu, r0.sd, r0.n_timesteps, device)


def build_batch_tensors(dfs, means, stds, n_timesteps, all_vars, all_counts, device=None):
    """[B, T, 2V] float32 on the device: build_stay_tensor of every frame in `dfs`, one kernel launch."""
    _no_worker("build_batch_tensors / build_stay_tensor")
    parts = [_rows(df, n_timesteps, all_vars, all_counts) for df in dfs]
    mu = np.asarray([float(means[v]) for v in all_vars], dtype=np.float64)
    sd = np.asarray([float(stds[v]) for v in all_vars], dtype=np.float64)
    return _bin_parts(parts, mu, sd, n_timesteps, device)


def build_stay_tensor(df_stay, means, stds, n_timesteps, all_vars, all_counts, device=None):
    """The reference's per-stay builder (same positional signature): [T, 2V] float32, on the device."""
    return build_batch_tensors([df_stay], means, stds, n_timesteps, all_vars, all_counts, device)[0]


def encode_static(row, age_mean, age_std, onehot_static):
    age = (float(row["age_at_intime"]) - age_mean) / (age_std + 1e-7)
    age = float(np.nan_to_num(age, nan=0.0))
    onehot = row[onehot_static].astype(float).values
    return torch.tensor([age, *onehot], dtype=torch.float32)


def collate_into_seqs(batch):
    xs, ys = zip(*batch)
    return tuple(zip(*xs)), ys


class MIMICDataset(torch.utils.data.Dataset):
    """Same surface as the reference's MIMICDataset (mimic_dataset.py:59-91); `batch(indices)` bins a whole batch with one
    launch and returns it in the collate format the model's feats_to_input takes."""

    def __init__(self, stay_ids, icu_df, static_df, meta, device=None):
        self.stay_ids = list(stay_ids)
        self.icu_df = icu_df.set_index("stay_id").sort_index()
        self.static_df = static_df.drop_duplicates("stay_id").set_index("stay_id")
        self.meta = meta
        self.n_timesteps = meta["N_TIMESTEPS"]
        self.label_col = meta["LABEL_COL"]
        self.device = device
        self.bin_ends = torch.arange(1, self.n_timesteps + 1).float() / 24.0
        self._mu = np.asarray([float(meta["means"][v]) for v in meta["ALL_VARS"]], dtype=np.float64)
        self._sd = np.asarray([float(meta["stds"][v]) for v in meta["ALL_VARS"]], dtype=np.float64)

    def __len__(self):
        return len(self.stay_ids)

    def _static(self, sid):
        return encode_static(self.static_df.loc[sid], self.meta["age_mean"], self.meta["age_std"], self.meta["ONEHOT_STATIC"])

    def __getitem__(self, i):
        """Host-only (safe in DataLoader workers): ((StayRows, static [S], bin_ends [T]), label).  The StayRows record is
        binned on the device by Model.feats_to_input, together with the rest of its batch."""
        sid = self.stay_ids[i]
        df_stay = self.icu_df.loc[[sid]].reset_index()
        slot, vals, cnts = _rows(df_stay, self.n_timesteps, self.meta["ALL_VARS"], self.meta["ALL_COUNTS"])
        return (StayRows(slot, vals, cnts, self._mu, self._sd, self.n_timesteps), self._static(sid), self.bin_ends), \
            float(self.static_df.loc[sid, self.label_col])

    def batch(self, indices):
        sids = [self.stay_ids[i] for i in indices]
        dfs = [self.icu_df.loc[[sid]].reset_index() for sid in sids]
        x = build_batch_tensors(dfs, self.meta["means"], self.meta["stds"], self.n_timesteps, self.meta["ALL_VARS"],
                                self.meta["ALL_COUNTS"], self.device)
        xs = (tuple(x.unbind(0)), tuple(self._static(s) for s in sids), tuple(self.bin_ends for _ in sids))
        return xs, tuple(float(self.static_df.loc[s, self.label_col]) for s in sids)

    def d_static_num(self):
        return self.meta["D_STATIC"]

    def d_time_series_num(self):
        return len(self.meta["ALL_VARS"])

    def d_target(self):
        return 1

    def pos_frac(self):
        s = self.static_df.loc[self.stay_ids, self.label_col].astype(float)
        return float(s.mean())
