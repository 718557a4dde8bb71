"""ORACLE / TEST INFRASTRUCTURE ONLY — generates tests/golden/g5_binning.npz by running the REFERENCE'S OWN
build_stay_tensor (duett/mimic_dataset.py:33-46, imported unmodified from /root/reference) on synthetic stay frames that
exercise: several rows per slot (later row wins), slots >= n_timesteps (skipped), negative slots (python indexing), zero and
NaN counts, NaN values, two variables sharing one count column, an empty stay.

Run in the authoring container only:   python oracle/make_golden_binning.py"""
import importlib.util
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DUETT_REFERENCE", "/root/reference")
spec = importlib.util.spec_from_file_location("ref_mimic_dataset", os.path.join(REF, "duett", "mimic_dataset.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(7)
T, V = 12, 9
all_vars = [f"v{j}" for j in range(V)]
all_counts = [f"count_v{j}" for j in range(V)]
all_counts[5] = all_counts[2]                    # shared count column (the reference's EXTRA variables)
uniq = list(dict.fromkeys(all_counts))
means = {v: float(rng.normal()) for v in all_vars}
stds = {v: float(abs(rng.normal()) + 0.3) for v in all_vars}
stds["v4"] = 0.0                                  # exercises the + 1e-7

frames = []
for b in range(7):
    R = [0, 5, 17, 30, 12, 9, 40][b]
    slot = rng.integers(-2, T + 3, size=R)
    if b == 2:
        slot = np.sort(slot)
    df = pd.DataFrame({"stay_id": 1000 + b, "slot_idx": slot})
    for v in all_vars:
        col = rng.normal(size=R) * 3
        col[rng.random(R) < 0.1] = np.nan
        df[v] = col
    for c in uniq:
        col = rng.integers(0, 4, size=R).astype(float) * (rng.random(R) < 0.6)
        col[rng.random(R) < 0.05] = np.nan
        df[c] = col
    frames.append(df)

outs = [ref.build_stay_tensor(df, means, stds, T, all_vars, all_counts).numpy() for df in frames]
row_start = np.zeros(len(frames) + 1, dtype=np.int64)
np.cumsum([len(df) for df in frames], out=row_start[1:])
slot = np.concatenate([df["slot_idx"].to_numpy() for df in frames]).astype(np.int64)
blobs = {
    "T": np.int64(T), "slot_raw": slot,
    "vals": np.concatenate([df[all_vars].to_numpy(dtype=np.float64) for df in frames]),
    "cnts": np.concatenate([df[all_counts].to_numpy(dtype=np.float64) for df in frames]),
    "row_start": row_start, "means": np.asarray([means[v] for v in all_vars]), "stds": np.asarray([stds[v] for v in all_vars]),
    "x": np.stack(outs), "shared_count_of": np.asarray([all_counts.index(c) for c in all_counts], dtype=np.int64),
}
path = os.path.join(ROOT, "tests", "golden", "g5_binning.npz")
np.savez_compressed(path, **blobs)
print("wrote", path, os.path.getsize(path), "bytes;", int(np.isnan(np.stack(outs)).sum()), "NaN cells,", int((np.stack(outs) != 0).sum()), "non-zero cells")
