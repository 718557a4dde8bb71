"""ORACLE / TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's input binning
(duett/mimic_dataset.py:33-46, build_stay_tensor) over plain arrays: rows are applied in order, `count > 0` gates the write,
value = (v - mean) / (std + 1e-7) in float64, stored as float32.  Pinned against the reference's own function through
tests/golden/g5_binning.npz (oracle/make_golden_binning.py)."""
import numpy as np


def bin_events(slot, vals, cnts, row_start, means, stds, T):
    B, V = len(row_start) - 1, len(means)
    x = np.zeros((B, T, 2 * V), dtype=np.float32)
    for b in range(B):
        for r in range(int(row_start[b]), int(row_start[b + 1])):
            t = int(slot[r])
            if t >= T:                 # mimic_dataset.py:38-39
                continue
            for j in range(V):
                c = cnts[r, j]
                if c > 0:              # mimic_dataset.py:42 (NaN compares false)
                    x[b, t, j] = (vals[r, j] - means[j]) / (stds[j] + 1e-7)
                    x[b, t, j + V] = c
    return x
