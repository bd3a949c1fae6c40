"""NumPy restatement of the reference's per-channel dataset statistics.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows
``calculate_dataset_stats.calculate_dataset_stats`` (calculate_dataset_stats.py:23-108): two passes
over ``inputs[S, C, H, W]``, float32 per-image sums accumulated in float64.
"""
from __future__ import annotations

import numpy as np

__all__ = ["channel_stats"]


def channel_stats(inputs: np.ndarray):
    """Returns ``{'means': [C floats], 'stds': [C floats]}`` exactly like the dict the reference
    stores in ``train_stats.npy`` (calculate_dataset_stats.py:82-89)."""
    S, C = inputs.shape[0], inputs.shape[1]
    means = np.zeros(C, dtype=np.float64)            # :54
    squared_sums = np.zeros(C, dtype=np.float64)     # :55
    for i in range(S):                               # pass 1, :59-64
        sample = inputs[i].astype(np.float32)
        for c in range(C):
            means[c] += np.sum(sample[c])
    pixel_count = inputs.shape[2] * inputs.shape[3]  # :67
    means = means / (S * pixel_count)                # :68
    for i in range(S):                               # pass 2, :71-76
        sample = inputs[i].astype(np.float32)
        for c in range(C):
            squared_sums[c] += np.sum((sample[c] - means[c]) ** 2)
    stds = np.sqrt(squared_sums / (S * pixel_count))  # :79 (population std, ddof=0)
    return {"means": means.tolist(), "stds": stds.tolist()}
