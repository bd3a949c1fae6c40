#!/usr/bin/env python
"""Drop-in for the reference ``calculate_dataset_stats.py`` on B200.

``calculate_dataset_stats(dataset_path, save_path=None)`` keeps the reference's contract
(calculate_dataset_stats.py:23-108): reads ``<dataset_path>/train_inputs.npy`` ([S,C,H,W], memory-mapped), computes the
per-channel mean and population standard deviation in two passes with float64 accumulation, writes
``train_stats.npy`` (pickled dict ``{'means': [...], 'stds': [...]}``, :82-89) and ``train_stats.json`` (:92-95) and
returns the dict (``None`` on error, like the reference).  The two passes stream the samples through pinned host
buffers into the ``nsm_channel_sums`` reduction kernel instead of looping over samples and channels in NumPy.
"""
from __future__ import annotations

import argparse
import json
import logging
import os

import numpy as np
import torch

import nsm


def setup_logging():
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s",
                        handlers=[logging.FileHandler("dataset_stats.log"), logging.StreamHandler()])


def _stream_pass(inputs, means, device, chunk_bytes=256 << 20):
    """One pass over the memory-mapped array: sum x (means None) or sum (x - mean_c)^2, float64 [C] on host."""
    S, C = inputs.shape[0], inputs.shape[1]
    per_sample = int(np.prod(inputs.shape[1:])) * 4
    step = max(1, chunk_bytes // per_sample)
    total = torch.zeros(C, dtype=torch.float64, device=device)
    m = None if means is None else torch.as_tensor(means, dtype=torch.float64, device=device)
    bufs = [torch.empty((step,) + tuple(inputs.shape[1:]), dtype=torch.float32).pin_memory() for _ in range(2)]
    events = [None, None]
    for k, i in enumerate(range(0, S, step)):
        n = min(step, S - i)
        b = bufs[k & 1]
        if events[k & 1] is not None:
            events[k & 1].synchronize()                      # buffer free again
        np.copyto(b[:n].numpy(), inputs[i:i + n], casting="same_kind")
        xd = b[:n].to(device, non_blocking=True)
        total += nsm.channel_sums(xd, m)
        ev = torch.cuda.Event()
        ev.record()
        events[k & 1] = ev
    return total.cpu().numpy()


def calculate_dataset_stats(dataset_path, save_path=None, device="cuda"):
    if save_path is None:
        save_path = dataset_path
    inputs_path = os.path.join(dataset_path, "train_inputs.npy")
    if not os.path.exists(inputs_path):
        logging.error(f"找不到训练数据文件: {inputs_path}")
        return None
    try:
        nsm.require_device()
        inputs = np.load(inputs_path, mmap_mode="r")
        logging.info(f"数据集形状: {inputs.shape}")
        S = inputs.shape[0]
        pixel_count = inputs.shape[2] * inputs.shape[3]
        means = _stream_pass(inputs, None, device) / (S * pixel_count)
        squared_sums = _stream_pass(inputs, means, device)
        stds = np.sqrt(squared_sums / (S * pixel_count))
        stats = {"means": means.tolist(), "stds": stds.tolist()}
        np.save(os.path.join(save_path, "train_stats.npy"), stats)
        with open(os.path.join(save_path, "train_stats.json"), "w", encoding="utf-8") as f:
            json.dump(stats, f, indent=4)
        logging.info(f"数据集通道均值: {means}")
        logging.info(f"数据集通道标准差: {stds}")
        return stats
    except Exception as e:  # same error convention as the reference (:102-108)
        logging.error(f"计算统计数据时出错: {str(e)}")
        import traceback
        logging.error(traceback.format_exc())
        return None


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="计算数据集的通道均值和标准差")
    parser.add_argument("--dataset_path", type=str, required=True)
    parser.add_argument("--save_path", type=str, default=None)
    args = parser.parse_args()
    setup_logging()
    calculate_dataset_stats(args.dataset_path, args.save_path)
