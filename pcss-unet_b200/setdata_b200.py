"""Input feed for the B200 training path.

The reference's ``MmapLiverDataset.__getitem__`` (setdata.py:296-328) memory-maps ``{split}_inputs.npy``, converts one
sample to float32, standardises it on ONE CPU thread and hands pageable tensors to a ``DataLoader(num_workers=0)``
(main.py:923-924); the H2D copy is synchronous (main.py:260-261).  At ~1000 samples/s per GPU x 4 MB per sample that starves
the GPU.  Here:

* ``MmapLiverDataset`` keeps the reference's constructor and file conventions (``{split}_inputs.npy``,
  ``{split}_labels.npy``, ``train_stats.npy`` with ``{'means','stds'}``) but returns RAW float32 samples;
* ``DeviceFeeder`` gathers batches into pinned double buffers, copies them on a side stream while the previous step
  computes, and applies ``(x - mean) / (std + 1e-8)`` on the GPU with ``nsm_standardize`` (bit-identical to setdata.py:316).
  The yielded inputs have ``requires_grad=True`` like the reference's samples (setdata.py:325-326).
  (Alternatively pass raw batches to the model after ``Unet.set_input_stats`` and the standardisation is fused into the
  first kernel's load.)
"""
from __future__ import annotations

import logging
import os

import numpy as np
import torch
from torch.utils.data import Dataset

import nsm


class MmapLiverDataset(Dataset):
    CHANNEL_MEANS = [0.0, 0.0, 0.0, 0.0]
    CHANNEL_STDS = [1.0, 1.0, 1.0, 1.0]

    def __init__(self, data_dir, split="train", stats_dir=None, transform=None, target_transform=None,
                 apply_normalization=True):
        stats_dir = data_dir if stats_dir is None else stats_dir
        self.inputs_path = os.path.join(data_dir, f"{split}_inputs.npy")
        self.labels_path = os.path.join(data_dir, f"{split}_labels.npy")
        self.split = split
        self.apply_normalization = apply_normalization
        for p in (self.inputs_path, self.labels_path):
            if not os.path.exists(p):
                raise FileNotFoundError(f"数据文件不存在 ({split} split): {p}")
        self.inputs = np.load(self.inputs_path, mmap_mode="r")
        self.labels = np.load(self.labels_path, mmap_mode="r")
        if self.inputs.shape[0] != self.labels.shape[0]:
            raise ValueError(f"输入数据 ({self.inputs.shape[0]}) 和标签 ({self.labels.shape[0]}) 数量不匹配 ({split} split)")
        self.transform, self.target_transform = transform, target_transform
        self.means = torch.tensor(self.CHANNEL_MEANS, dtype=torch.float32)
        self.stds = torch.tensor(self.CHANNEL_STDS, dtype=torch.float32)
        stats_path = os.path.join(stats_dir, "train_stats.npy")      # always the TRAIN statistics (setdata.py:262)
        if apply_normalization and os.path.exists(stats_path):
            try:
                stats = np.load(stats_path, allow_pickle=True).item()
                if len(stats.get("means", [])) == 4 and len(stats.get("stds", [])) == 4:
                    self.means = torch.tensor(stats["means"], dtype=torch.float32)
                    self.stds = torch.tensor(stats["stds"], dtype=torch.float32)
            except Exception as e:  # same fallback as the reference: keep the defaults
                logging.warning(f"({split} split) 从 {stats_path} 加载统计数据失败: {e}")

    def __len__(self):
        return len(self.inputs)

    def __getitem__(self, index):
        """RAW sample (float32, not standardised): standardisation happens on the GPU in DeviceFeeder / the model."""
        x = torch.from_numpy(np.ascontiguousarray(self.inputs[index], dtype=np.float32))
        y = torch.from_numpy(np.ascontiguousarray(self.labels[index], dtype=np.float32))
        if self.transform is not None:
            x = self.transform(x)
        if self.target_transform is not None:
            y = self.target_transform(y)
        return x, y


class DeviceFeeder:
    """for inputs, labels in DeviceFeeder(dataset, batch_size, device): ...   (sequential order like main.py's
    shuffle=False loaders; drop_last=False)."""

    def __init__(self, dataset: MmapLiverDataset, batch_size, device="cuda", standardize=True, requires_grad=True,
                 rank=0, world=1):
        self.ds, self.bs, self.dev = dataset, batch_size, torch.device(device)
        self.standardize = standardize and dataset.apply_normalization
        self.requires_grad = requires_grad
        self.indices = list(range(rank, len(dataset), world)) if world > 1 else list(range(len(dataset)))
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        xs, ys = dataset.inputs.shape[1:], dataset.labels.shape[1:]
        self.pin_x = [torch.empty((batch_size,) + tuple(xs), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.pin_y = [torch.empty((batch_size,) + tuple(ys), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.free = [None, None]      # events: the pinned buffer may be overwritten again
        self.mean = dataset.means.to(self.dev)
        self.std = dataset.stds.to(self.dev)

    def __len__(self):
        return (len(self.indices) + self.bs - 1) // self.bs

    def _stage(self, k, idx):
        """Gather batch `idx` into pinned buffer k and start its H2D copy on the side stream."""
        slot = k & 1
        if self.free[slot] is not None:
            self.free[slot].synchronize()
        n = len(idx)
        px, py = self.pin_x[slot][:n], self.pin_y[slot][:n]
        for j, i in enumerate(idx):       # mmap read + float32 conversion straight into pinned memory
            np.copyto(px[j].numpy(), self.ds.inputs[i], casting="unsafe")
            np.copyto(py[j].numpy(), self.ds.labels[i], casting="unsafe")
        with torch.cuda.stream(self.copy_stream):
            xd = px.to(self.dev, non_blocking=True)
            yd = py.to(self.dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.free[slot] = ev
        return xd, yd, ev

    def __iter__(self):
        nsm.require_device()
        batches = [self.indices[i:i + self.bs] for i in range(0, len(self.indices), self.bs)]
        nxt = self._stage(0, batches[0]) if batches else None
        for k in range(len(batches)):
            xd, yd, ev = nxt
            nxt = self._stage(k + 1, batches[k + 1]) if k + 1 < len(batches) else None   # overlaps the step below
            torch.cuda.current_stream().wait_event(ev)
            xd.record_stream(torch.cuda.current_stream())
            yd.record_stream(torch.cuda.current_stream())
            if self.standardize:
                xd = nsm.standardize(xd, self.mean, self.std)
            yield xd.requires_grad_(self.requires_grad), yd
