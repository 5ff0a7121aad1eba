"""Input pipeline of the hot path -- SURVEY 8f row 2 (reference: pgan_pytorch/data.py:16-86,
main.py:84-118, train.py:143-144).

The reference's dataset is a directory ``{res}x{res}/NNNN.npy`` of uint16 ``(D,H,W)`` volumes
(data_scripts/create_lidc_idri_dataset.py:185-212); its loader is ``np.load -> float32 ->
[None] / 1024`` on the host followed by ``+ 0.01*randn`` on the device.  Here the raw uint16 voxels
travel to the GPU (half the H2D bytes of fp32) through double-buffered pinned staging and ONE
kernel (``sg_prepare_real``) does the cast, the 1/1024 scale and the instance noise.
"""
from __future__ import annotations

import os
import threading
from queue import Queue
from typing import Iterator, List, Optional

import numpy as np
import torch

from ._lib import call


def list_volumes(root: str, extension: str = ".npy") -> List[str]:
    """Sorted recursive file list, like data.py:16-31 (`make_dataset`)."""
    out = []
    for dirpath, _, files in sorted(os.walk(os.path.expanduser(root), followlinks=True)):
        for f in sorted(files):
            if f.lower().endswith(extension):
                out.append(os.path.join(dirpath, f))
    return out


def prepare_real(raw_u16: torch.Tensor, noise: Optional[torch.Tensor] = None, scale: float = 1.0 / 1024,
                 sigma: float = 1e-2, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B,D,H,W) or (B,1,D,H,W) uint16 on the GPU -> (B,1,D,H,W) fp32 = raw*scale + sigma*noise."""
    if raw_u16.dtype != torch.uint16:
        raise TypeError("prepare_real expects the raw uint16 voxels")
    shape = raw_u16.shape if raw_u16.dim() == 5 else (raw_u16.shape[0], 1, *raw_u16.shape[1:])
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=raw_u16.device)
    call("sg_prepare_real", raw_u16.contiguous(), None if noise is None else noise.contiguous(), out,
         raw_u16.numel(), float(scale), float(sigma if noise is not None else 0.0))
    return out


class VolumeLoader:
    """Batches of raw uint16 volumes, prefetched by a host thread into pinned buffers and copied to
    the device on a side stream (double buffered).  Rank r of `world` takes files r, r+world, ...
    (the reference's DistributedSampler without shuffling when `shuffle=False`)."""

    def __init__(self, root: str, batch_size: int, device, rank: int = 0, world: int = 1, shuffle: bool = True,
                 seed: int = 0, extension: str = ".npy", drop_last: bool = True):
        self.files = list_volumes(root, extension)[rank::world]
        if not self.files:
            raise FileNotFoundError(f"no {extension} volumes under {root}")
        self.batch_size, self.device, self.shuffle, self.drop_last = batch_size, torch.device(device), shuffle, drop_last
        self.rng = np.random.default_rng(seed + rank)
        self.stream = torch.cuda.Stream(device=self.device)
        # per staging slot: `_free` is set by the consumer once the copy out of the slot has been ENQUEUED, `_copied`
        # is the CUDA event recorded right behind that copy; the producer re-fills a slot only after both
        self._free = [threading.Event(), threading.Event()]
        self._copied = [None, None]

    def __len__(self):
        n = len(self.files) // self.batch_size
        return n if self.drop_last or len(self.files) % self.batch_size == 0 else n + 1

    def _host_batches(self, q: Queue):
        order = self.rng.permutation(len(self.files)) if self.shuffle else np.arange(len(self.files))
        pinned = [None, None]
        for bi in range(len(self)):
            idx = order[bi * self.batch_size:(bi + 1) * self.batch_size]
            vols = [np.load(self.files[i]) for i in idx]
            arr = np.stack(vols).astype(np.uint16, copy=False)
            slot = bi & 1
            if pinned[slot] is None or pinned[slot].shape != arr.shape:
                pinned[slot] = torch.empty(arr.shape, dtype=torch.uint16).pin_memory()
            self._free[slot].wait()
            self._free[slot].clear()
            if self._copied[slot] is not None:
                self._copied[slot].synchronize()     # the asynchronous copy out of this slot has finished
            pinned[slot].numpy()[...] = arr
            q.put((slot, pinned[slot]))
        q.put(None)

    def __iter__(self) -> Iterator[torch.Tensor]:
        q: Queue = Queue(maxsize=1)      # one batch being filled while one is in flight
        for f in self._free:
            f.set()
        self._copied = [None, None]
        t = threading.Thread(target=self._host_batches, args=(q,), daemon=True)
        t.start()
        while True:
            item = q.get()
            if item is None:
                break
            slot, host = item
            with torch.cuda.stream(self.stream):
                dev = host.to(self.device, non_blocking=True)
                done = torch.cuda.Event()
                done.record(self.stream)
            self._copied[slot] = done
            self._free[slot].set()
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
            dev.record_stream(torch.cuda.current_stream(self.device))
            yield dev
        t.join()
