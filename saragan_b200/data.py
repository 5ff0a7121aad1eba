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
from queue import Full, Queue
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
    the device on a side stream (double buffered).

    Sharding follows torch's DistributedSampler as the reference uses it (main.py:106-107, `set_epoch` per epoch):
    the file list is shuffled GLOBALLY with a generator seeded by (seed, epoch) -- the same on every rank --, padded by
    wrapping around to a multiple of `world`, and rank r takes entries r, r + world, ...  Every rank therefore sees
    the same number of batches (a rank with one batch more would block forever in the per-step gradient
    all-reduce) and a different subset every epoch."""

    def __init__(self, root: str, batch_size: int, device, rank: int = 0, world: int = 1, shuffle: bool = True,
                 seed: int = 0, extension: str = ".npy", drop_last: bool = True):
        self.all_files = list_volumes(root, extension)
        if not self.all_files:
            raise FileNotFoundError(f"no {extension} volumes under {root}")
        self.rank, self.world, self.seed, self.epoch = rank, world, seed, 0
        self.per_rank = (len(self.all_files) + world - 1) // world
        self.batch_size, self.device, self.shuffle, self.drop_last = batch_size, torch.device(device), shuffle, drop_last
        self.stream = torch.cuda.Stream(device=self.device)

    def set_epoch(self, epoch: int) -> None:
        """Reshuffle for the next pass (DistributedSampler.set_epoch)."""
        self.epoch = int(epoch)

    def indices(self) -> np.ndarray:
        """This rank's file indices for the current epoch."""
        n = len(self.all_files)
        order = np.random.default_rng([self.seed, self.epoch]).permutation(n) if self.shuffle else np.arange(n)
        total = self.per_rank * self.world
        if total > n:
            order = np.concatenate([order, np.resize(order, total - n)])       # wrap around, like DistributedSampler
        return order[self.rank:total:self.world]

    @property
    def files(self) -> List[str]:
        return [self.all_files[i] for i in self.indices()]

    def __len__(self):
        n = self.per_rank // self.batch_size
        return n if self.drop_last or self.per_rank % self.batch_size == 0 else n + 1

    def _host_batches(self, q: Queue, st: dict):
        """Producer thread of ONE pass; `st` holds that pass's staging-slot hand-over state.  A staging slot is
        re-filled only after (a) the consumer has enqueued the copy out of it (`free`) and (b) that copy has
        finished (`copied`, a CUDA event)."""
        files = self.files
        pinned = [None, None]
        try:
            for bi in range(len(self)):
                vols = [np.load(f) for f in files[bi * self.batch_size:(bi + 1) * self.batch_size]]
                arr = np.stack(vols).astype(np.uint16, copy=False)
                slot = bi & 1
                if pinned[slot] is None or pinned[slot].shape != arr.shape:
                    pinned[slot] = torch.empty(arr.shape, dtype=torch.uint16).pin_memory()
                while not st["free"][slot].wait(timeout=0.1):
                    if st["stop"].is_set():
                        return
                st["free"][slot].clear()
                if st["copied"][slot] is not None:
                    st["copied"][slot].synchronize()     # the asynchronous copy out of this slot has finished
                pinned[slot].numpy()[...] = arr
                while True:
                    if st["stop"].is_set():
                        return
                    try:
                        q.put((slot, pinned[slot]), timeout=0.1)
                        break
                    except Full:
                        continue
        finally:
            if not st["stop"].is_set():
                q.put(None)

    def __iter__(self) -> Iterator[torch.Tensor]:
        q: Queue = Queue(maxsize=1)      # one batch being filled while one is in flight
        st = dict(free=[threading.Event(), threading.Event()], copied=[None, None], stop=threading.Event())
        for f in st["free"]:
            f.set()
        t = threading.Thread(target=self._host_batches, args=(q, st), daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                slot, host = item
                with torch.cuda.stream(self.stream):
                    dev = host.to(self.device, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(self.stream)
                st["copied"][slot] = done
                st["free"][slot].set()
                torch.cuda.current_stream(self.device).wait_stream(self.stream)
                dev.record_stream(torch.cuda.current_stream(self.device))
                yield dev
        finally:
            # the consumer may abandon the pass early (break, exception): release the producer and wait for it, so
            # that nothing of this pass is still running when the next one starts
            st["stop"].set()
            t.join()


def synthetic_reals(n: int, volume, seed: int, passes: int = 2) -> torch.Tensor:
    """Synthetic CT-like reals of SURVEY.md 8(d): clip(1024 + 350*smooth(N(0,1)), 0, 3072) / 1024 as (n,1,D,H,W) fp32
    on the host, deterministic per seed (box-filtered white noise; bench.py, the full-size parity fixtures)."""
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 1, *volume, generator=gen)
    k = torch.ones(1, 1, 3, 3, 3) / 27
    for _ in range(passes):
        x = torch.nn.functional.conv3d(x, k, padding=1)
    x = x / x.std()
    return (torch.clamp(1024 + 350 * x, 0, 3072) / 1024).contiguous()


def step_draws(batch: int, volume, latent_dim: int, seed: int):
    """The random draws of one train step (train.py:144-145,178, loss.py:11) from a host generator: a step can be
    replayed bit-for-bit on another implementation (the full-size parity fixtures, bench.py's parity check)."""
    gen = torch.Generator().manual_seed(seed)
    return dict(noise=torch.randn(batch, 1, *volume, generator=gen), z_d=torch.randn(batch, latent_dim, generator=gen),
                z_g=torch.randn(batch, latent_dim, generator=gen), eps=torch.rand(batch, 1, 1, 1, 1, generator=gen))
