"""Host logic of the input pipeline (saragan_b200/data.py) without a GPU: file order like the
reference's make_dataset (data.py:16-31), rank sharding like DistributedSampler without shuffling, batching,
drop_last, deterministic shuffling, and the hand-over of the double-buffered staging slots.  CUDA streams, events
and pinned memory are replaced by inert stand-ins; the copies and the kernel are covered by the GPU tests."""
import os

import numpy as np
import pytest
import torch

from saragan_b200 import data


class _Stream:
    def __init__(self, device=None):
        pass

    def wait_stream(self, other):
        pass


class _Event:
    log = []

    def record(self, stream=None):
        _Event.log.append("record")

    def synchronize(self):
        _Event.log.append("sync")


class _Ctx:
    def __init__(self, stream):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


@pytest.fixture
def no_gpu(monkeypatch):
    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "stream", _Ctx)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: _Stream())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    orig_to = torch.Tensor.to
    # host -> "device" must be a COPY like the real H2D transfer (Tensor.to on a CPU tensor returns the tensor itself,
    # which would alias the staging slot the producer thread re-fills)
    monkeypatch.setattr(torch.Tensor, "to", lambda self, *a, **k: orig_to(self, *a, **k).clone())
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None)
    _Event.log = []


def _dataset(tmp_path, n):
    d = tmp_path / "16x16"
    os.makedirs(d)
    vols = []
    for i in range(n):
        v = np.full((2, 4, 4), i, dtype=np.uint16)
        np.save(d / f"{i:04d}.npy", v)
        vols.append(v)
    (d / "notes.txt").write_text("ignored")
    return vols


def test_loader_order_sharding_and_slots(tmp_path, no_gpu):
    _dataset(tmp_path, 11)
    assert [os.path.basename(f) for f in data.list_volumes(str(tmp_path))] == [f"{i:04d}.npy" for i in range(11)]
    seen = []
    for rank in range(2):
        loader = data.VolumeLoader(str(tmp_path), batch_size=2, device="cpu", rank=rank, world=2, shuffle=False)
        # 11 files over 2 ranks: padded by wrapping to 12 like DistributedSampler, so BOTH ranks run 3 batches (a
        # rank with one batch more would block in the per-step gradient all-reduce)
        assert len(loader) == 3
        batches = [b.clone() for b in loader]
        assert all(b.dtype == torch.uint16 and tuple(b.shape) == (2, 2, 4, 4) for b in batches)
        seen.append([int(b[i, 0, 0, 0]) for b in batches for i in range(2)])
    assert seen == [[0, 2, 4, 6, 8, 10], [1, 3, 5, 7, 9, 0]]      # rank r takes entries r, r + world, ... of the padded list
    # a staging slot is re-filled only after the copy out of it was waited for: every refill of a used slot syncs
    assert _Event.log.count("record") == 6 and _Event.log.count("sync") >= 1


def test_loader_matches_distributed_sampler_semantics(tmp_path, no_gpu):
    """Same contract as torch.utils.data.DistributedSampler (main.py:106-107): one GLOBAL permutation per (seed, epoch)
    shared by the ranks, padded by wrap-around, strided by rank; set_epoch reshuffles; equal length on every rank."""
    _dataset(tmp_path, 10)
    world = 4
    loaders = [data.VolumeLoader(str(tmp_path), 1, "cpu", rank=r, world=world, seed=5) for r in range(world)]
    per_epoch = []
    for epoch in range(2):
        for ld in loaders:
            ld.set_epoch(epoch)
        idx = [list(ld.indices()) for ld in loaders]
        assert {len(i) for i in idx} == {3} and {len(ld) for ld in loaders} == {3}
        merged = [idx[r][k] for k in range(3) for r in range(world)]          # undo the stride
        assert sorted(merged[:10]) == list(range(10)) and merged[10:] == merged[:2]   # a permutation + wrap-around pad
        ref = torch.utils.data.DistributedSampler(list(range(10)), num_replicas=world, rank=1, shuffle=True, seed=5)
        ref.set_epoch(epoch)
        assert len(list(ref)) == len(idx[1])                                  # same per-rank count as torch's sampler
        per_epoch.append(merged)
    assert per_epoch[0] != per_epoch[1]                                       # set_epoch reshuffles


def test_loader_survives_an_abandoned_pass(tmp_path, no_gpu):
    """Breaking out of a pass stops its producer thread; the next pass starts clean (own slot state per pass)."""
    import threading
    _dataset(tmp_path, 12)
    loader = data.VolumeLoader(str(tmp_path), 2, "cpu", shuffle=False)
    n_threads = threading.active_count()
    it = iter(loader)
    first = next(it)
    it.close()                                                                # consumer walks away after one batch
    assert threading.active_count() == n_threads
    again = [int(b[0, 0, 0, 0]) for b in loader]
    assert int(first[0, 0, 0, 0]) == 0 and again == [0, 2, 4, 6, 8, 10]


def test_loader_keeps_the_last_partial_batch_on_request(tmp_path, no_gpu):
    _dataset(tmp_path, 5)
    loader = data.VolumeLoader(str(tmp_path), batch_size=2, device="cpu", shuffle=False, drop_last=False)
    shapes = [tuple(b.shape) for b in loader]
    assert len(loader) == 3 and shapes == [(2, 2, 4, 4), (2, 2, 4, 4), (1, 2, 4, 4)]


def test_loader_shuffle_is_deterministic_per_seed_and_rank(tmp_path, no_gpu):
    _dataset(tmp_path, 8)
    order = lambda seed: [int(b[i, 0, 0, 0]) for b in data.VolumeLoader(str(tmp_path), 2, "cpu", seed=seed) for i in range(2)]   # noqa: E731
    a, b, c = order(3), order(3), order(4)
    assert a == b and sorted(a) == list(range(8)) and a != c


def test_empty_directory_raises(tmp_path, no_gpu):
    with pytest.raises(FileNotFoundError):
        data.VolumeLoader(str(tmp_path), 2, "cpu")


def test_synthetic_pyramid_has_the_reference_layout(tmp_path, no_gpu):
    """tools/make_synthetic_volumes.py writes what main.py:69-91 expects: {res}x{res}/NNNN.npy, (res/4, res, res)
    uint16 in [0, 3072]; the loader reads it back in file order."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(os.path.dirname(os.path.dirname(__file__)),
                                                                    "tools", "make_synthetic_volumes.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    d = mk.write_phase(str(tmp_path), 3, 6)
    assert os.path.basename(d) == "16x16" and sorted(os.listdir(d)) == [f"{i:04d}.npy" for i in range(6)]
    v = np.load(os.path.join(d, "0003.npy"))
    assert v.shape == (4, 16, 16) and v.dtype == np.uint16 and v.max() <= 3072 and 600 < v.mean() < 1400
    again = mk.synthetic_volume(np.random.default_rng(1234 + 3), 16)
    assert np.array_equal(again, np.load(os.path.join(d, "0000.npy")))             # deterministic per phase
    batches = list(data.VolumeLoader(d, batch_size=3, device="cpu", shuffle=False))
    assert len(batches) == 2 and np.array_equal(batches[1][0].numpy(), v)


@pytest.mark.skipif(not os.path.exists("/root/reference/pgan_pytorch/data.py"), reason="reference tree not present")
def test_loader_agrees_with_the_reference_dataset_class(tmp_path, no_gpu):
    """The reference's own DatasetFolder + DataLoader (data.py:34-89 with main.py:84-118's loader / transform) and
    VolumeLoader read the same files in the same order; raw / 1024 (what sg_prepare_real computes without noise)
    equals the reference's float batches."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_data", "/root/reference/pgan_pytorch/data.py")
    ref_data = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_data)
    spec = importlib.util.spec_from_file_location("mk", os.path.join(os.path.dirname(os.path.dirname(__file__)),
                                                                    "tools", "make_synthetic_volumes.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    d = mk.write_phase(str(tmp_path), 3, 7)
    dataset = ref_data.DatasetFolder(d, loader=lambda path: torch.from_numpy(np.load(path).astype(np.float32)),
                                     extensions=(".npy",), transform=lambda x: x.unsqueeze(0) / 1024)     # main.py:84-91
    ref_batches = list(torch.utils.data.DataLoader(dataset, batch_size=2, shuffle=False, drop_last=True))
    ours = list(data.VolumeLoader(d, batch_size=2, device="cpu", shuffle=False))
    assert len(ours) == len(ref_batches) == 3
    for a, b in zip(ours, ref_batches):
        assert torch.equal(torch.from_numpy(a.numpy().astype(np.float32))[:, None] / 1024, b)
