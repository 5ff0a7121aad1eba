"""SURVEY 8f "next" rows built so far: fused Adam (+EMA) and the uint16 input preparation."""
import os

import numpy as np
import pytest
import torch

import saragan_b200 as sg
from saragan_b200 import data

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("betas", [(0.0, 0.99), (0.9, 0.999)])
def test_fused_adam_matches_torch_adam(betas):
    gen = torch.Generator().manual_seed(0)
    shapes = [(64, 32, 3, 3, 3), (64,), (1, 7), (1025,), (3, 5, 1, 1, 1)]
    ps = [torch.randn(s, generator=gen) for s in shapes]
    a = [torch.nn.Parameter(p.clone().cuda()) for p in ps] + [torch.nn.Parameter(torch.zeros(5, device="cuda"))]
    b = [torch.nn.Parameter(p.clone().cuda()) for p in ps] + [torch.nn.Parameter(torch.zeros(5, device="cuda"))]
    oa = torch.optim.Adam(a, lr=1e-3, betas=betas)
    ob = sg.FusedAdam(b, lr=1e-3, betas=betas, ema_beta=0.99)
    ema_ref = [p.detach().clone() for p in a[:-1]]
    for step in range(5):
        for pa, pb in zip(a[:-1], b[:-1]):            # the last parameter never gets a gradient (inactive level)
            g = torch.randn(pa.shape, generator=gen).cuda() * (10.0 ** (step - 2))
            pa.grad, pb.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
        ema_ref = [0.99 * e + 0.01 * p.detach() for e, p in zip(ema_ref, a[:-1])]
    for pa, pb in zip(a, b):
        assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-6), float((pa - pb).abs().max())
    ema = ob.ema_state()
    for e, pb in zip(ema_ref, b[:-1]):
        assert torch.allclose(e, ema[pb], rtol=1e-5, atol=1e-6)
    assert b[-1] not in ema and float(b[-1].abs().max()) == 0.0


def test_prepare_real_and_loader(tmp_path):
    rng = np.random.default_rng(1)
    d = tmp_path / "16x16"
    os.makedirs(d)
    vols = []
    for i in range(5):
        v = rng.integers(0, 3072, size=(4, 16, 16), dtype=np.uint16)
        np.save(d / f"{i:04d}.npy", v)
        vols.append(v)
    assert [os.path.basename(f) for f in data.list_volumes(str(tmp_path))] == [f"{i:04d}.npy" for i in range(5)]
    loader = data.VolumeLoader(str(tmp_path), batch_size=2, device="cuda", shuffle=False)
    assert len(loader) == 2
    batches = list(loader)
    assert len(batches) == 2 and batches[0].dtype == torch.uint16 and tuple(batches[0].shape) == (2, 4, 16, 16)
    assert np.array_equal(batches[1].cpu().numpy(), np.stack(vols[2:4]))
    noise = torch.randn(2, 1, 4, 16, 16, device="cuda")
    x = data.prepare_real(batches[0], noise)
    want = torch.from_numpy(np.stack(vols[0:2]).astype(np.float32))[:, None].cuda() / 1024 + 1e-2 * noise
    assert tuple(x.shape) == (2, 1, 4, 16, 16) and torch.allclose(x, want, rtol=1e-6, atol=1e-6)
    x0 = data.prepare_real(batches[0])
    assert torch.allclose(x0, want - 1e-2 * noise, rtol=1e-6, atol=1e-6)
    ragged = torch.arange(7, dtype=torch.int32).to(torch.uint16).cuda().reshape(1, 1, 1, 7)   # scalar tail path
    assert torch.allclose(data.prepare_real(ragged).flatten(), torch.arange(7, device="cuda") / 1024)


def test_fused_adam_step_invalidates_packed_weight_caches():
    """FusedAdam writes the parameters through raw pointers; the conv layers cache packed copies of their weights
    keyed on the version counter.  After a step the next forward must see the NEW weights."""
    import saragan_b200 as sg
    from saragan_b200.optim import FusedAdam
    conv = sg.EqualizedConv3d(16, 16, 3, padding=1).cuda()
    opt = FusedAdam(conv.parameters(), lr=0.5, betas=(0.0, 0.99))
    x = torch.randn(2, 16, 2, 8, 8, device="cuda")
    y0 = conv(x)
    y0.sum().backward()
    v0 = conv.weight._version
    opt.step()
    assert conv.weight._version > v0
    with torch.no_grad():
        y1 = conv(x)
        ref = torch.nn.functional.conv3d(x, conv.weight * conv.std, conv.bias, 1, 1)
    assert float((y1 - y0.detach()).abs().max()) > 1e-2          # lr 0.5: the output must move
    assert float((y1.float() - ref).abs().max()) < 5e-2 * float(ref.abs().max())

