"""SURVEY 8f row 4 on the GPU: the evaluation-metric kernels (csrc/metrics.cu) through the C ABI against their
numpy/scipy restatements, and saragan_b200.metrics end to end against the golden outputs of the unmodified reference
(tests/golden/metrics_*.npz, minted by oracle/pin_metrics_against_reference.py).

Tolerances: the KS distance is bit-exact (integer counts on the device, numpy's own histogram arithmetic on <= 3073
distinct values on the host); pyramid levels 1e-6 norm-wise (fp64 stencil sums like scipy, one rounding); sliced
Wasserstein distances 1e-4 relative (fp32 projections of ~1e5 terms in a different summation order than BLAS)."""
import numpy as np
import pytest
import torch

from saragan_b200 import kernels as K
from saragan_b200 import metrics as M
from tests import cpu_emul as E
from tests.test_metrics_oracle_cpu import CASES, load_metrics_golden
from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _vol(shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("shape", [(2, 1, 8, 32, 32), (1, 2, 3, 5, 7), (3, 1, 1, 2, 9), (1, 1, 2, 2, 2), (2, 1, 6, 10, 4)])
def test_pyr_down_matches_scipy(shape):
    x = _vol(shape, 0)
    got = K.pyr_down(x.cuda())
    want = E.pyr_down(x)
    assert got.shape == want.shape and rel_err(got, want) < 1e-6


@pytest.mark.parametrize("cshape", [(2, 1, 4, 16, 16), (1, 2, 1, 3, 5), (3, 1, 2, 1, 1), (1, 1, 3, 4, 2)])
def test_pyr_up_sub_matches_scipy(cshape):
    coarse = _vol(cshape, 1)
    fine = _vol(cshape[:2] + tuple(2 * s for s in cshape[2:]), 2)
    got = K.pyr_up_sub(fine.cuda(), coarse.cuda())
    assert rel_err(got, E.pyr_up_sub(fine, coarse)) < 1e-6
    with pytest.raises(ValueError):
        K.pyr_up_sub(fine.cuda()[..., :-1].contiguous(), coarse.cuda())


@pytest.mark.parametrize("name", ["metrics_w32", "metrics_w64"])
def test_laplacian_pyramid_matches_reference_goldens(name):
    z, real, _, _ = load_metrics_golden(name)
    pyr = M.generate_laplacian_pyramid(torch.from_numpy(real).cuda(), len(M.swd_resolutions(real.shape[-1])))
    for i, p in enumerate(pyr):
        want = torch.from_numpy(z[f"ref.pyr{i}"])
        assert p.shape == want.shape and rel_err(p, want) < 1e-6, i


@pytest.mark.parametrize("b,d,h,w,n", [(2, 8, 32, 32, 256), (3, 3, 9, 12, 16), (5, 4, 16, 64, 640)])
def test_descriptors_match_restatement(b, d, h, w, n):
    level = _vol((b, 1, d, h, w), 3)
    g = torch.Generator().manual_seed(4)
    pz = torch.randint(1, d - 1, (n,), generator=g, dtype=torch.int32)
    py = torch.randint(4, h - 4, (n,), generator=g, dtype=torch.int32)
    px = torch.randint(4, w - 4, (n,), generator=g, dtype=torch.int32)
    want = torch.zeros((b, n * 243))
    E.swd_descriptors(level, pz, py, px, want)
    both = torch.full((2 * b, n * 243), float("nan"), device="cuda")        # written as one arm of a stacked matrix
    K.swd_descriptors(level.cuda(), pz.cuda(), py.cuda(), px.cuda(), both[b:])
    assert rel_err(both[b:], want) < 1e-5
    assert torch.isnan(both[:b]).all()
    # a constant neighbourhood standardises to 0/0 = NaN, as the reference's numpy arithmetic does
    flat = torch.zeros((1, 1, 3, 9, 9), device="cuda")
    one = torch.zeros((1, 243), device="cuda")
    K.swd_descriptors(flat, *(torch.tensor([v], dtype=torch.int32, device="cuda") for v in (1, 4, 4)), one)
    assert torch.isnan(one).all()


@pytest.mark.parametrize("r,k", [(4, 243 * 256), (6, 1000), (20, 243 * 128 * 3), (1, 63)])
def test_project_and_finish(r, k):
    g = torch.Generator().manual_seed(5)
    a = torch.randn(r, k, generator=g)
    dirs = torch.randn(k, 128, generator=g)
    p, colsq = K.swd_project(a.cuda(), dirs.cuda(), True)
    assert rel_err(p, a.double() @ dirs.double()) < 2e-5
    assert rel_err(colsq, (dirs.double() ** 2).sum(0)) < 1e-5
    p2, none = K.swd_project(a.cuda(), dirs.cuda(), False)
    assert none is None and rel_err(p2, p) < 1e-6
    if r % 2 == 0:
        for cs in (None, colsq):
            out, want = torch.zeros(4, device="cuda"), torch.zeros(4)
            K.swd_finish(p, cs, out)
            E.swd_finish(p.cpu(), None if cs is None else cs.cpu(), want)
            assert abs(float(out[0]) - float(want[0])) < 1e-5 * abs(float(want[0]))


@pytest.mark.parametrize("n,v", [(3, 70001), (1, 5), (4, 1 << 20)])
def test_value_hist_is_exact(n, v):
    x = (torch.rand(n, v, generator=torch.Generator().manual_seed(6)) * 4.4 - 2.2)      # some values clip on both sides
    x[0, :3] = torch.tensor([-1.0, 0.0, 0.99951171875])[: min(3, v)]
    got = K.value_hist(x.cuda(), -1024.0, -1024, 2048)
    assert torch.equal(got.cpu(), E.value_hist(x, -1024.0, -1024, 2048))
    assert int(got.sum()) == n * v


@pytest.mark.parametrize("name", CASES)
def test_metrics_match_reference_goldens(name):
    """the reference's own outputs for the same volumes and the same numpy random stream"""
    z, real, fake, seed = load_metrics_golden(name)
    swd = M.sliced_wasserstein_distance(real, fake, rng=np.random.RandomState(seed))
    assert len(swd) == len(z["ref.swd"])
    assert np.allclose(swd, z["ref.swd"], rtol=1e-4), (swd, z["ref.swd"])
    kms = M.kolmogorov_smirnov_distance(torch.from_numpy(real).cuda(), torch.from_numpy(fake).cuda(), -1024, (-1024, 2048))
    assert float(kms) == float(z["ref.kms"])
    np.random.seed(seed)
    d = M.get_metrics(real, fake, rng=np.random)                # the global generator, as the reference uses it
    want_keys = {"kms", "mean_swd"} | {f"swd_{32 * 2 ** i}" for i in range(1, len(swd))}
    assert set(d) == want_keys and np.isclose(d["mean_swd"], z["ref.swd"][-1], rtol=1e-4)


def test_full_size_properties():
    """BASELINE cfg3 shape (B=4, 32x128x128), device random numbers: the KS distance of a batch to itself is 0 and is
    symmetric; the sliced Wasserstein distance is invariant under an affine intensity map applied to both arms
    (descriptors are standardised) when the random stream is the same, and close to the host-stream estimate."""
    g = torch.Generator(device="cuda").manual_seed(7)
    real = torch.randn(4, 1, 32, 128, 128, device="cuda", generator=g)
    real = torch.nn.functional.avg_pool3d(real, 3, 1, 1) * 0.9
    fake = torch.randn(4, 1, 32, 128, 128, device="cuda", generator=g) * 0.4 + 0.05
    assert M.kolmogorov_smirnov_distance(real, real.clone(), -1024, (-1024, 2048)) == 0.0
    k1 = M.kolmogorov_smirnov_distance(real, fake, -1024, (-1024, 2048))
    assert k1 > 0 and k1 == M.kolmogorov_smirnov_distance(fake, real, -1024, (-1024, 2048))
    torch.manual_seed(11)
    a = M.sliced_wasserstein_distance(real, fake)
    torch.manual_seed(11)
    b = M.sliced_wasserstein_distance(2.0 * real + 3.0, 2.0 * fake + 3.0)
    assert len(a) == 3 and all(np.isfinite(a)) and np.allclose(a, b, rtol=2e-3), (a, b)
    torch.manual_seed(12)
    c = M.sliced_wasserstein_distance(real, fake)
    assert np.allclose(a, c, rtol=0.3), (a, c)                  # another random stream: same estimate, new noise
