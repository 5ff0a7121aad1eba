"""Host logic of saragan_b200.metrics (draw order of the reference's numpy RNG, pyramid levels, the 128-fold row
de-duplication, the weighted-histogram identity behind the KS distance, get_metrics labels) on the emulated
kernels against the golden fixtures minted from the unmodified reference."""
import numpy as np
import pytest
import torch

from tests.test_metrics_oracle_cpu import load_metrics_golden


@pytest.fixture
def cpu_metrics(cpu_kernels, monkeypatch):
    from saragan_b200 import metrics
    monkeypatch.setattr(metrics, "_device", lambda: torch.device("cpu"))
    return metrics


@pytest.mark.parametrize("name", ["metrics_w32", "metrics_w64"])
def test_swd_host_logic_matches_reference(name, cpu_metrics):
    z, real, fake, seed = load_metrics_golden(name)
    got = cpu_metrics.sliced_wasserstein_distance(real, fake, rng=np.random.RandomState(seed))
    want = z["ref.swd"]
    assert len(got) == len(want)
    assert np.allclose(got, want, rtol=1e-4), (got, want)


@pytest.mark.parametrize("name", ["metrics_w32", "metrics_w64", "metrics_w128_b3"])
def test_kms_is_exact(name, cpu_metrics):
    z, real, fake, _ = load_metrics_golden(name)
    got = cpu_metrics.kolmogorov_smirnov_distance(real, fake, -1024, (-1024, 2048))
    assert float(got) == float(z["ref.kms"])


def test_global_numpy_rng_is_accepted(cpu_metrics):
    """rng=numpy.random: the global generator, which is what the reference consumes"""
    z, real, fake, seed = load_metrics_golden("metrics_w32")
    np.random.seed(seed)
    got = cpu_metrics.get_metrics(real, fake, rng=np.random)
    assert set(got) == {"mean_swd", "kms"}
    assert np.isclose(got["mean_swd"], z["ref.swd"][-1], rtol=1e-4) and got["kms"] == float(z["ref.kms"])


def test_device_rng_estimates_the_same_distance(cpu_metrics):
    z, real, fake, _ = load_metrics_golden("metrics_w64")
    torch.manual_seed(3)
    got = cpu_metrics.sliced_wasserstein_distance(torch.from_numpy(real), torch.from_numpy(fake))
    assert np.allclose(got, z["ref.swd"], rtol=0.25), (got, z["ref.swd"])
