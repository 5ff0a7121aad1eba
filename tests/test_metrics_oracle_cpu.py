"""The numpy restatement of the reference's evaluation metrics (oracle/metrics_oracle.py) against the golden
fixtures minted from the UNMODIFIED reference (oracle/pin_metrics_against_reference.py): bit-exact."""
import os

import numpy as np
import pytest

from oracle import metrics_oracle as M
from tests.util import GOLDEN

CASES = ["metrics_w32", "metrics_w64", "metrics_w128_b3"]


def load_metrics_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, z["real"].astype(np.float32), z["fake"].astype(np.float32), int(z["seed"])


@pytest.mark.parametrize("name", CASES)
def test_swd_and_kms_match_reference_goldens(name):
    z, real, fake, seed = load_metrics_golden(name)
    got = M.sliced_wasserstein_distance(real.copy(), fake.copy(), np.random.RandomState(seed))
    assert [float(v) for v in got] == [float(v) for v in z["ref.swd"]]
    kms = M.kolmogorov_smirnov_distance(real.copy(), fake.copy(), -1024, (-1024, 2048))
    assert float(kms) == float(z["ref.kms"])


def test_laplacian_pyramid_matches_reference_goldens():
    z, real, _, _ = load_metrics_golden("metrics_w64")
    pyr = M.generate_laplacian_pyramid(real.copy(), len(M.swd_resolutions(real.shape[-1])))
    for i, p in enumerate(pyr):
        assert np.array_equal(p, z[f"ref.pyr{i}"])


def test_get_metrics_labels():
    """train.py:12-27: the lowest-resolution level is logged as 'mean_swd', the others as swd_<res>."""
    z, real, fake, seed = load_metrics_golden("metrics_w128_b3")
    d = M.get_metrics(real.copy(), fake.copy(), np.random.RandomState(seed))
    assert set(d) == {"mean_swd", "swd_64", "swd_128", "kms"}
    assert d["mean_swd"] == float(z["ref.swd"][-1]) and d["swd_128"] == float(z["ref.swd"][0])
