"""GPU parity of the population diversity metric: pairwise centred kernel alignment
(src/autoks/core/covariance.py:278-338) against a verbatim NumPy restatement (two N^3 matmuls per pair)."""
import numpy as np
import pytest

from conftest import to_oracle_expr, from_string, make_data
from oracle import kernels as ok, scoring as osc

pytestmark = pytest.mark.gpu

NAMES = ['SE_1', 'RQ_2', 'PER_1', 'LIN_2', 'SE_1 + RQ_2', 'SE_1 * PER_1 + RQ_2', '(SE_1 + RQ_2) * LIN_2', 'LIN_1 * LIN_1 * PER_1',
         '(SE_1 + PER_2) * (RQ_1 + LIN_2)', '(SE_1 + RQ_2) * (PER_1 + LIN_2 + SE_2) * RQ_1',
         '(SE_1 + RQ_1 + PER_1) * (SE_2 + RQ_2 + LIN_2) * (PER_2 + LIN_1 + SE_2)', 'SE_2']


@pytest.mark.parametrize('n', [7, 130, 300])
def test_pairwise_centered_alignments_match_reference_formula(n):
    from ms_project_b200.covariance import Covariance, pairwise_centered_alignments, centered_alignment_of_kernels
    X, _ = make_data(n, 2, seed=n)
    rng = np.random.RandomState(n)
    kernels = [from_string(s) for s in NAMES]
    for k in kernels:
        k[:] = np.exp(rng.uniform(-0.7, 0.7, size=k.size))
    covs = [Covariance(k) for k in kernels]
    got = pairwise_centered_alignments(covs, X)
    mats = [ok.K(to_oracle_expr(k), k.param_array, X, None, 'direct') for k in kernels]
    ref = osc.pairwise_centered_alignments(mats)
    assert got.shape == ref.shape == (len(kernels), len(kernels))
    assert np.all(np.diag(got) == 0) and np.allclose(got, got.T)
    # the reference's quirk: (dists + dists.T) / 2 of an upper-triangular matrix halves every alignment
    assert np.allclose(got, ref, rtol=1e-9, atol=1e-11), np.max(np.abs(got - ref))
    a01 = centered_alignment_of_kernels(covs[0], covs[5], X)
    assert abs(a01 - osc.centered_alignment(mats[0], mats[5])) <= 1e-9
    assert np.all(2 * got <= 1 + 1e-9) and np.all(2 * got >= -1 - 1e-9)
