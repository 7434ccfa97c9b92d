"""Evaluation-stage statistics on the device (gn_kde2d_pdf_f32, gn_overlap_sums_f32 through bbh.gaussian_kde2d /
bbh.overlap_beta) against the float64 oracle and the SciPy-made golden vectors.  Tolerance: the density is a sum of
n float32 exponentials of arguments formed from float32 differences: 1e-5 of the peak density, overlap score 1e-5."""
import os

import numpy as np
import pytest

from oracle import stats_oracle as st

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'kde_overlap.npz')


@pytest.fixture(scope='module')
def bbh():
    from gennet_b200 import bbh as m
    return m


def test_kde_pdf_and_overlap_match_scipy_golden(bbh):
    g = np.load(GOLD)
    pos = g['positions']
    k1, k2 = bbh.gaussian_kde2d(g['pred']), bbh.gaussian_kde2d(g['lal'])
    assert abs(k1.factor - 1500 ** (-1.0 / 6)) < 1e-15
    p1, p2 = k1.pdf(pos), k2(pos)
    assert np.abs(p1 - g['cnn_pdf']).max() < 1e-5 * g['cnn_pdf'].max()
    assert np.abs(p2 - g['lalinf_pdf']).max() < 1e-5 * g['lalinf_pdf'].max()
    # signal_pe.predict format: [mc (n,1), q (n,1)]; lalinference samples: [mc (m,), q (m,)]
    beta = bbh.overlap_beta([g['pred'][0][:, None], g['pred'][1][:, None]], [g['lal'][0], g['lal'][1]])
    assert abs(beta - float(g['beta'])) < 1e-5
    # combined-model format (n,2)
    assert abs(bbh.overlap_beta(g['pred'].T, [g['lal'][0], g['lal'][1]]) - float(g['beta'])) < 1e-5


@pytest.mark.parametrize('n,m', [(3, 1), (257, 1000), (4000, 10000)])
def test_kde_pdf_matches_oracle_ragged_sizes(bbh, n, m):
    rs = np.random.RandomState(n + m)
    data = np.stack([35.0 + 2.0 * rs.normal(size=n), 0.7 + 0.1 * rs.normal(size=n)])
    if n == 3:
        data = np.array([[30.0, 31.0, 30.5], [0.5, 0.9, 0.6]])
    pts = np.stack([rs.uniform(28, 42, m), rs.uniform(0.3, 1.1, m)])
    if n == 3:
        pts = np.array([[30.4], [0.62]])
    ref = st.gaussian_kde(data).pdf(pts)
    got = bbh.gaussian_kde2d(data).pdf(pts)
    assert got.shape == ref.shape and ref.max() > 0
    assert np.abs(got - ref).max() < 1e-5 * ref.max()


def test_overlap_properties_at_evaluation_size(bbh):
    """Sizes of the reference's evaluation stage (4000 generated samples against a lalinference posterior, 100 x 100
    grid): identical sets overlap completely, the score is symmetric and decreases with the offset."""
    rs = np.random.RandomState(3)
    a = [30 + rs.normal(size=4000), 0.8 + 0.05 * rs.normal(size=4000)]
    assert abs(bbh.overlap_beta(a, a) - 1.0) < 1e-6
    prev = 1.0
    for shift in (0.5, 1.5, 4.0):
        b = [a[0][:3000] + shift, a[1][:3000]]
        ab, ba = bbh.overlap_beta(a, b), bbh.overlap_beta(b, a)
        assert abs(ab - ba) < 1e-6 and ab < prev
        prev = ab


def test_kde_argument_validation(bbh):
    import torch
    from gennet_b200 import _lib
    with pytest.raises(ValueError):
        bbh.gaussian_kde2d(np.zeros((3, 10)))
    d = torch.zeros(8, device='cuda')
    with pytest.raises(_lib.GennetError):
        _lib.call('gn_kde2d_pdf_f32', _lib.ptr(d), 4, _lib.ptr(d), 4, 1.0, 2.0, 1.0, 1.0, _lib.ptr(d), _lib.stream())


def test_overlap_tests_triplet(bbh):
    """bbh.overlap_tests, the drop-in for bbhMahoGANy.py:811-871: SciPy K-S / Anderson-Darling + device beta."""
    import warnings
    g = np.load(GOLD)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ks, ad, beta = bbh.overlap_tests([g['pred'][0][:, None], g['pred'][1][:, None]], [g['lal'][0], g['lal'][1]], [30, 0.8])
    assert ks.shape == (2, 2) and len(ad) == 2 and abs(beta - float(g['beta'])) < 1e-5
