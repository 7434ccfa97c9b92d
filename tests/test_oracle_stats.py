"""The evaluation-stage oracle (oracle/stats_oracle.py) against SciPy's gaussian_kde -- the reference's own dependency
(bbhMahoGANy.py:62,790) -- live and through the committed golden vectors."""
import os

import numpy as np
import pytest

from oracle import stats_oracle as st

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'kde_overlap.npz')


def test_kde_oracle_matches_installed_scipy():
    scipy_stats = pytest.importorskip('scipy.stats')
    rs = np.random.RandomState(5)
    data = np.stack([rs.normal(3.0, 2.0, 400), rs.normal(-1.0, 0.3, 400)])
    data[1] += 0.1 * data[0]
    pts = np.stack([rs.uniform(-4, 10, 300), rs.uniform(-3, 1, 300)])
    ref = scipy_stats.gaussian_kde(data)
    mine = st.gaussian_kde(data)
    assert abs(mine.factor - ref.factor) < 1e-15
    assert np.allclose(mine.covariance, ref.covariance, rtol=1e-13, atol=0)
    assert np.allclose(mine.pdf(pts), ref.pdf(pts), rtol=1e-11, atol=1e-300)


def test_kde_oracle_matches_golden_vectors():
    g = np.load(GOLD)
    pos = g['positions']
    assert np.allclose(st.gaussian_kde(g['pred']).pdf(pos), g['cnn_pdf'], rtol=1e-11, atol=1e-300)
    assert np.allclose(st.gaussian_kde(g['lal']).pdf(pos), g['lalinf_pdf'], rtol=1e-11, atol=1e-300)
    assert abs(st.overlap_beta(g['pred'], g['lal']) - float(g['beta'])) < 1e-12


def test_overlap_is_one_for_identical_sets_and_small_for_distant_ones():
    rs = np.random.RandomState(1)
    a = np.stack([rs.normal(size=300), rs.normal(size=300)])
    assert abs(st.overlap_beta(a, a) - 1.0) < 1e-12
    b = a + np.array([[25.0], [25.0]])
    assert st.overlap_beta(a, b) < 1e-6
