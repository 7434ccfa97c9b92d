"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy float64) of the evaluation-stage statistics of the reference.

The reference calls ``scipy.stats.gaussian_kde`` (a third-party dependency; the repo pins SciPy 1.1.0 in
requirements.txt) at BBH_version/bbhMahoGANy.py:790 and evaluates it at :861,866; the overlap score is :868-870.
``gaussian_kde`` below restates SciPy's published algorithm (scipy/stats/kde.py, 1.1.0: Scott's factor
n**(-1/(d+4)), covariance = unbiased data covariance * factor**2, normalisation sqrt(det(2 pi covariance)) * n,
``evaluate`` = sum of Gaussians / normalisation).  Pinned: tests/test_oracle_stats.py checks it against the SciPy
installed in this image (the same definition in every SciPy release since 0.11) and against the committed vectors
tests/golden/kde_overlap.npz made by tests/golden/make_kde_golden.py with SciPy itself.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""
import numpy as np


class gaussian_kde(object):
    """scipy.stats.gaussian_kde with the default (Scott) bandwidth; dataset (d, n)."""

    def __init__(self, dataset):
        self.dataset = np.atleast_2d(np.asarray(dataset, dtype=np.float64))
        self.d, self.n = self.dataset.shape
        self.factor = np.power(self.n, -1.0 / (self.d + 4))          # scotts_factor
        self._data_covariance = np.atleast_2d(np.cov(self.dataset, rowvar=1, bias=False))
        self._data_inv_cov = np.linalg.inv(self._data_covariance)
        self.covariance = self._data_covariance * self.factor ** 2
        self.inv_cov = self._data_inv_cov / self.factor ** 2
        self._norm_factor = np.sqrt(np.linalg.det(2 * np.pi * self.covariance)) * self.n

    def evaluate(self, points):
        points = np.atleast_2d(np.asarray(points, dtype=np.float64))
        d, m = points.shape
        assert d == self.d
        result = np.zeros((m,), dtype=np.float64)
        for i in range(self.n):                                      # kde.py evaluate(), m >= n branch
            diff = self.dataset[:, i, np.newaxis] - points
            tdiff = np.dot(self.inv_cov, diff)
            energy = np.sum(diff * tdiff, axis=0) / 2.0
            result = result + np.exp(-energy)
        return result / self._norm_factor

    pdf = evaluate
    __call__ = evaluate


def overlap_beta(pred_xy, lalinf_xy, n_grid=100):
    """bbhMahoGANy.py:853-870 with both kernels built from the sample sets themselves (:790): pred_xy, lalinf_xy are
    (2, n) / (2, m) arrays of (chirp mass, mass ratio) samples."""
    pred_xy, lalinf_xy = np.asarray(pred_xy, np.float64), np.asarray(lalinf_xy, np.float64)
    comb_mc = np.concatenate((pred_xy[0], lalinf_xy[0]))
    comb_q = np.concatenate((pred_xy[1], lalinf_xy[1]))
    X, Y = np.mgrid[np.min(comb_mc):np.max(comb_mc):complex(0, n_grid), np.min(comb_q):np.max(comb_q):complex(0, n_grid)]
    positions = np.vstack([X.ravel(), Y.ravel()])
    cnn_pdf = gaussian_kde(pred_xy).pdf(positions)
    lalinf_pdf = gaussian_kde(lalinf_xy).pdf(positions)
    return np.divide(np.sum(cnn_pdf * lalinf_pdf), np.sqrt(np.sum(cnn_pdf ** 2) * np.sum(lalinf_pdf ** 2)))
