"""Golden vectors for the evaluation-stage statistics: SciPy's own gaussian_kde (the reference's dependency,
bbhMahoGANy.py:62,790) on seeded sample sets, and the overlap score of overlap_tests (:853-870) computed with it.
Run in the build container:  python tests/golden/make_kde_golden.py  ->  tests/golden/kde_overlap.npz"""
import os

import numpy as np
import scipy
from scipy.stats import gaussian_kde

rs = np.random.RandomState(2024)
# CNN-like and lalinference-like (chirp mass, mass ratio) posteriors: correlated, different widths and offsets
pred = np.stack([30.0 + 1.5 * rs.normal(size=1500), 0.80 + 0.08 * rs.normal(size=1500)])
pred[1] += 0.03 * (pred[0] - 30.0)
lal = np.stack([30.6 + 1.1 * rs.normal(size=900), 0.84 + 0.05 * rs.normal(size=900)])
lal[1] -= 0.02 * (lal[0] - 30.6)
comb_mc, comb_q = np.concatenate((pred[0], lal[0])), np.concatenate((pred[1], lal[1]))
X, Y = np.mgrid[np.min(comb_mc):np.max(comb_mc):100j, np.min(comb_q):np.max(comb_q):100j]
positions = np.vstack([X.ravel(), Y.ravel()])
k1, k2 = gaussian_kde(pred), gaussian_kde(lal)
p1, p2 = k1.pdf(positions), k2.pdf(positions)
beta = np.sum(p1 * p2) / np.sqrt(np.sum(p1 ** 2) * np.sum(p2 ** 2))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'kde_overlap.npz')
np.savez_compressed(out, pred=pred, lal=lal, positions=positions.astype(np.float64), cnn_pdf=p1, lalinf_pdf=p2,
                    beta=beta, scipy_version=scipy.__version__)
print(out, 'beta = %.12f' % beta, 'scipy', scipy.__version__)
