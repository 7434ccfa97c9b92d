"""Pin the synthesis oracle against vectors produced by the reference's own source
(tests/golden/make_golden.py -> synth_ref.npz) and against SciPy / closed forms."""
import random

import numpy as np
import pytest

from oracle import synth_oracle as so
from tests.golden.make_golden import toy_psd


def test_tukey_matches_reference_and_scipy(golden):
    from scipy.signal.windows import tukey as sp_tukey
    for key in [k for k in golden.files if k.startswith('tukey_')]:
        _, M, a = key.split('_')
        w = so.tukey(int(M), float(a))
        assert np.array_equal(w, golden[key])
        np.testing.assert_allclose(w, sp_tukey(int(M), float(a)), rtol=0, atol=1e-15)


def test_convert_beta(golden):
    cases = [([0.75, 0.95], 1024, 4), ([0.45, 0.55], 1024, 4), ([0.5, 0.5], 2048, 4), ([0.45, 0.55], 4096, 8)]
    got = np.array([so.convert_beta(*c) for c in cases])
    assert np.array_equal(got, golden['convert_beta'])


@pytest.mark.parametrize('fs', [1024, 2048])
def test_gen_noise_and_whiten_match_reference(golden, fs):
    T = 4
    psd = toy_psd(fs, T)
    rs = np.random.RandomState(1234 + fs)
    Nf = fs * T // 2 + 1
    normals = np.stack([rs.normal(0, 1, Nf), rs.normal(0, 1, Nf)])
    x = so.gen_noise(fs, T, psd, normals=normals)
    assert np.array_equal(x, golden['noise_td_%d' % fs])
    w = so.whiten_data(x, T, fs, psd, 'td')
    assert np.array_equal(w, golden['whiten_td_%d' % fs])
    wf = so.whiten_data(golden['fd_in_%d' % fs], T, fs, psd, 'fd')
    assert np.array_equal(wf, golden['whiten_fd_%d' % fs])
    # whitened coloured noise is ~unit variance away from the window taper
    assert abs(np.std(so.crop_central(w, fs, T)) - 1.0) < 0.08


def test_gen_masses_and_gen_par_follow_reference_rng_order(golden):
    for k, mdist in enumerate(['astro', 'hunt_constrain', 'gh', 'metric']):
        rs = np.random.RandomState(50 + k)
        rows = []
        for _ in range(8):
            m12, mc, eta = so.gen_masses(rs, 5.0, 100.0, mdist)
            rows.append([m12[0], m12[1], mc, eta])
        np.testing.assert_allclose(np.array(rows, dtype=np.float64), golden['gen_masses_' + mdist], rtol=1e-15)
    rs = np.random.RandomState(99)
    rows = []
    for _ in range(6):
        p = so.gen_par(rs, 1024, 4, mdist='hunt_constrain', beta=[0.45, 0.55])
        rows.append([p.mc, p.M, p.eta, p.m1, p.m2, p.idx])
    p = so.gen_par(rs, 1024, 4, mdist='hunt_constrain', beta=[0.45, 0.55], gw_tmp=True)
    rows.append([p.mc, p.M, p.eta, p.m1, p.m2, p.idx])
    np.testing.assert_allclose(np.array(rows, dtype=np.float64), golden['gen_par'], rtol=1e-15)


def test_burst_and_sinusoid_generators(golden):
    random.seed(5)
    draws = [(random.uniform(0.25, 0.75), random.uniform(1.0 / 60.0, 1.0 / 15.0)) for _ in range(6)]
    d, p = so.make_burst_waveforms(6, rand5=True, draws=draws)
    assert np.array_equal(d, golden['burst_data']) and np.array_equal(p, golden['burst_pars'])
    d1, _ = so.make_burst_waveforms(1)
    assert np.array_equal(d1, golden['burst_fixed'])
    rs = np.random.RandomState(8)
    draws = [(rs.random_sample(), rs.random_sample()) for _ in range(5)]
    assert np.array_equal(so.sample_data(draws), golden['nn_sample_data'])


def test_whiten_linearity_and_parseval():
    fs, T = 1024, 4
    psd = so.analytic_psd(fs, T)
    rs = np.random.RandomState(3)
    a, b = rs.normal(size=fs * T), rs.normal(size=fs * T)
    wa, wb = so.whiten_data(a, T, fs, psd), so.whiten_data(b, T, fs, psd)
    np.testing.assert_allclose(so.whiten_data(2 * a - 3 * b, T, fs, psd), 2 * wa - 3 * wb, rtol=1e-10, atol=1e-9 * np.abs(wa).max())


def test_gen_bbh_tail_pipeline_shapes():
    fs, T = 1024, 4
    psd = so.analytic_psd(fs, T)
    hp, hc = so.newtonian_chirp_fd(36.0, 29.0, fs, T)
    lo, hi = so.convert_beta([0.5, 0.5], fs, T)
    ts, ref_idx = so.gen_bbh_from_fd(hp, hc, fs, T, psd, lo, 0.3, -0.4)
    assert ts.shape == (fs * T,)
    peak = int(np.argmax(np.abs(ts)))
    assert abs(peak - lo) < 64          # peak lands at the requested index (within the 11-sample lead + envelope)
    assert np.all(ts[: fs // 2] == 0)   # aggressive window zeroes the safe margins


def test_resample_restatement_against_scipy():
    """oracle.resample_fft restates scipy 1.1.0's resample (requirements.txt:42).  Current SciPy differs only in the
    new-Nyquist bin (carried over since 1.4), so it must agree exactly on input band-limited below it, and the
    restatement must drop that bin on broadband input."""
    import scipy.signal
    from oracle import synth_oracle as so
    rs = np.random.RandomState(0)
    for Nx in (2000, 4096, 3001, 700):
        x = rs.normal(size=Nx)
        X = np.fft.rfft(x)
        X[200:] = 0
        xb = np.fft.irfft(X, Nx)
        assert np.abs(so.resample_fft(xb, 512) - scipy.signal.resample(xb, 512)).max() < 1e-12
        y = so.resample_fft(x, 512)
        assert abs(np.fft.fft(y)[256]) < 1e-9 * np.abs(np.fft.fft(y)).max()
    # upsampling branch (Nx < num) and the full ingest: max-normalised, rolled
    x = rs.normal(size=300)
    assert np.abs(so.resample_fft(x, 512)[::1].mean() - x.mean()) < 1e-9 * 512
    d = so.ingest_waveform(rs.normal(size=5000), -37)
    assert d.shape == (512,) and abs(d.max() - 1.0) < 1e-12
    assert np.argmax(d) == (np.argmax(np.roll(d, 37)) - 37) % 512
