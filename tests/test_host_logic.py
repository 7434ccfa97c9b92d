"""Host-side logic (graph executor, Keras protocol, trainable sets, arenas, optimizer wiring) checked on
the CPU against the float64 oracle, with the C ABI answered by tests/fake_backend.py."""
import numpy as np
import pytest

from tests import fake_backend, parity_cases as pc


@pytest.fixture
def fake(monkeypatch):
    return fake_backend.install(monkeypatch)


def test_pe_model_step_matches_oracle(fake):
    prod, orc, x, y = pc.pe_case(128, 4)
    out = prod.predict(x)
    assert isinstance(out, list) and out[0].shape == (4, 1) and out[1].shape == (4, 1)
    errs, w0 = pc.compare_step(prod, orc, x, y)
    pc.compare_weights(prod, orc, w0)
    # second step exercises Adam's t=2 bias correction and the persistent moments
    pc.resync([(prod, orc)])
    errs, w0 = pc.compare_step(prod, orc, x, y, check_predict=False)
    pc.compare_weights(prod, orc, w0)


def test_pe_returns_keras_shaped_list(fake):
    prod, orc, x, y = pc.pe_case(128, 4)
    r = prod.train_on_batch(x, y)
    assert len(r) == 5          # [total, mc_loss, q_loss, mc_acc, q_acc]
    assert abs(r[0] - (r[1] + r[2])) < 1e-6 * max(1.0, abs(r[0]))


def test_gan_wiring_matches_oracle(fake):
    (g, d, dg), (og, od, ocomp), z, sX, sy = pc.gan_case(64, 4)
    # generator.predict uses BN moving statistics and no dropout
    pc.assert_close(g.predict(z), og.predict(z), 'generator.predict')
    dw0 = [w.copy() for w in d.get_weights()]
    errs, w0 = pc.compare_step(d, od, sX, sy)
    pc.compare_weights(d, od, w0)
    assert any(np.abs(a - b).max() > 0 for a, b in zip(dw0, d.get_weights()))
    # G step through the frozen D: D's weights must not move, G's must, BN moving stats must update
    pc.resync([(d, od)])
    dw1 = [w.copy() for w in d.get_weights()]
    gw1 = [w.copy() for w in g.get_weights()]
    errs, w0 = pc.compare_step(dg, ocomp, z, [1] * 4, check_predict=False)
    assert all(np.array_equal(a, b) for a, b in zip(dw1, d.get_weights()))
    assert any(not np.array_equal(a, b) for a, b in zip(gw1, g.get_weights()))
    pc.compare_weights(g, og, [w for w in w0[:len(gw1)]])


@pytest.fixture
def f16x2(fake):
    from gennet_b200 import nn
    nn.set_compute_dtype('f16x2')
    try:
        yield nn
    finally:
        nn.set_compute_dtype('float32')


def test_f16x2_mode_wiring_pe(f16x2):
    """The benchmark's arithmetic mode on the host side: operand planes with their scale scalars travel between layers
    (producer-computed max |x|, planes of weights per weight version), bias gradients come out of the data-gradient
    column sums -- checked at rtol 1e-4 against the oracle with the C ABI answered by the CPU stand-in."""
    nn = f16x2
    prod, orc, x, y = pc.pe_case(128, 4)
    assert sorted(l._path() for l in prod.all_layers() if isinstance(l, nn.Conv1D)) == ['smallcin32'] * 2 + ['tc3'] * 7
    errs, w0 = pc.compare_step(prod, orc, x, y)
    pc.compare_weights(prod, orc, w0)
    pc.resync([(prod, orc)])
    errs, w0 = pc.compare_step(prod, orc, x, y, check_predict=False)
    pc.compare_weights(prod, orc, w0)


def test_f16x2_mode_wiring_gan(f16x2):
    """Generator / discriminator in the benchmark's mode: BatchNormalization statistics taken from the convolution that
    feeds it, operand planes written by the BN-tanh-dropout apply pass under the a-priori bound, bias gradients from the
    gradient split, dropout applying the fused LeakyReLU mask -- D step and G step through the frozen D."""
    nn = f16x2
    (g, d, dg), (og, od, ocomp), z, sX, sy = pc.gan_case(64, 4)
    bns = [l for l in g.all_layers() if isinstance(l, nn.BatchNormalization)]
    assert sum(l.planes_consumer is not None for l in bns) >= 4
    assert sum(getattr(l, 'bn_consumer', None) is not None for l in g.all_layers() if isinstance(l, nn.Conv1D)) >= 4
    pc.assert_close(g.predict(z), og.predict(z), 'generator.predict')
    errs, w0 = pc.compare_step(d, od, sX, sy)
    pc.compare_weights(d, od, w0)
    pc.resync([(d, od)])
    dw1 = [w.copy() for w in d.get_weights()]
    errs, w0 = pc.compare_step(dg, ocomp, z, [1] * 4, check_predict=False)
    assert all(np.array_equal(a, b) for a, b in zip(dw1, d.get_weights()))
    pc.compare_weights(g, og, w0[:len(g.get_weights())])


def test_burst_three_step_iteration(fake):
    (g, d, dg, sub_g), (og, od, ocomp, osub), z, sX, sy, ny = pc.burst_case(64, 4)
    errs, w0 = pc.compare_step(d, od, sX, sy)
    pc.compare_weights(d, od, w0)
    pc.resync([(d, od)])
    errs, w0 = pc.compare_step(sub_g, osub, z, ny, check_predict=False)     # MSE on batch-global residual moments
    pc.compare_weights(g, og, w0[:len(g.get_weights())])
    pc.resync([(g, og)])
    errs, w0 = pc.compare_step(dg, ocomp, z, [1] * 4, check_predict=False)
    pc.compare_weights(g, og, w0[:len(g.get_weights())])


def test_wvf_functional_models(fake):
    (G, D, GAN), (og, od, ogan), X, y, z, yz = pc.wvf_case(256, 4)
    pc.assert_close(G.predict(z), og.predict(z), 'G.predict')
    errs, w0 = pc.compare_step(D, od, X, y)
    pc.compare_weights(D, od, w0)
    pc.resync([(D, od)])
    errs, w0 = pc.compare_step(GAN, ogan, z, yz, check_predict=False)
    pc.compare_weights(G, og, w0[:len(G.get_weights())])


def test_upsampling_is_fused_into_following_conv(fake):
    from gennet_b200 import nn, bbh
    nn.clear_session()
    bbh.n_pix = 64
    g = bbh.generator_model()
    ups = [l for l in g.layers if isinstance(l, nn.UpSampling1D)]
    convs = [l for l in g.layers if isinstance(l, nn.Conv1D)]
    assert len(ups) == 2 and all(u.fused for u in ups)
    assert [c.fused_up for c in convs] == [2, 2, 1, 1, 1, 1]
    assert g.output_shape == (64, 1)


def test_param_arena_is_contiguous_and_shared(fake):
    (g, d, dg), _, z, sX, sy = pc.gan_case(64, 2)
    segs_g = dg._compiled['segments']
    assert len(segs_g) == 1, 'all generator parameters should form one flat segment'
    n_train = sum(p.numel() for l in g.all_layers() for p in l.params if p.trainable)
    assert segs_g[0][1].numel() >= n_train
    # the discriminator's parameters live in a different arena and form one segment of their own
    assert len(d._compiled['segments']) == 1
    assert d._compiled['segments'][0][0] != segs_g[0][0]


def test_summary_and_counts(fake):
    from gennet_b200 import nn, bbh
    nn.clear_session()
    bbh.n_pix = 1024
    pe = bbh.signal_pe_model()
    assert abs(pe.count_params() - 4.63e6) < 0.02e6
    lines = []
    pe.summary(print_fn=lines.append)
    assert any('Total params' in l for l in lines)


def test_waveform_ingest_host_logic(fake, tmp_path):
    """load_txtwfs.py:31-77 mirror: files -> resample operator -> max-normalise -> roll, grouped by input length."""
    from gennet_b200 import wvf
    from oracle import synth_oracle as so
    rs = np.random.RandomState(2)
    series = [rs.normal(size=n) for n in (3000, 3000, 2048)]
    for i, s_ in enumerate(series):
        np.savetxt(str(tmp_path / ('wf%d.txt' % i)), s_)
    data, pars = wvf.load_data(str(tmp_path), 10, frequencies=[70.0, 70.0, 70.0], rng=np.random.RandomState(4))
    assert data.shape == (3, 512) and pars.shape == (3, 2)
    import glob
    files = list(glob.iglob('%s/*.txt' % str(tmp_path)))
    for row, p, f in zip(data, pars, files):
        ref = so.ingest_waveform(np.loadtxt(f), int(p[0] - 256))
        assert np.abs(row - ref).max() < 5e-5          # float32 operator in the stand-in backend
        assert p[1] == 70.0 and -100 <= p[0] - 256 < 100


def test_two_model_version_wiring(fake):
    """BASELINE config 5: Conv2DTranspose generator + Conv1D discriminator, D step then G step through the frozen D."""
    (G, D, GAN), (og, od, ogan), X, y, z, yz = pc.two_model_case(8)
    assert G.output_shape == (50,) and G.count_params() == sum(int(np.prod(w.shape)) for w in og.get_weights())
    widths = [l.output_shape[1] for l in G.layers if type(l).__name__ == 'Conv2DTranspose']
    assert widths == [4, 11, 26, 57]                              # no_mode_collapse_network.py:79-90
    pc.assert_close(G.predict(z), og.predict(z), 'G.predict')
    errs, w0 = pc.compare_step(D, od, X, y)
    pc.compare_weights(D, od, w0)
    pc.resync([(D, od)])
    dw = [w.copy() for w in D.get_weights()]
    errs, w0 = pc.compare_step(GAN, ogan, z, yz, check_predict=False)
    assert all(np.array_equal(a, b) for a, b in zip(dw, D.get_weights()))
    pc.compare_weights(G, og, w0[:len(G.get_weights())])


def test_overlap_beta_host_algebra_matches_scipy_golden(fake):
    """bbh.gaussian_kde2d / overlap_beta: Scott factor, covariance, normalisation, centring and the comparison grid are
    host code; with the two device entry points answered by the CPU stand-in the score must equal SciPy's."""
    import os
    from gennet_b200 import bbh
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'kde_overlap.npz'))
    k = bbh.gaussian_kde2d(g['pred'])
    assert abs(k.factor - 1500 ** (-1.0 / 6)) < 1e-15
    p = k.pdf(g['positions'])
    assert np.abs(p - g['cnn_pdf']).max() < 2e-6 * g['cnn_pdf'].max()          # positions pass through float32
    beta = bbh.overlap_beta([g['pred'][0][:, None], g['pred'][1][:, None]], [g['lal'][0], g['lal'][1]])
    assert abs(beta - float(g['beta'])) < 1e-6


def test_overlap_tests_mirrors_the_reference_triplet(fake):
    """bbh.overlap_tests (bbhMahoGANy.py:811-871): K-S and Anderson-Darling are SciPy's, beta is the device path."""
    import os
    import warnings
    from scipy.stats import anderson_ksamp, ks_2samp
    from gennet_b200 import bbh
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'kde_overlap.npz'))
    pred, lal = g['pred'], g['lal']
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ks, ad, beta = bbh.overlap_tests([pred[0][:, None], pred[1][:, None]], [lal[0], lal[1]], true_vals=[30.0, 0.8])
        ref_ad = anderson_ksamp([pred[1], lal[1]])
    assert ks.shape == (2, 2) and np.allclose(ks[0], ks_2samp(pred[0], lal[0]))
    assert len(ad) == 2 and ad[1].statistic == ref_ad.statistic
    assert abs(beta - float(g['beta'])) < 1e-6
