"""Host logic of the 2_model_version scripts (gennet_b200.two_model.*) on the CPU: sampling helpers reproduce the
reference's shapes, label layouts and RNG call order; the subtract-stage networks (ELU, l1 / l2 regularisers, Dropout;
GaussianNoise, BatchNormalization(axis=1), global average pooling) step like the float64 oracle (C ABI answered by
tests/fake_backend.py); the shipped best_*_weights.hdf5 load into the builders."""
import os

import numpy as np
import pytest

from tests import fake_backend, parity_cases as pc

REF_DIR = '/root/reference/2_model_version/weight_version'


@pytest.fixture
def fake(monkeypatch):
    return fake_backend.install(monkeypatch)


def test_subtract_stage_steps_match_oracle(fake):
    (G, D, GAN), (og, od, ogan), X, y, z, yz = pc.subtract_case(6)
    pc.assert_close(G.predict(z), og.predict(z), 'G.predict')
    errs, w0 = pc.compare_step(D, od, X, y)
    pc.compare_weights(D, od, w0)
    pc.resync([(D, od)])
    dw = [w.copy() for w in D.get_weights()]
    # the generator step carries the l1 activity / l2 kernel penalties of its first Conv2DTranspose in loss AND gradient
    errs, w0 = pc.compare_step(GAN, ogan, z, yz, check_predict=False)
    assert all(np.array_equal(a, b) for a, b in zip(dw, D.get_weights()))
    pc.compare_weights(G, og, w0[:len(G.get_weights())])


def test_regularisation_terms_enter_loss_and_gradient(fake):
    """Reported loss = BCE + 0.001 sum|y1| + 0.01 sum w1^2 (keras `model.losses`); without them it is plain BCE."""
    (G, D, GAN), (og, od, ogan), X, y, z, yz = pc.subtract_case(4)
    conv = [l for l in G.all_layers() if type(l).__name__ == 'Conv2DTranspose'][0]
    assert conv.kernel_regularizer.l2 == 0.01 and conv.activity_regularizer.l1 == 0.001
    w = conv.get_weights()[0].astype(np.float64)
    with_reg = GAN.train_on_batch(z, yz)
    pc.resync([(G, og)])
    conv.kernel_regularizer = conv.activity_regularizer = None
    plain = GAN.train_on_batch(z, yz)
    assert with_reg - plain > 0.01 * (w ** 2).sum() * 0.999          # kernel term alone is a lower bound of the gap


def test_nw_discriminator_step_matches_oracle(fake):
    D, od, X, y = pc.nw_disc_case(8)
    assert [type(l).__name__ for l in D.layers].count('BatchNormalization') == 3
    assert all(l.axis == 1 for l in D.layers if type(l).__name__ == 'BatchNormalization')
    pc.assert_close(D.predict(X), od.predict(X), 'D.predict')
    errs, w0 = pc.compare_step(D, od, X, y)
    pc.compare_weights(D, od, w0)
    r = D.train_on_batch(X, y)
    assert len(r) == 2 and 0.0 <= r[1] <= 1.0                        # [loss, categorical accuracy]


def test_nw_generator_fails_like_keras_2_2_4(fake):
    from gennet_b200 import nn
    from gennet_b200.two_model import subtract_model_nw as nw
    nn.clear_session()
    with pytest.raises(AssertionError, match='dilation_rate'):
        nw.get_generative(nn.Input(shape=(1, nw.hyperparams.noise_dim)), lr=1e-4)


def test_sampling_helpers_follow_the_scripts(fake):
    from gennet_b200 import nn
    from gennet_b200.two_model import noise_gan as ng, subtract_model as sm, subtract_model_nw as nw
    nn.clear_session()
    nn.set_seed(3)
    sm.hyperparams.noise_dim, sm.hyperparams.outdim = 1, 50
    G, _ = sm.get_generative(nn.Input(shape=(1, 1)), lr=1e-4)
    rs = np.random.RandomState(5)
    ht = sm.sample_data(1, rng=rs)
    assert ht.shape == (1, 50) and np.abs(ht).max() <= 5.0
    xt = ht + rs.normal(0, 5, size=[1, 50])
    # weight_version: all-ones labels, residual rows = x_t - G(z), RNG order XT -> latents
    rs2 = np.random.RandomState(9)
    X, y = sm.sample_data_and_gen(G, xt, [], noise_dim=1, n_samples=6, noise_samples=4, rng=rs2)
    rs3 = np.random.RandomState(9)
    XT = rs3.normal(0, 5, size=[6, 50])
    zz = rs3.normal(0, 1, size=[4, 1, 1])
    assert X.shape == (10, 50) and (y == 1).all() and np.allclose(X[:6], XT)
    assert np.allclose(X[6:], xt[0] - G.predict(zz), atol=1e-5)
    Xn, yn = sm.sample_noise(G, xt, [], noise_dim=1, n_samples=4, rng=rs2)
    assert Xn.shape == (4, 1, 1) and (yn == 1).all()
    Xt, res = sm.test_data_and_gen(G, xt, [], noise_dim=1, n_samples=3, noise_samples=2, rng=rs2)
    assert Xt.shape == (5, 50) and res.shape == (2, 1, 50)
    # noise_gan: noise rows then generated rows, one-hot labels
    X, y = ng.sample_data_and_gen(G, noise_dim=1, n_samples=5, noise_samples=3, rng=np.random.RandomState(1))
    assert X.shape == (8, 50) and (y[:5, 1] == 1).all() and (y[5:, 0] == 1).all() and y.sum() == 8
    # no_weight_code: half batches, soft labels drawn once per group, residual = G(z) - x_t
    nw.hyperparams.batch_size = 4
    X, y = nw.sample_data_and_gen(G, xt, [], 0, noise_dim=1, rng=np.random.RandomState(2))
    assert X.shape == (4, 50) and y.shape == (4, 2)
    assert 0.7 <= y[0, 0] <= 1 and y[0, 0] == y[1, 0] and 0 <= y[0, 1] <= 0.3 and 0.7 <= y[2, 1] <= 1
    Xn, yn = nw.sample_noise(G, xt, [], noise_dim=1)
    assert Xn.shape == (4, 1, 1) and (yn[:, 0] == 1).all() and (yn[:, 1] == 0).all()


def test_subtract_stage_training_loop_runs_and_saves(fake, tmp_path):
    """noise_gan.train saves best_d_weights.hdf5 / d_model.hdf5; subtract_model.main loads such files and trains."""
    from gennet_b200 import nn
    from gennet_b200.two_model import noise_gan as ng, subtract_model as sm
    nn.clear_session()
    nn.set_seed(4)
    G, _ = ng.get_generative(nn.Input(shape=(1, 1)), lr=1e-3)
    D, _ = ng.get_discriminative(nn.Input(shape=(50,)), lr=1e-3)
    GAN, _ = ng.make_gan(nn.Input((1, 1)), G, D)
    rs = np.random.RandomState(0)
    ng.pretrain(G, D, n_samples=8, noise_samples=8, noise_dim=1, batch_size=4, rng=rs)
    d_loss, g_loss = ng.train(GAN, G, D, epochs=2, n_samples=6, noise_samples=6, noise_dim=1, rng=rs, save_to=str(tmp_path))
    assert len(d_loss) == 2 and len(g_loss) == 2 and np.isfinite(d_loss + g_loss).all()
    assert os.path.exists(tmp_path / 'best_d_weights.hdf5') and os.path.exists(tmp_path / 'd_model.hdf5')
    D2 = nn.load_model(str(tmp_path / 'd_model.hdf5'))
    x = rs.normal(size=(3, 50)).astype(np.float32)
    assert np.allclose(D2.predict(x), D.predict(x), atol=1e-6)
    sm.hyperparams.n_samples = sm.hyperparams.noise_samples = 6
    sm.hyperparams.batch_size, sm.hyperparams.noise_dim, sm.hyperparams.outdim = 4, 1, 50
    out = sm.main(epochs=2, rng=np.random.RandomState(1))
    assert len(out['d_loss']) == 2 and np.isfinite(out['g_loss']).all() and out['residuals'].shape == (25, 1, 50)


@pytest.mark.skipif(not os.path.isdir(REF_DIR), reason='reference checkout not present')
def test_shipped_best_weights_load_into_the_subtract_builders(fake):
    """subtract_model.py:312-320: G.load_weights('best_g_weights.hdf5'), D.load_weights('best_d_weights.hdf5')."""
    from gennet_b200 import nn, hdf5
    from gennet_b200.two_model import subtract_model as sm, noise_gan as ng
    nn.clear_session()
    G, _ = sm.get_generative(nn.Input(shape=(1, 1)), lr=1e-4)
    G.load_weights(os.path.join(REF_DIR, 'best_g_weights.hdf5'))
    f = hdf5.File(os.path.join(REF_DIR, 'best_g_weights.hdf5'))
    names = [n.decode() for n in f.attrs['layer_names']]
    first = [n for n in names if n.startswith('conv2d_transpose')][0]
    k = np.asarray(f[first][first + '/kernel:0'][...])
    conv = [l for l in G.layers if type(l).__name__ == 'Conv2DTranspose'][0]
    assert np.array_equal(conv.get_weights()[0], k)
    D, _ = ng.get_discriminative(nn.Input(shape=(50,)), lr=1e-4)
    D.load_weights(os.path.join(REF_DIR, 'best_d_weights.hdf5'))
    assert np.isfinite(D.predict(np.zeros((2, 50), np.float32))).all()
