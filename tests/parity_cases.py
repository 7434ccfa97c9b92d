"""Shared parity scaffolding: build the same network in the product (gennet_b200) and in the float64
oracle with identical weights, inputs and dropout draws, run one Keras-style step in both and compare
forward outputs, losses, gradients, updated weights and BatchNorm moving statistics.

Used by the CPU host-logic tests (product routed to tests/fake_backend.py) and by the `-m gpu` parity
tests (product on the real sm_100a kernels).  Tolerance: rtol 1e-4 of the tensor's max magnitude (float32
product vs float64 oracle), as BASELINE.json's north_star states.
"""
import os

import numpy as np
import torch

from oracle import keras_oracle as ko

RTOL = 1e-4


def case_seed(case):
    """Seed of a parametrised test case that is the same in every process (hash() of a tuple holding a string is salted
    per interpreter run, which made the test data -- and the distance to a tolerance -- differ from run to run)."""
    import zlib
    return zlib.crc32(repr(case).encode()) & 0x7fffffff


def assert_close(got, ref, what, rtol=RTOL, floor=1e-30):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, '%s: shape %s vs %s' % (what, got.shape, ref.shape)
    scale = max(np.abs(ref).max(), floor)
    err = np.abs(got - ref).max() / scale
    assert np.isfinite(got).all(), '%s: non-finite values' % what
    if os.environ.get('GN_TEST_MARGINS') and err > 0.5 * rtol:      # developer aid: which checks sit near their tolerance
        print('[margin] %s: %.3e of rtol %.1e' % (what, err, rtol))
    assert err <= rtol, '%s: max error %.3e of scale %.3e exceeds rtol %.1e' % (what, err, scale, rtol)
    return err


def noise_layers(model):
    return [l for l in model.all_layers() if type(l).__name__ in ('Dropout', 'GaussianDropout', 'GaussianNoise')]


def sync_weights(product, oracle):
    ws = oracle.get_weights()
    product.set_weights([w.astype(np.float32) for w in ws])
    # the oracle continues from the float32-rounded values so both start identically
    oracle.set_weights([w.astype(np.float32).astype(np.float64) for w in ws])


def resync(pairs):
    """Copy the oracle's (float32-rounded) weights back into the product so that consecutive steps are
    compared independently: Adam's first updates are ~lr*sign(g), which amplifies float32-level gradient
    differences on near-zero gradients into weight differences that would otherwise compound."""
    for product, oracle in pairs:
        sync_weights(product, oracle)


def draw_noise(oracle_model, x, seed):
    """Run the oracle forward once in training mode to draw every dropout tensor; returns {oracle_name: tensor}."""
    noise = {'__gen__': torch.Generator().manual_seed(seed)}
    st = [([s.clone() for s in l.state], getattr(l, 'biased', None), getattr(l, 'local_step', 0))
          for l in oracle_model.all_layers()]
    with torch.no_grad():
        oracle_model.forward(torch.as_tensor(np.asarray(x, dtype=np.float64)), True, noise)
    for l, (s, biased, step) in zip(oracle_model.all_layers(), st):      # undo the BN moving-average side effect
        l.state = s
        if hasattr(l, 'biased'):
            l.biased, l.local_step = biased, step
    noise.pop('__gen__')
    noise.pop('__losses__', None)
    return noise


def kink_layers(model):
    """Layers with a piecewise-linear response (ReLU family, max pooling), in all_layers order."""
    out = []
    for l in model.all_layers():
        n = type(l).__name__
        act = getattr(l, 'activation', None) or getattr(l, 'act', None)
        if n in ('ReLU', 'LeakyReLU', 'MaxPooling1D') or (n in ('Activation', 'Dense', 'Conv1D') and act == 'relu'):
            out.append(l)
    return out


def record_kinks(product_model):
    """Wrap the product's piecewise-linear layers so that their outputs of the NEXT forward are kept."""
    rec = {}
    for l in kink_layers(product_model):
        if getattr(l, '_pc_wrapped', False):
            l._pc_rec = rec
            continue
        orig = l.forward

        def fw(x, ctx, _l=l, _orig=orig):
            y = _orig(x, ctx)
            if type(_l).__name__ == 'MaxPooling1D':
                # the window element the product selected: the first one equal to the pooled output (gn_maxpool1d_bwd_f32)
                xv = x.detach().float().cpu().numpy()
                B, L, C = xv.shape
                n = L // _l.pool
                win = xv[:, :n * _l.pool].reshape(B, n, _l.pool, C)
                _l._pc_rec[_l.name] = {'pool_idx': (win == win.max(axis=2, keepdims=True)).argmax(axis=2)}
            else:
                _l._pc_rec[_l.name] = y.detach().float().cpu().numpy().copy()
            return y
        l.forward = fw
        l._pc_wrapped = True
        l._pc_rec = rec
    return rec


def map_kinks(rec, oracle_model, product_model):
    ok, pk = kink_layers(oracle_model), kink_layers(product_model)
    assert len(ok) == len(pk), (len(ok), len(pk))
    return {o.name: rec[p.name] for o, p in zip(ok, pk) if p.name in rec}


def map_noise(noise, oracle_model, product_model):
    on, pn = noise_layers(oracle_model), noise_layers(product_model)
    assert len(on) == len(pn)
    return {p.name: noise[o.name].numpy().astype(np.float32) for o, p in zip(on, pn) if o.name in noise}


def compare_step(product, oracle, x, y, seed=0, rtol=RTOL, check_predict=True):
    """One train_on_batch in both; returns dict of observed relative errors."""
    errs = {}
    if check_predict:
        po, oo = product.predict(x), oracle.predict(x)
        if isinstance(oo, list):
            for k, (a, b) in enumerate(zip(po, oo)):
                errs['predict%d' % k] = assert_close(a, b, 'predict[%d]' % k, rtol)
        else:
            errs['predict'] = assert_close(po, oo, 'predict', rtol)
    noise = draw_noise(oracle, x, seed)
    pnoise = map_noise(noise, oracle, product)
    w_before = [w.copy() for w in oracle.get_weights()]
    # the product runs first; the side it takes at every ReLU-type kink is fed to the oracle (which checks
    # that any disagreement is within rounding distance of the kink) so both differentiate the same piece
    rec = record_kinks(product)
    rp = product.train_on_batch(x, y, _noise=pnoise)
    onoise = dict(noise)
    onoise['__kinks__'] = map_kinks(rec, oracle, product)
    ro = oracle.train_on_batch(x, y, noise=onoise)
    if not isinstance(rp, list):          # compiled without metrics: Keras returns the scalar loss
        ro = ro[0]
    errs['loss'] = assert_close(rp, ro, 'train_on_batch return', max(rtol, 2e-4))
    gp = product.get_gradients()
    # structurally-zero gradients (e.g. a bias feeding BatchNorm) are compared against the largest
    # gradient of the step instead of their own rounding noise
    gfloor = 1e-3 * max(np.abs(b).max() for b in oracle.last_grads)
    assert len(gp) == len(oracle.last_grads)
    bad = []
    gmax = 1e3 * gfloor
    for i, (a, b) in enumerate(zip(gp, oracle.last_grads)):
        scale = max(np.abs(b).max(), gfloor)
        if np.abs(b).max() < 1e-9 * gmax:
            # exactly zero in exact arithmetic (the bias of a layer feeding BatchNormalization): what the product holds
            # is the float32 rounding residue of a sum over batch x length terms; bound it against the step's largest gradient
            scale = 1e-2 * gmax
        emax = np.abs(a.astype(np.float64) - b).max() / scale
        errs['grad%d' % i] = emax
        if not (np.isfinite(a).all() and emax <= rtol):
            bad.append('gradient %d %s: max err %.3e (scale %.3e)' % (i, b.shape, emax, scale))
    assert not bad, '; '.join(bad)
    return errs, w_before


def compare_weights(product, oracle, w_before, rtol=RTOL):
    """Updated weights: compare the UPDATE (w_after - w_before) so that a no-op cannot pass."""
    pw, ow = product.get_weights(), oracle.get_weights()
    gmax = max(np.abs(b - w0).max() for b, w0 in zip(ow, w_before))
    for i, (a, b, w0) in enumerate(zip(pw, ow, w_before)):
        if np.abs(b - w0).max() == 0:
            assert np.abs(a - w0).max() == 0, 'weight %d should be unchanged' % i
            continue
        # float32 storage rounds each weight to ~6e-8 relative, and Adam's m/(sqrt(v)+eps) amplifies the
        # relative error of gradients that are small against eps: allow 5e-3 of the update scale + 2 ulp
        # (Keras Adam step 1: update = lr*g/(|g| + eps/sqrt(1-beta2)), i.e. slope lr/3e-6 near g = 0.)
        # The kernel itself is pinned against the closed form in test_adam_and_sgd_kernels; here the check
        # is on wiring: relative L2 error of the update <= 2e-2 and no element off by more than 25 %.
        upd = np.abs(b - w0).max()
        d = a.astype(np.float64) - b
        err = np.abs(d).max()
        el2 = np.linalg.norm(d.ravel()) / max(np.linalg.norm((b - w0).ravel()), 1e-30)
        bound = 0.25 * upd + 2.4e-7 * np.abs(b).max() + 1e-3 * gmax   # last term: structurally-zero gradients
        assert err <= bound and (el2 <= 2e-2 or upd < 1e-3 * gmax), \
            'weight update %d %s: max error %.3e (bound %.3e), L2 %.3e (update scale %.3e)' % (
                i, b.shape, err, bound, el2, upd)


# ---- case builders ----------------------------------------------------------------------------------
def pe_case(n_pix, B, seed=0):
    from gennet_b200 import nn, bbh
    nn.clear_session()
    ko.clear_session()
    nn.set_seed(seed)
    bbh.n_pix = n_pix
    prod = bbh.signal_pe_model()
    prod.compile(loss='mean_squared_error', optimizer=nn.Adam(lr=9e-5, beta_1=0.5), metrics=['accuracy'])
    orc = ko.build(ko.bbh_signal_pe_model(n_pix), seed=seed + 1)
    # positive head biases keep the ReLU heads active so the gradient check is not vacuous
    orc.branches[0][-2].weights[1].data += 0.5
    orc.branches[1][-2].weights[1].data += 0.5
    orc.compile('mean_squared_error', ko.Adam(9e-5, beta_1=0.5))
    sync_weights(prod, orc)
    rs = np.random.RandomState(seed)
    x = rs.normal(size=(B, n_pix, 1)).astype(np.float32)
    y = [rs.uniform(0.2, 1.0, B).astype(np.float32), rs.uniform(0.5, 1.0, B).astype(np.float32)]
    return prod, orc, x, y


def gan_case(n_pix, B, seed=0):
    """bbhMahoGANy.py GAN wiring: returns product and oracle (G, D, D_on_G) + inputs."""
    from gennet_b200 import nn, bbh
    nn.clear_session()
    ko.clear_session()
    nn.set_seed(seed)
    bbh.n_pix = n_pix
    rs = np.random.RandomState(seed)
    noise_signal = rs.normal(size=(n_pix, 1)).astype(np.float32)
    g, d, dg, sub_g = bbh.build_gan(noise_signal)
    og = ko.build(ko.bbh_generator_model(n_pix), seed=seed + 1)
    od = ko.build(ko.bbh_signal_discriminator_model(n_pix), seed=seed + 2)
    osub = ko.Sequential([ko.StackResidual(noise_signal.astype(np.float64))])
    ocomp = ko.Sequential([ko.Sequential([og, osub]), od])
    ocomp.build((100,))
    ko.set_trainable(od, False)
    ocomp.compile('binary_crossentropy', ko.Adam(9e-5, beta_1=0.5))
    ko.set_trainable(od, True)
    od.compile('binary_crossentropy', ko.Adam(9e-5, beta_1=0.5))
    sync_weights(g, og)
    sync_weights(d, od)
    z = rs.uniform(-1, 1, (B, 100)).astype(np.float32)
    sX = rs.normal(size=(2 * B, n_pix, 2, 1)).astype(np.float32)
    sy = np.array([1.0] * B + [0.0] * B, dtype=np.float32)
    return (g, d, dg), (og, od, ocomp), z, sX, sy


def burst_case(n_pix, B, seed=0):
    from gennet_b200 import nn, burst
    nn.clear_session()
    ko.clear_session()
    nn.set_seed(seed)
    burst.n_pix = n_pix
    rs = np.random.RandomState(seed)
    noise_signal = rs.normal(size=(n_pix, 1)).astype(np.float32) * 0.25
    g, d, dg, sub_g = burst.build_gan(noise_signal)
    og = ko.build(ko.burst_generator_model(n_pix), seed=seed + 1)
    od = ko.build(ko.burst_signal_discriminator_model(n_pix), seed=seed + 2)
    osub = ko.Sequential([og, ko.Sequential([ko.ResidualMoments(noise_signal.astype(np.float64))])])
    osub.build((100,))
    osub.compile('mean_squared_error', ko.Adam(2e-4, beta_1=0.5))
    ocomp = ko.Sequential([og, od])
    ocomp.build((100,))
    ko.set_trainable(od, False)
    ocomp.compile('binary_crossentropy', ko.Adam(2e-4, beta_1=0.5))
    ko.set_trainable(od, True)
    od.compile('binary_crossentropy', ko.Adam(2e-4, beta_1=0.5))
    sync_weights(g, og)
    sync_weights(d, od)
    z = rs.uniform(-1, 1, (B, 100)).astype(np.float32)
    sX = rs.normal(size=(2 * B, n_pix, 1)).astype(np.float32)
    sy = np.array([1.0] * B + [0.0] * B, dtype=np.float32)
    ny = np.zeros((B, 2), dtype=np.float32)
    ny[:, 1] = 0.25 ** 2
    return (g, d, dg, sub_g), (og, od, ocomp, osub), z, sX, sy, ny


def wvf_case(out_dim, B, seed=0):
    from gennet_b200 import nn, wvf
    nn.clear_session()
    ko.clear_session()
    nn.set_seed(seed)
    G_in = nn.Input(shape=[10])
    G, _ = wvf.get_generative(G_in, out_dim=out_dim)
    D_in = nn.Input(shape=[out_dim])
    D, _ = wvf.get_discriminative(D_in)
    GAN_in = nn.Input(shape=[10])
    GAN, _ = wvf.make_gan(GAN_in, G, D)
    og = ko.build(ko.wvf_get_generative(10, 300, out_dim), seed=seed + 1)
    og.compile('binary_crossentropy', ko.SGD(0.425e-1))
    od = ko.build(ko.wvf_get_discriminative(out_dim), seed=seed + 2)
    od.compile('binary_crossentropy', ko.Adam(1e-6))
    ogan = ko.Sequential([og, od])
    ogan.build((10,))
    ko.set_trainable(od, False)
    ogan.compile('binary_crossentropy', og.optimizer)
    sync_weights(G, og)
    sync_weights(D, od)
    rs = np.random.RandomState(seed)
    X = rs.uniform(0, 1, (2 * B, out_dim)).astype(np.float32)
    y = np.zeros((2 * B, 2), np.float32)
    y[:B, 1] = 1
    y[B:, 0] = 1
    z = rs.uniform(0, 1, (B, 10)).astype(np.float32)
    yz = np.zeros((B, 2), np.float32)
    yz[:, 1] = 1
    return (G, D, GAN), (og, od, ogan), X, y, z, yz


def two_model_case(B, seed=0, out_dim=50):
    """2_model_version (BASELINE config 5) wiring: transposed-conv generator, Conv1D discriminator, stacked GAN."""
    from gennet_b200 import nn, twomodel
    nn.clear_session()
    ko.clear_session()
    nn.set_seed(seed)
    G_in = nn.Input(shape=(1, 1))
    G, _ = twomodel.get_generative(G_in, out_dim=out_dim, lr=4e-3)
    D_in = nn.Input(shape=(out_dim,))
    D, _ = twomodel.get_discriminative(D_in, lr=4e-3)
    GAN_in = nn.Input((1, 1))
    GAN, _ = twomodel.make_gan(GAN_in, G, D)
    og = ko.build(ko.two_model_get_generative(1, out_dim), seed=seed + 1)
    og.compile('binary_crossentropy', ko.SGD(4e-3))
    od = ko.build(ko.two_model_get_discriminative(out_dim), seed=seed + 2)
    od.layers[-2].activation = None          # the script's Dense(n_channels) is linear (the shipped file has tanh)
    od.compile('binary_crossentropy', ko.Adam(4e-3, beta_1=0.5))
    ogan = ko.Sequential([og, od])
    ogan.build((1, 1))
    ko.set_trainable(od, False)
    ogan.compile('binary_crossentropy', og.optimizer)
    ko.set_trainable(od, True)
    sync_weights(G, og)
    sync_weights(D, od)
    rs = np.random.RandomState(seed)
    X = rs.normal(size=(2 * B, out_dim)).astype(np.float32)
    y = np.zeros((2 * B, 2), np.float32)
    y[:B, 1] = 1
    y[B:, 0] = 1
    z = rs.uniform(-5, 5, (B, 1, 1)).astype(np.float32)
    yz = np.zeros((B, 2), np.float32)
    yz[:, 1] = 1
    return (G, D, GAN), (og, od, ogan), X, y, z, yz


def subtract_case(B, seed=0, noise_dim=10, out_dim=50):
    """2_model_version/weight_version/subtract_model.py (the subtract stage of BASELINE config 5): ELU transposed-conv
    generator with the l1 activity / l2 kernel regularisers, Dropout discriminator, stacked GAN; inputs as the script
    samples them (noise rows, residual rows x_t - G(z); latents)."""
    from gennet_b200 import nn
    from gennet_b200.two_model import subtract_model as sm
    nn.clear_session()
    ko.clear_session()
    nn.set_seed(seed)
    sm.hyperparams.noise_dim, sm.hyperparams.outdim = noise_dim, out_dim
    G, _ = sm.get_generative(nn.Input(shape=(1, noise_dim)), out_dim=out_dim, lr=4e-3)
    D, _ = sm.get_discriminative(nn.Input(shape=(out_dim,)), lr=4e-3)
    GAN, _ = sm.make_gan(nn.Input((1, noise_dim)), G, D)
    og = ko.build(ko.subtract_get_generative(noise_dim, out_dim), seed=seed + 1)
    og.compile('binary_crossentropy', ko.Adam(4e-3, beta_1=0.5))
    od = ko.build(ko.subtract_get_discriminative(out_dim), seed=seed + 2)
    od.compile('binary_crossentropy', ko.Adam(4e-3, beta_1=0.5))
    ogan = ko.Sequential([og, od])
    ogan.build((1, noise_dim))
    ko.set_trainable(od, False)
    ogan.compile('binary_crossentropy', og.optimizer)
    ko.set_trainable(od, True)
    sync_weights(G, og)
    sync_weights(D, od)
    rs = np.random.RandomState(seed)
    X = rs.normal(0, 5, size=(2 * B, out_dim)).astype(np.float32)
    y = np.ones((2 * B, 2), np.float32)               # the script's labels: all ones (subtract_model.py:95-97)
    z = rs.normal(0, 1, size=(B, 1, noise_dim)).astype(np.float32)
    yz = np.ones((B, 2), np.float32)
    return (G, D, GAN), (og, od, ogan), X, y, z, yz


def nw_disc_case(B, seed=0, out_dim=50):
    """2_model_version/no_weight_code/subtract_model.py:322-390: the MSE discriminator with GaussianNoise and
    BatchNormalization(axis=1) blocks, global average pooling, Adam with decay, soft labels."""
    from gennet_b200 import nn
    from gennet_b200.two_model import subtract_model_nw as nw
    nn.clear_session()
    ko.clear_session()
    nn.set_seed(seed)
    D, _ = nw.get_discriminative(nn.Input(shape=(out_dim,)), lr=1e-3)
    od = ko.build(ko.nw_get_discriminative(out_dim), seed=seed + 2)
    od.compile('mean_squared_error', ko.Adam(1e-3, beta_1=0.5, decay=1e-4))
    sync_weights(D, od)
    rs = np.random.RandomState(seed)
    X = rs.normal(0, 5, size=(B, out_dim)).astype(np.float32)
    y = np.zeros((B, 2), np.float32)
    y[:B // 2, 0], y[B // 2:, 1] = rs.uniform(0.7, 1), rs.uniform(0.7, 1)
    y[:B // 2, 1], y[B // 2:, 0] = rs.uniform(0, 0.3), rs.uniform(0, 0.3)
    return D, od, X, y
