"""GPU parity of whole training steps: product (sm_100a kernels via the C ABI) vs the float64 oracle on
identical inputs, weights and dropout draws -- forward outputs, losses, one-step gradients (rtol 1e-4),
updated weights, BatchNorm moving statistics."""
import numpy as np
import pytest
import torch

from tests import parity_cases as pc

pytestmark = pytest.mark.gpu


def test_pe_step_parity():
    prod, orc, x, y = pc.pe_case(256, 8)
    errs, w0 = pc.compare_step(prod, orc, x, y)
    pc.compare_weights(prod, orc, w0)
    pc.resync([(prod, orc)])
    errs, w0 = pc.compare_step(prod, orc, x, y, check_predict=False)      # t=2: bias correction, persistent moments
    pc.compare_weights(prod, orc, w0)


def test_gan_steps_parity():
    (g, d, dg), (og, od, ocomp), z, sX, sy = pc.gan_case(128, 8)
    pc.assert_close(g.predict(z), og.predict(z), 'generator.predict')
    errs, w0 = pc.compare_step(d, od, sX, sy)
    pc.compare_weights(d, od, w0)
    pc.resync([(d, od)])
    dw = [w.copy() for w in d.get_weights()]
    errs, w0 = pc.compare_step(dg, ocomp, z, [1] * 8, check_predict=False)
    assert all(np.array_equal(a, b) for a, b in zip(dw, d.get_weights()))
    pc.compare_weights(g, og, w0[:len(g.get_weights())])
    # moving statistics were updated by the training-mode pass and drive the next predict
    pc.resync([(g, og)])
    pc.assert_close(g.predict(z), og.predict(z), 'generator.predict after step')


def test_burst_iteration_parity():
    (g, d, dg, sub_g), (og, od, ocomp, osub), z, sX, sy, ny = pc.burst_case(64, 8)
    errs, w0 = pc.compare_step(d, od, sX, sy)
    pc.compare_weights(d, od, w0)
    pc.resync([(d, od)])
    errs, w0 = pc.compare_step(sub_g, osub, z, ny, check_predict=False)
    pc.compare_weights(g, og, w0[:len(g.get_weights())])
    pc.resync([(g, og)])
    errs, w0 = pc.compare_step(dg, ocomp, z, [1] * 8, check_predict=False)
    pc.compare_weights(g, og, w0[:len(g.get_weights())])


def test_wvf_models_parity():
    (G, D, GAN), (og, od, ogan), X, y, z, yz = pc.wvf_case(512, 8)
    pc.assert_close(G.predict(z), og.predict(z), 'G.predict')
    errs, w0 = pc.compare_step(D, od, X, y)
    pc.compare_weights(D, od, w0)
    pc.resync([(D, od)])
    errs, w0 = pc.compare_step(GAN, ogan, z, yz, check_predict=False)
    pc.compare_weights(G, og, w0[:len(G.get_weights())])


def test_device_resident_steps_and_on_device_dropout():
    """bbh.pe_train_step / gan_train_step run from device-resident inputs with Philox dropout."""
    from gennet_b200 import nn, bbh
    nn.clear_session()
    nn.set_seed(3)
    bbh.n_pix = 128
    pe = bbh.signal_pe_model()
    pe.compile(loss='mean_squared_error', optimizer=nn.Adam(lr=9e-5, beta_1=0.5), metrics=['accuracy'])
    g = torch.Generator(device='cuda').manual_seed(0)
    templates = torch.randn(50, 128, device='cuda', generator=g)
    pars = torch.rand(50, 2, device='cuda', generator=g)
    idx = torch.randint(0, 50, (16,), device='cuda', generator=g, dtype=torch.int32)
    noise = torch.randn(2, 128, device='cuda', generator=g)
    l0 = bbh.pe_train_step(pe, templates, pars, idx, noise, 2.5)
    assert len(l0) == 5 and all(np.isfinite(l0))
    for _ in range(20):
        l1 = bbh.pe_train_step(pe, templates, pars, idx, noise, 2.5)
    assert l1[0] < l0[0]                        # the step actually trains
    noise_signal = np.random.RandomState(0).normal(size=(128, 1)).astype(np.float32)
    G, D, DG, _ = bbh.build_gan(noise_signal)
    ns = torch.as_tensor(noise_signal.reshape(-1)).cuda()
    real = torch.randn(8, 128, device='cuda', generator=g)
    z1 = torch.rand(8, 100, device='cuda', generator=g) * 2 - 1
    z2 = torch.rand(8, 100, device='cuda', generator=g) * 2 - 1
    rn = torch.randn(8, 128, device='cuda', generator=g)
    sd, sg = bbh.gan_train_step(G, D, DG, ns, real, z1, rn, z2)
    assert len(sd) == 2 and len(sg) == 2 and np.isfinite(sd + sg).all()


def test_pe_step_bf16_within_stated_tolerance():
    """Throughput mode (bf16 operands on the tcgen05 tensor cores, fp32 accumulation / weights / optimizer).
    Stated tolerance vs the float64 oracle: forward outputs 2e-2 of the output scale, losses 3e-2, every one-step
    gradient 1e-1 relative L2 and 3e-1 of its max.  (bf16 carries 8 mantissa bits; activations AND back-propagated
    gradients are rounded at each of the five stacked layers, and weight gradients are cancellation-heavy sums, so
    the first layer sees ~7e-2.  The kernels themselves are pinned to 2^-8 / 1e-4 on bf16-rounded operands in
    tests/test_gpu_conv_tc.py; this test pins the wiring of the bf16 graph.)"""
    from gennet_b200 import nn
    try:
        nn.set_compute_dtype('bfloat16')
        prod, orc, x, y = pc.pe_case(256, 8)
        po, oo = prod.predict(x), orc.predict(x)
        for a, b in zip(po, oo):
            assert np.abs(a - b).max() <= 2e-2 * max(np.abs(b).max(), 1e-6)
        ro = orc.train_on_batch(x, y)
        rp = prod.train_on_batch(x, y)
        assert np.allclose(rp[:3], ro[:3], rtol=3e-2, atol=1e-4)
        gp = prod.get_gradients()
        gmax = max(np.abs(g).max() for g in orc.last_grads)
        for i, (a, b) in enumerate(zip(gp, orc.last_grads)):
            el2 = np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-3 * gmax * np.sqrt(b.size))
            emax = np.abs(a - b).max() / max(np.abs(b).max(), 1e-3 * gmax)
            assert np.isfinite(a).all() and el2 <= 1e-1 and emax <= 3e-1, (i, b.shape, el2, emax)
        # the tensor-core kernels were really used
        convs = [l for l in prod.all_layers() if isinstance(l, nn.Conv1D)]
        assert sorted(c._path() for c in convs) == ['smallcin'] * 2 + ['tc'] * 7
    finally:
        nn.set_compute_dtype('float32')


def test_gan_steps_bf16_within_stated_tolerance():
    """Throughput mode on the GAN of bbhMahoGANy.py (BASELINE config 3): generator convolutions (UpSampling folded,
    Cout = 1 tail) and the discriminator's packed Conv2D layers on the tensor-core / streaming bf16 kernels.
    Stated tolerance vs the float64 oracle: outputs 2e-2 of scale, losses 3e-2 (+1e-4), gradients 1.5e-1 relative L2 /
    3e-1 of max (the generator's Dense(100 -> 128 n_pix) also runs with bf16 operands on the tensor cores).  Dropout
    masks are fed; BatchNorm runs in float32 between the bf16 convolutions."""
    from gennet_b200 import nn
    try:
        nn.set_compute_dtype('bfloat16')
        (g, d, dg), (og, od, ocomp), z, sX, sy = pc.gan_case(256, 8)
        a, b = g.predict(z), og.predict(z)
        assert np.abs(a - b).max() <= 2e-2 * np.abs(b).max()
        a, b = d.predict(sX), od.predict(sX)
        assert np.abs(a - b).max() <= 2e-2 * max(np.abs(b).max(), 1e-6)

        def step(prod, orc, x, y, seed):
            noise = pc.draw_noise(orc, x, seed)
            rp = prod.train_on_batch(x, y, _noise=pc.map_noise(noise, orc, prod))
            ro = orc.train_on_batch(x, y, noise=dict(noise))
            assert np.allclose(rp[0], ro[0], rtol=3e-2, atol=1e-4), (rp, ro)
            gp = prod.get_gradients()
            gmax = max(np.abs(q).max() for q in orc.last_grads)
            assert len(gp) == len(orc.last_grads)
            for i, (p_, q) in enumerate(zip(gp, orc.last_grads)):
                # biases in front of BatchNorm have structurally zero gradients: compare those against 1e-2 of the
                # largest gradient of the step instead of their own rounding noise
                el2 = np.linalg.norm((p_ - q).ravel()) / max(np.linalg.norm(q.ravel()), 1e-2 * gmax * np.sqrt(q.size))
                emax = np.abs(p_ - q).max() / max(np.abs(q).max(), 1e-2 * gmax)
                assert np.isfinite(p_).all() and el2 <= 1.5e-1 and emax <= 3e-1, (i, q.shape, el2, emax)
        step(d, od, sX, sy, 0)
        pc.resync([(d, od)])
        step(dg, ocomp, z, [1] * 8, 1)
        convs = [l for l in g.all_layers() if isinstance(l, nn.Conv1D)]
        assert [c._path() for c in convs] == ['tc', 'tc', 'tc', 'tc', 'tc', 'cout1']
        assert [l._path() for l in d.all_layers() if isinstance(l, nn.Conv2D)] == ['smallcin', 'tc']
    finally:
        nn.set_compute_dtype('float32')


@pytest.mark.gpu
def test_keras_hdf5_roundtrip_on_device(tmp_path):
    """save -> load_model on the GPU path: identical predictions, optimizer state carried over (bbhMahoGANy.py:1173,1135)."""
    from gennet_b200 import nn
    prod, orc, x, y = pc.pe_case(256, 4)
    prod.train_on_batch(x, y)
    p = str(tmp_path / 'signal_pe.h5')
    prod.save(p, True)
    m2 = nn.load_model(p)
    for a, b in zip(prod.predict(x), m2.predict(x)):
        assert np.array_equal(a, b)
    r1, r2 = prod.train_on_batch(x, y), m2.train_on_batch(x, y)
    assert np.allclose(r1, r2, rtol=1e-5, atol=1e-7)
    for a, b in zip(prod.get_weights(), m2.get_weights()):
        assert np.allclose(a, b, rtol=1e-5, atol=1e-7)
    prod.save_weights(str(tmp_path / 'w.h5'), True)
    m2.load_weights(str(tmp_path / 'w.h5'))
    for a, b in zip(prod.get_weights(), m2.get_weights()):
        assert np.array_equal(a, b)


def test_two_model_version_parity():
    """BASELINE config 5 (2_model_version): Conv2DTranspose generator (widths 1->4->11->26->57), Conv1D discriminator:
    predict, D step, G step through the frozen D -- outputs, losses, every gradient, updates vs the float64 oracle."""
    (G, D, GAN), (og, od, ogan), X, y, z, yz = pc.two_model_case(16)
    pc.assert_close(G.predict(z), og.predict(z), 'G.predict')
    errs, w0 = pc.compare_step(D, od, X, y)
    pc.compare_weights(D, od, w0)
    pc.resync([(D, od)])
    dw = [w.copy() for w in D.get_weights()]
    errs, w0 = pc.compare_step(GAN, ogan, z, yz, check_predict=False)
    assert all(np.array_equal(a, b) for a, b in zip(dw, D.get_weights()))
    pc.compare_weights(G, og, w0[:len(G.get_weights())])


def test_posterior_sampling_chain_on_device():
    """bbhMahoGANy.py:1311-1343: generator.predict -> signal_pe.predict, chained on the device, equals the two
    host-level predict calls; the percentile curves equal the reference's per-sample np.percentile loop."""
    from gennet_b200 import nn, bbh
    nn.clear_session()
    nn.set_seed(5)
    bbh.n_pix = 128
    g = bbh.generator_model()
    pe = bbh.signal_pe_model()
    z = np.random.RandomState(0).uniform(-1, 1, (24, 100)).astype(np.float32)
    samples, gen = bbh.posterior_samples(g, pe, n=24, batch=10, z=z)
    gen_ref = g.predict(z)
    ref = pe.predict(gen_ref.reshape(24, 128, 1))
    assert np.array_equal(gen, gen_ref)
    for a, b in zip(samples, ref):
        assert a.shape == (24, 1) and np.array_equal(a, b)
    s2, _ = bbh.posterior_samples(g, pe, n=50, seed=3, batch=16)          # Philox latents: batch-size invariant
    s3, _ = bbh.posterior_samples(g, pe, n=50, seed=3, batch=50)
    assert np.allclose(s2[0], s3[0], rtol=0, atol=1e-6) and np.allclose(s2[1], s3[1], rtol=0, atol=1e-6)
    pc_ = bbh.waveform_percentiles(gen)
    for p in (90, 75, 25, 5):
        assert np.allclose(pc_[p], [np.percentile(gen[:, n, 0], p) for n in range(128)], rtol=1e-5, atol=1e-7)


# ---- split-operand tensor-core modes ('bf16x3': three bf16 planes; 'f16x2': scaled fp16 pairs): the SAME rtol 1e-4
# as the float32 SIMT path
@pytest.fixture(params=['bf16x3', 'f16x2'])
def bf16x3(request):
    from gennet_b200 import nn
    nn.set_compute_dtype(request.param)
    try:
        yield nn
    finally:
        nn.set_compute_dtype('float32')


@pytest.mark.parametrize('n_pix,B', [(256, 8), (2048, 8)])
def test_pe_step_parity_bf16x3(bf16x3, n_pix, B):
    """CNN point estimator on the tcgen05 tensor cores with three-plane split operands: predict, losses, every
    one-step gradient within rtol 1e-4 of the float64 oracle, updated weights -- at the BASELINE n_pix too."""
    nn = bf16x3
    prod, orc, x, y = pc.pe_case(n_pix, B)
    convs = [l for l in prod.all_layers() if isinstance(l, nn.Conv1D)]
    assert sorted(c._path() for c in convs) == ['smallcin32'] * 2 + ['tc3'] * 7
    errs, w0 = pc.compare_step(prod, orc, x, y)
    pc.compare_weights(prod, orc, w0)
    print('%s PE n_pix %d: max gradient error %.2e' % (nn.compute_dtype(), n_pix, max(v for k, v in errs.items() if k.startswith('grad'))))
    if n_pix == 256:
        pc.resync([(prod, orc)])
        errs, w0 = pc.compare_step(prod, orc, x, y, check_predict=False)
        pc.compare_weights(prod, orc, w0)


@pytest.mark.parametrize('n_pix,B', [(128, 8), (2048, 8)])
def test_gan_steps_parity_bf16x3(bf16x3, n_pix, B):
    nn = bf16x3
    (g, d, dg), (og, od, ocomp), z, sX, sy = pc.gan_case(n_pix, B)
    convs = [l for l in g.all_layers() if isinstance(l, nn.Conv1D)]
    assert [c._path() for c in convs] == ['tc3', 'tc3', 'tc3', 'tc3', 'tc3', 'cout1_32']
    assert [l._path() for l in d.all_layers() if isinstance(l, nn.Conv2D)] == ['smallcin32', 'tc3']
    pc.assert_close(g.predict(z), og.predict(z), 'generator.predict')
    errs, w0 = pc.compare_step(d, od, sX, sy)
    pc.compare_weights(d, od, w0)
    pc.resync([(d, od)])
    dw = [w.copy() for w in d.get_weights()]
    errs, w0 = pc.compare_step(dg, ocomp, z, [1] * B, check_predict=False)
    assert all(np.array_equal(a, b) for a, b in zip(dw, d.get_weights()))
    pc.compare_weights(g, og, w0[:len(g.get_weights())])
    print('%s GAN n_pix %d: max gradient error %.2e' % (nn.compute_dtype(), n_pix, max(v for k, v in errs.items() if k.startswith('grad'))))


def test_burst_iteration_parity_bf16x3(bf16x3):
    (g, d, dg, sub_g), (og, od, ocomp, osub), z, sX, sy, ny = pc.burst_case(512, 8)
    errs, w0 = pc.compare_step(d, od, sX, sy)
    pc.compare_weights(d, od, w0)
    pc.resync([(d, od)])
    errs, w0 = pc.compare_step(sub_g, osub, z, ny, check_predict=False)
    pc.compare_weights(g, og, w0[:len(g.get_weights())])
    pc.resync([(g, og)])
    errs, w0 = pc.compare_step(dg, ocomp, z, [1] * 8, check_predict=False)
    pc.compare_weights(g, og, w0[:len(g.get_weights())])


def test_posterior_samples_and_percentiles_vs_oracle():
    """bbhMahoGANy.py:1311-1343 / :913-921 against the float64 oracle carrying the same weights: generator.predict ->
    signal_pe.predict chained on the device equals og.predict -> ope.predict (rtol 1e-4), and the device percentile
    curves equal np.percentile of the oracle's generated waveforms."""
    from gennet_b200 import nn, bbh
    from oracle import keras_oracle as ko
    nn.clear_session()
    ko.clear_session()
    nn.set_seed(5)
    bbh.n_pix = 256
    g, pe = bbh.generator_model(), bbh.signal_pe_model()
    og = ko.build(ko.bbh_generator_model(256), seed=11)
    ope = ko.build(ko.bbh_signal_pe_model(256), seed=12)
    ope.branches[0][-2].weights[1].data += 0.5          # keep the ReLU heads active
    ope.branches[1][-2].weights[1].data += 0.5
    pc.sync_weights(g, og)
    pc.sync_weights(pe, ope)
    z = np.random.RandomState(3).uniform(-1, 1, (37, 100)).astype(np.float32)
    samples, gen = bbh.posterior_samples(g, pe, n=37, batch=16, z=z)
    gen_ref = og.predict(z)
    pc.assert_close(gen, gen_ref, 'generated waveforms')
    ref = ope.predict(gen_ref)
    for k, (a, b) in enumerate(zip(samples, ref)):
        pc.assert_close(a, b, 'posterior sample %d' % k)
    curves = bbh.waveform_percentiles(gen, percentiles=(90, 75, 50, 25, 5, 0, 100))
    gr = gen_ref.reshape(37, 256)
    for p, c in curves.items():
        want = np.array([np.percentile(gr[:, n], p) for n in range(256)])      # the reference's loop (:918-921)
        pc.assert_close(c, want, 'percentile %g' % p)


@pytest.mark.parametrize('n,L', [(1000, 64), (4000, 130), (37, 9), (2, 5), (4097, 16)])
def test_percentile_kernel_matches_numpy(n, L):
    from gennet_b200 import bbh
    rs = np.random.RandomState(n + L)
    x = rs.normal(size=(n, L)).astype(np.float32)
    x[:, 0] = 1.0                                           # ties
    got = bbh.waveform_percentiles(x, percentiles=(90, 75, 25, 5, 33.3))
    for p, c in got.items():
        want = np.percentile(x.astype(np.float64), p, axis=0)
        assert np.abs(c - want).max() <= 1e-6 * max(np.abs(want).max(), 1.0), (p, np.abs(c - want).max())


def test_pe_train_step_batch_assembly_vs_oracle():
    """bbhMahoGANy.py:1153-1166 on the device (bbh.pe_train_step: gn_gather_rows_f32 + gn_add_scaled_f32 + train step)
    against the reference's host assembly `x = templates[idx]; x[:B/8] += sigma * n; y = pars[idx]` fed to the float64
    oracle: the assembled batch and labels are identical, the losses and gradients agree within rtol 1e-4."""
    from gennet_b200 import bbh
    prod, orc, _, _ = pc.pe_case(256, 16)
    rs = np.random.RandomState(7)
    templates = rs.normal(size=(50, 256)).astype(np.float32)
    pars = rs.uniform(0.3, 1.0, (50, 2)).astype(np.float32)
    idx = rs.randint(0, 50, 16).astype(np.int32)
    noise = rs.normal(size=(2, 256)).astype(np.float32)
    sigma = 2.5
    want_x = templates[idx].copy()
    want_x[:2] += np.float32(sigma) * noise
    want_y = pars[idx]
    seen = {}
    orig = prod.train_on_batch

    def spy(x, y, **kw):
        seen['x'] = x.detach().cpu().numpy().copy()
        seen['y'] = [t.detach().cpu().numpy().copy() for t in y]
        return orig(x, y, **kw)
    prod.train_on_batch = spy
    dev = lambda a, dt=torch.float32: torch.as_tensor(a).cuda().to(dt)
    rec = pc.record_kinks(prod)
    rp = bbh.pe_train_step(prod, dev(templates), dev(pars), dev(idx, torch.int32), dev(noise), sigma)
    assert np.abs(seen['x'].reshape(16, 256) - want_x).max() <= 2.4e-7 * np.abs(want_x).max()      # one fma vs mul + add
    assert np.array_equal(seen['y'][0], want_y[:, 0]) and np.array_equal(seen['y'][1], want_y[:, 1])
    ro = orc.train_on_batch(seen['x'].reshape(16, 256, 1), [want_y[:, 0], want_y[:, 1]],
                            noise={'__kinks__': pc.map_kinks(rec, orc, prod)})
    pc.assert_close(rp, ro, 'pe_train_step losses', 2e-4)
    for i, (a, b) in enumerate(zip(prod.get_gradients(), orc.last_grads)):
        pc.assert_close(a, b, 'pe_train_step gradient %d' % i, 1e-4, floor=1e-3 * max(np.abs(q).max() for q in orc.last_grads))


@pytest.mark.parametrize('which', ['pe', 'gan'])
def test_step_parity_bf16x2(which):
    """Two-plane split operands (three plane products per K step, 16 mantissa bits per operand): the whole-step parity
    suite at the same rtol 1e-4, at the BASELINE n_pix.  Reported by bench.py as an extra line, not the headline."""
    from gennet_b200 import nn
    nn.set_compute_dtype('bf16x2')
    try:
        if which == 'pe':
            prod, orc, x, y = pc.pe_case(2048, 8)
            errs, w0 = pc.compare_step(prod, orc, x, y)
            pc.compare_weights(prod, orc, w0)
        else:
            (g, d, dg), (og, od, ocomp), z, sX, sy = pc.gan_case(2048, 8)
            pc.assert_close(g.predict(z), og.predict(z), 'generator.predict')
            errs, w0 = pc.compare_step(d, od, sX, sy)
            pc.resync([(d, od)])
            errs2, w0 = pc.compare_step(dg, ocomp, z, [1] * 8, check_predict=False)
            errs.update({'g_' + k: v for k, v in errs2.items()})
        print('bf16x2 %s n_pix 2048: max gradient error %.2e' % (which, max(v for k, v in errs.items() if 'grad' in k)))
    finally:
        nn.set_compute_dtype('float32')


@pytest.mark.parametrize('mode', ['float32', 'bf16x3', 'f16x2'])
def test_subtract_stage_parity(mode):
    """2_model_version/weight_version/subtract_model.py (config 5 subtract stage): ELU transposed-conv generator with the
    l1 activity / l2 kernel regularisers, Dropout discriminator: predict, D step, G step through the frozen D."""
    from gennet_b200 import nn
    nn.set_compute_dtype(mode)
    try:
        (G, D, GAN), (og, od, ogan), X, y, z, yz = pc.subtract_case(16)
        pc.assert_close(G.predict(z), og.predict(z), 'G.predict')
        errs, w0 = pc.compare_step(D, od, X, y)
        pc.compare_weights(D, od, w0)
        pc.resync([(D, od)])
        dw = [w.copy() for w in D.get_weights()]
        errs, w0 = pc.compare_step(GAN, ogan, z, yz, check_predict=False)
        assert all(np.array_equal(a, b) for a, b in zip(dw, D.get_weights()))
        pc.compare_weights(G, og, w0[:len(G.get_weights())])
    finally:
        nn.set_compute_dtype('float32')


@pytest.mark.parametrize('mode', ['float32', 'bf16x3', 'f16x2'])
def test_nw_discriminator_parity(mode):
    """2_model_version/no_weight_code/subtract_model.py:322-390: Conv1D(tanh) -> LeakyReLU -> GaussianNoise(1.6) ->
    BatchNormalization(axis=1) blocks, GlobalAveragePooling1D, MSE, Adam with decay."""
    from gennet_b200 import nn
    nn.set_compute_dtype(mode)
    try:
        D, od, X, y = pc.nw_disc_case(16)
        if mode != 'float32':
            assert [l._path() for l in D.all_layers() if isinstance(l, nn.Conv1D)] == ['smallcin32', 'tc3', 'tc3']
        pc.assert_close(D.predict(X), od.predict(X), 'D.predict')
        errs, w0 = pc.compare_step(D, od, X, y)
        pc.compare_weights(D, od, w0)
    finally:
        nn.set_compute_dtype('float32')


@pytest.mark.parametrize('mode', ['f16x2', 'bf16x3'])
def test_gan_iterations_track_the_float32_path(mode):
    """Six consecutive device-resident GAN iterations (Philox dropout, Adam updates, moving statistics, cached operand
    planes of the weights rebuilt every step) in a split-operand tensor-core mode against the float32 SIMT path from the
    same seed: the loss trajectories stay together (nothing stale is carried from one iteration to the next)."""
    from gennet_b200 import nn, bbh

    def run(m):
        nn.clear_session()
        nn.set_seed(5)
        nn.set_compute_dtype(m)
        bbh.n_pix = 256
        noise_signal = np.random.RandomState(0).normal(size=(256, 1)).astype(np.float32)
        G, D, DG, _ = bbh.build_gan(noise_signal)
        ns = torch.as_tensor(noise_signal.reshape(-1)).cuda()
        g = torch.Generator(device='cuda').manual_seed(1)
        out = []
        w0 = [w.copy() for w in G.get_weights()]
        for it in range(6):
            real = torch.randn(16, 256, device='cuda', generator=g)
            z1 = torch.rand(16, 100, device='cuda', generator=g) * 2 - 1
            z2 = torch.rand(16, 100, device='cuda', generator=g) * 2 - 1
            rn = torch.randn(16, 256, device='cuda', generator=g)
            sd, sg = bbh.gan_train_step(G, D, DG, ns, real, z1, rn, z2)
            out.append([sd[0], sg[0]])
        return np.array(out), [w.copy() for w in G.get_weights()], w0
    try:
        ref, wref, w0 = run('float32')
        got, wgot, w0b = run(mode)
    finally:
        nn.set_compute_dtype('float32')
    assert all(np.array_equal(a, b) for a, b in zip(w0, w0b))
    assert np.isfinite(got).all()
    # the order of the float atomics differs from run to run: observed 3e-5 .. 2e-4 on losses of ~0.7
    assert np.allclose(got, ref, rtol=2e-3, atol=2e-4), np.abs(got - ref).max()
    # the generator's weights after six Adam steps, relative to how far they moved (Adam normalises the update, so a
    # gradient component near zero turns a 1e-6 difference into a step of the order of the learning rate)
    num = sum(float(((a - b) ** 2).sum()) for a, b in zip(wgot, wref))
    den = sum(float(((b - c) ** 2).sum()) for b, c in zip(wref, w0))
    print('%s vs float32 after 6 GAN iterations: max loss difference %.2e, weight difference / movement %.2e' % (
        mode, np.abs(got - ref).max(), (num / den) ** 0.5))
    assert (num / den) ** 0.5 < 5e-2, (num / den) ** 0.5
