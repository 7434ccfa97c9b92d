"""Self-consistency checks of the Keras-semantics oracle (no TF available: parity unpinned,
so the oracle is checked against closed forms and finite differences)."""
import math

import numpy as np
import torch

from oracle import keras_oracle as ko


def test_same_padding_rule():
    assert ko.same_pad(1024, 5, 2) == (1, 2)      # [A2] k5,s2,even L
    assert ko.same_pad(1024, 5, 1) == (2, 2)
    assert ko.same_pad(2, 5, 1) == (2, 2)         # D's width-2 axis


def test_conv1d_same_stride2_against_direct_sum():
    torch.manual_seed(0)
    c = ko.Conv1D(3, 5, strides=2, padding='same')
    c.build((10, 2), torch.Generator().manual_seed(1), torch.float64)
    x = torch.randn(1, 10, 2, dtype=torch.float64)
    y = c.forward(x, False, {})
    W, b = c.weights
    xp = torch.zeros(1, 13, 2, dtype=torch.float64)
    xp[:, 1:11] = x
    ref = torch.stack([sum(xp[0, 2 * l + k] @ W[k] for k in range(5)) + b for l in range(5)])
    assert y.shape == (1, 5, 3)
    assert torch.allclose(y[0], ref, atol=1e-12)


def test_bn_train_and_moving_update():
    bn = ko.BatchNormalization(momentum=0.9)
    bn.build((4, 3), None, torch.float64)
    x = torch.randn(5, 4, 3, dtype=torch.float64)
    y = bn.forward(x, True, {})
    flat = x.reshape(-1, 3)
    assert torch.allclose(y.reshape(-1, 3).mean(0), torch.zeros(3, dtype=torch.float64), atol=1e-12)
    n = 20.0
    var_b = flat.var(0, unbiased=False)
    # Keras 2.2.4 on TF 1.12: zero-debiased moving average -- after the first update the moving statistics ARE the
    # batch statistics (the initial 0 / 1 are discarded), after the second the debiased mix of the two batches
    assert torch.allclose(bn.state[0], flat.mean(0))
    assert torch.allclose(bn.state[1], var_b * n / (n - 1.001))
    x2 = torch.randn(5, 4, 3, dtype=torch.float64) + 1.0
    bn.forward(x2, True, {})
    f2 = x2.reshape(-1, 3)
    want = (0.9 * 0.1 * flat.mean(0) + 0.1 * f2.mean(0)) / (1 - 0.9 ** 2)
    assert torch.allclose(bn.state[0], want)
    y2 = bn.forward(x, False, {})
    assert torch.allclose(y2, (x - bn.state[0]) / torch.sqrt(bn.state[1] + 1e-3))
    # plain exponential average (tf.keras / zero_debias=False)
    try:
        ko.ZERO_DEBIAS = False
        bn = ko.BatchNormalization(momentum=0.9)
        bn.build((4, 3), None, torch.float64)
        bn.forward(x, True, {})
        assert torch.allclose(bn.state[0], 0.1 * flat.mean(0))
        assert torch.allclose(bn.state[1], 0.9 + 0.1 * var_b * n / (n - 1.001))
    finally:
        ko.ZERO_DEBIAS = True


def test_adam_first_step_closed_form():
    p = torch.tensor([1.0, -2.0], dtype=torch.float64, requires_grad=True)
    g = torch.tensor([0.5, -0.25], dtype=torch.float64)
    opt = ko.Adam(lr=1e-3, beta_1=0.5)
    opt.step([p], [g])
    lr_t = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.5)
    m, v = 0.5 * g, 0.001 * g * g
    assert torch.allclose(p.detach(), torch.tensor([1.0, -2.0], dtype=torch.float64) - lr_t * m / (v.sqrt() + 1e-7))


def test_bce_matches_clipped_log_form():
    yp = torch.tensor([[0.2], [0.9], [1.0], [0.0]], dtype=torch.float64)
    yt = torch.tensor([[1.0], [0.0], [1.0], [1.0]], dtype=torch.float64)
    l = ko.binary_crossentropy(yt, yp)
    pc = yp.clamp(1e-7, 1 - 1e-7)
    ref = -(yt * pc.log() + (1 - yt) * (1 - pc).log()).squeeze(-1)
    assert torch.allclose(l, ref, atol=1e-12)


def test_builders_param_counts_match_survey():
    def count(m):
        return sum(w.numel() for l in m.all_layers() for w in l.weights + l.state)
    g = ko.build(ko.bbh_generator_model(1024))
    assert g.out_shape == (1024, 1)
    assert abs(count(g) - 17.08e6) < 0.3e6          # 17.08 M (incl. BN moving stats)
    d = ko.build(ko.bbh_signal_discriminator_model(1024))
    assert abs(count(d) - 3.55e6) < 0.02e6
    pe = ko.build(ko.bbh_signal_pe_model(1024))
    assert abs(count(pe) - 4.63e6) < 0.02e6
    bg = ko.build(ko.burst_generator_model(512))
    assert abs(count(bg) - 7.46e6) < 0.02e6
    bd = ko.build(ko.burst_signal_discriminator_model(512))
    assert abs(count(bd) - 16.56e6) < 0.02e6
    wd = ko.build(ko.wvf_get_discriminative(8192))
    assert abs(count(wd) - 5.12e6) < 0.02e6


def test_finite_difference_gradient_small_pe():
    m = ko.build(ko.bbh_signal_pe_model(128), seed=3)
    m.compile('mean_squared_error', ko.SGD(lr=0.0))
    rs = np.random.RandomState(0)
    x = rs.normal(size=(3, 128, 1))
    y = [rs.uniform(20, 35, 3) / 35.0, rs.uniform(0.5, 1, 3)]
    # bias the heads positive so ReLU outputs are active
    m.branches[0][-2].weights[1].data += 0.5
    m.branches[1][-2].weights[1].data += 0.5

    def loss():
        outs = m.forward(torch.as_tensor(x), True, {})
        return float(sum(((o - torch.as_tensor(t)[:, None]) ** 2).mean() for o, t in zip(outs, y)))
    m.train_on_batch(x, y)
    W = m.branches[1][2].weights[0]          # second conv of q tower
    gidx = [i for i, w in enumerate(m.collected) if w is W][0]
    g = m.last_grads[gidx]
    for idx in [(0, 0, 0), (2, 5, 7), (4, 63, 127)]:
        old = W.data[idx].item()
        W.data[idx] = old + 1e-6
        lp = loss()
        W.data[idx] = old - 1e-6
        lm = loss()
        W.data[idx] = old
        assert abs((lp - lm) / 2e-6 - g[idx]) < 1e-6 * max(1.0, abs(g[idx]))


def test_gan_composite_only_updates_generator():
    n = 64
    g = ko.build(ko.bbh_generator_model(n), seed=1)
    d = ko.build(ko.bbh_signal_discriminator_model(n), seed=2)
    sub = ko.Sequential([ko.StackResidual(np.zeros((n, 1)))])
    comp = ko.Sequential([ko.Sequential([g, sub]), d])
    comp.build((100,))
    ko.set_trainable(d, False)
    comp.compile('binary_crossentropy', ko.Adam(9e-5, beta_1=0.5))
    ko.set_trainable(d, True)
    d.compile('binary_crossentropy', ko.Adam(9e-5, beta_1=0.5))
    dw0 = [w.copy() for w in d.get_weights()]
    gw0 = [w.copy() for w in g.get_weights()]
    z = np.random.RandomState(0).uniform(-1, 1, (4, 100))
    out = comp.train_on_batch(z, [1] * 4)
    assert len(out) == 2 and np.isfinite(out[0])
    assert all(np.array_equal(a, b) for a, b in zip(dw0, d.get_weights()))
    assert any(not np.array_equal(a, b) for a, b in zip(gw0, g.get_weights()))
    sx = np.random.RandomState(1).normal(size=(8, n, 2, 1))
    d.train_on_batch(sx, [1.0] * 4 + [0.0] * 4)
    assert any(not np.array_equal(a, b) for a, b in zip(dw0, d.get_weights()))
