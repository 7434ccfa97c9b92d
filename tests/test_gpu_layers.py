"""GPU parity of the individual layer kernels (through the C ABI) against float64 torch-CPU math."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import keras_oracle as ko
from tests.parity_cases import assert_close, case_seed

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def lib():
    from gennet_b200 import _lib
    return _lib


def conv_ref(x, w, b, k, s, padding, up):
    xt = torch.as_tensor(x, dtype=torch.float64)
    if up > 1:
        xt = xt.repeat_interleave(up, dim=1)
    xt = xt.requires_grad_(True)
    wt = torch.as_tensor(w, dtype=torch.float64).requires_grad_(True)
    bt = torch.as_tensor(b, dtype=torch.float64).requires_grad_(True)
    xp = xt.permute(0, 2, 1)
    if padding == 'same':
        xp = F.pad(xp, ko.same_pad(xt.shape[1], k, s))
    y = F.conv1d(xp, wt.permute(2, 1, 0), bt, stride=s).permute(0, 2, 1)
    return xt, wt, bt, y


CONV_CASES = [
    # B, L(after up), Cin, Cout, k, s, padding, up
    (3, 64, 1, 64, 5, 2, 'same', 1),        # PE mc first layer
    (3, 64, 1, 64, 5, 1, 'same', 1),        # PE q first layer
    (2, 61, 16, 40, 5, 2, 'valid', 1),      # odd length, ragged channel counts
    (2, 40, 24, 136, 5, 1, 'valid', 1),
    (2, 32, 32, 16, 5, 2, 'same', 2),       # generator: UpSampling1D(2) -> Conv1D(s2)
    (2, 32, 32, 24, 5, 1, 'same', 2),       # generator: UpSampling1D(2) -> Conv1D(s1)
    (2, 48, 160, 1, 5, 1, 'same', 1),       # generator last layer (Cout=1 -> row-dot kernels)
    (3, 50, 1, 25, 5, 1, 'valid', 1),       # nn.py discriminator
    (2, 33, 4, 6, 3, 1, 'same', 1),
    (1, 300, 8, 130, 5, 1, 'same', 1),      # > one 128-row and > one 128-col tile
]


@pytest.mark.parametrize('case', CONV_CASES)
def test_conv1d_fwd_dgrad_wgrad(case):
    L_ = lib()
    B, L, Cin, Cout, k, s, padding, up = case
    rs = np.random.RandomState(case_seed(case))
    x = rs.normal(size=(B, L // up, Cin)).astype(np.float32)
    w = (rs.normal(size=(k, Cin, Cout)) / math.sqrt(k * Cin)).astype(np.float32)
    b = rs.normal(size=Cout).astype(np.float32)
    xt, wt, bt, y = conv_ref(x, w, b, k, s, padding, up)
    Lout = y.shape[1]
    pad = ko.same_pad(L, k, s)[0] if padding == 'same' else 0
    dy = rs.normal(size=(B, Lout, Cout)).astype(np.float32)
    (y * torch.as_tensor(dy, dtype=torch.float64)).sum().backward()
    st = L_.stream()
    dx_, dw_, x_, dy_ = cu(np.zeros_like(x)), cu(np.zeros_like(w)), cu(x), cu(dy)
    w_, b_ = cu(w), cu(b)
    y_ = torch.empty(B, Lout, Cout, device='cuda')
    db_ = torch.empty(Cout, device='cuda')
    L_.call('gn_conv1d_fwd_f32', L_.ptr(x_), L_.ptr(w_), L_.ptr(b_), L_.ptr(y_), B, L, Cin, Lout, Cout, k, s, pad, up,
            0, 0.0, st)
    assert_close(y_.cpu().numpy(), y.detach().numpy(), 'conv fwd', 2e-5)
    L_.call('gn_conv1d_dgrad_f32', L_.ptr(dy_), L_.ptr(w_), L_.ptr(dx_), B, L, Cin, Lout, Cout, k, s, pad, up, st)
    gx = xt.grad.numpy()
    if up > 1:
        gx = gx.reshape(B, L // up, up, Cin).sum(2)
    assert_close(dx_.cpu().numpy(), gx, 'conv dgrad', 2e-5)
    L_.call('gn_conv1d_wgrad_f32', L_.ptr(x_), L_.ptr(dy_), L_.ptr(dw_), L_.ptr(db_), B, L, Cin, Lout, Cout, k, s, pad, up, st)
    assert_close(dw_.cpu().numpy(), wt.grad.numpy(), 'conv wgrad', 2e-5)
    assert_close(db_.cpu().numpy(), bt.grad.numpy(), 'conv bias grad', 2e-5)
    # fused activation epilogue
    L_.call('gn_conv1d_fwd_f32', L_.ptr(x_), L_.ptr(w_), L_.ptr(b_), L_.ptr(y_), B, L, Cin, Lout, Cout, k, s, pad, up,
            L_.ACT_TANH, 0.0, st)
    assert_close(y_.cpu().numpy(), torch.tanh(y).detach().numpy(), 'conv fwd+tanh', 2e-5)


def test_conv1d_wgrad_split_k_large_reduction():
    """B*Lout large enough to take the split-K + atomics path."""
    L_ = lib()
    B, L, Cin, Cout, k, s = 16, 1024, 16, 32, 5, 1
    rs = np.random.RandomState(1)
    x = rs.normal(size=(B, L, Cin)).astype(np.float32)
    dy = rs.normal(size=(B, L - 4, Cout)).astype(np.float32)
    xt, wt, bt, y = conv_ref(x, np.zeros((k, Cin, Cout), np.float32), np.zeros(Cout, np.float32), k, s, 'valid', 1)
    (y * torch.as_tensor(dy, dtype=torch.float64)).sum().backward()
    dw_ = torch.empty(k, Cin, Cout, device='cuda')
    x_, dy_ = cu(x), cu(dy)
    L_.call('gn_conv1d_wgrad_f32', L_.ptr(x_), L_.ptr(dy_), L_.ptr(dw_), None, B, L, Cin, L - 4, Cout, k, s, 0, 1,
            L_.stream())
    assert_close(dw_.cpu().numpy(), wt.grad.numpy(), 'split-K wgrad', 2e-5)


@pytest.mark.parametrize('M,K,N', [(5, 100, 300), (7, 300, 1), (4, 513, 2), (9, 64, 1024), (130, 20, 140)])
def test_dense(M, K, N):
    L_ = lib()
    rs = np.random.RandomState(M * 1000 + N)
    x = rs.normal(size=(M, K)).astype(np.float32)
    w = (rs.normal(size=(K, N)) / math.sqrt(K)).astype(np.float32)
    b = rs.normal(size=N).astype(np.float32)
    dy = rs.normal(size=(M, N)).astype(np.float32)
    y_ = torch.empty(M, N, device='cuda')
    st = L_.stream()
    x_, w_, b_, dy_ = cu(x), cu(w), cu(b), cu(dy)
    L_.call('gn_dense_fwd_f32', L_.ptr(x_), L_.ptr(w_), L_.ptr(b_), L_.ptr(y_), M, K, N, L_.ACT_SIGMOID, 0.0, st)
    ref = 1 / (1 + np.exp(-(x.astype(np.float64) @ w.astype(np.float64) + b)))
    assert_close(y_.cpu().numpy(), ref, 'dense fwd', 2e-5)
    dx_ = torch.empty(M, K, device='cuda')
    L_.call('gn_dense_dgrad_f32', L_.ptr(dy_), L_.ptr(w_), L_.ptr(dx_), M, K, N, st)
    assert_close(dx_.cpu().numpy(), dy.astype(np.float64) @ w.astype(np.float64).T, 'dense dgrad', 2e-5)
    dw_, db_ = torch.empty(K, N, device='cuda'), torch.empty(N, device='cuda')
    L_.call('gn_dense_wgrad_f32', L_.ptr(x_), L_.ptr(dy_), L_.ptr(dw_), L_.ptr(db_), M, K, N, st)
    assert_close(dw_.cpu().numpy(), x.astype(np.float64).T @ dy.astype(np.float64), 'dense wgrad', 2e-5)
    assert_close(db_.cpu().numpy(), dy.astype(np.float64).sum(0), 'dense bias grad', 2e-5)


def test_conv2d_width2_equals_conv2d_same():
    """Conv2D(5x5, strides (2,1), same) on (L,2,C) executed as Conv1D must equal a true 2-D convolution."""
    from gennet_b200 import nn
    nn.clear_session()
    rs = np.random.RandomState(3)
    for cin, cout, L in [(1, 8, 32), (6, 10, 30)]:
        layer = nn.Conv2D(cout, (5, 5), strides=(2, 1), padding='same', input_shape=(L, 2, cin))
        m = nn.Sequential([layer])
        w = rs.normal(size=(5, 5, cin, cout)).astype(np.float32) * 0.2
        b = rs.normal(size=cout).astype(np.float32)
        layer.set_weights([w, b])
        x = rs.normal(size=(3, L, 2, cin)).astype(np.float32)
        o = ko.Conv2D(cout, (5, 5), strides=(2, 1), padding='same')
        o.build((L, 2, cin), torch.Generator().manual_seed(0), torch.float64)
        o.weights[0].data = torch.as_tensor(w, dtype=torch.float64)
        o.weights[1].data = torch.as_tensor(b, dtype=torch.float64)
        ref = o.forward(torch.as_tensor(x, dtype=torch.float64), False, {}).detach().numpy()
        assert_close(m.predict(x), ref, 'conv2d width-2', 2e-5)


def test_batchnorm_fwd_bwd_and_moving_stats():
    from gennet_b200 import nn
    nn.clear_session()
    rs = np.random.RandomState(4)
    for shape in [(6, 37), (4, 19, 24)]:
        bn = nn.BatchNormalization(momentum=0.9, input_shape=shape[1:])
        m = nn.Sequential([bn])
        m.compile(loss='mean_squared_error', optimizer=nn.SGD(lr=0.0))
        g = rs.uniform(0.5, 1.5, shape[-1]).astype(np.float32)
        be = rs.normal(size=shape[-1]).astype(np.float32)
        bn.set_weights([g, be, np.zeros(shape[-1], np.float32), np.ones(shape[-1], np.float32)])
        x = (rs.normal(size=shape) * 3 + 100).astype(np.float32)     # large mean: two-pass variance matters
        o = ko.BatchNormalization(0.9)
        om = ko.Sequential([o])
        om.build(shape[1:])
        o.weights[0].data = torch.as_tensor(g, dtype=torch.float64)
        o.weights[1].data = torch.as_tensor(be, dtype=torch.float64)
        om.compile('mean_squared_error', ko.SGD(0.0))
        y = rs.normal(size=shape).astype(np.float32)
        om.train_on_batch(x, y)
        m.train_on_batch(x, y)
        gp = m.get_gradients()
        assert_close(gp[0], om.last_grads[0], 'dgamma', 1e-4)
        assert_close(gp[1], om.last_grads[1], 'dbeta', 1e-4)
        pw = bn.get_weights()
        assert_close(pw[2], o.state[0].numpy(), 'moving_mean', 1e-5)
        assert_close(pw[3], o.state[1].numpy(), 'moving_var', 1e-4)
        assert_close(m.predict(x), om.predict(x), 'bn inference', 1e-4)


def test_elementwise_layers_and_pooling():
    L_ = lib()
    st = L_.stream()
    rs = np.random.RandomState(5)
    x = rs.normal(size=(3, 10, 7)).astype(np.float32)
    dy = rs.normal(size=(3, 5, 7)).astype(np.float32)
    y_ = torch.empty(3, 5, 7, device='cuda')
    x_, dy_ = cu(x), cu(dy)
    L_.call('gn_maxpool1d_fwd_f32', L_.ptr(x_), L_.ptr(y_), 3, 10, 7, 2, st)
    xt = torch.as_tensor(x, dtype=torch.float64).requires_grad_(True)
    ref = F.max_pool1d(xt.permute(0, 2, 1), 2).permute(0, 2, 1)
    assert np.array_equal(y_.cpu().numpy(), ref.detach().numpy().astype(np.float32))
    (ref * torch.as_tensor(dy, dtype=torch.float64)).sum().backward()
    dx_ = torch.empty(3, 10, 7, device='cuda')
    L_.call('gn_maxpool1d_bwd_f32', L_.ptr(x_), L_.ptr(y_), L_.ptr(dy_), L_.ptr(dx_), 3, 10, 7, 2, st)
    assert_close(dx_.cpu().numpy(), xt.grad.numpy(), 'maxpool bwd', 1e-7)
    up_ = torch.empty(3, 20, 7, device='cuda')
    L_.call('gn_upsample1d_fwd_f32', L_.ptr(x_), L_.ptr(up_), 3, 10, 7, 2, st)
    assert np.array_equal(up_.cpu().numpy(), np.repeat(x, 2, axis=1))
    for act, p, f in [(L_.ACT_RELU, 0.0, lambda t: torch.relu(t)), (L_.ACT_TANH, 0.0, torch.tanh),
                      (L_.ACT_SIGMOID, 0.0, torch.sigmoid), (L_.ACT_LEAKY, 0.2, lambda t: torch.where(t >= 0, t, 0.2 * t)),
                      (L_.ACT_RELU_MAX, 1.0, lambda t: torch.clamp(t, 0, 1.0))]:
        xt = torch.as_tensor(x, dtype=torch.float64).requires_grad_(True)
        r = f(xt)
        g = torch.as_tensor(rs.normal(size=x.shape))
        (r * g).sum().backward()
        a_ = torch.empty_like(x_)
        L_.call('gn_act_fwd_f32', L_.ptr(x_), L_.ptr(a_), x.size, act, p, st)
        assert_close(a_.cpu().numpy(), r.detach().numpy(), 'act fwd %d' % act, 1e-6)
        d_ = torch.empty_like(a_)
        g_ = cu(g.numpy())
        L_.call('gn_act_bwd_f32', L_.ptr(g_), L_.ptr(a_), L_.ptr(d_), x.size, act, p, st)
        assert_close(d_.cpu().numpy(), xt.grad.numpy(), 'act bwd %d' % act, 1e-5)


def test_losses_and_metrics():
    L_ = lib()
    rs = np.random.RandomState(6)
    B = 37
    for kind, D, name in [(L_.LOSS_BCE, 1, 'bce'), (L_.LOSS_BCE, 2, 'bce2'), (L_.LOSS_MSE, 1, 'mse'), (L_.LOSS_MSE, 2, 'mse2'),
                          (L_.LOSS_CHISQ, 1, 'chisq')]:
        p = rs.uniform(0, 1, (B, D)).astype(np.float32)
        # clipped probabilities.  (An exact 1.0 is left out of the strict check: in float32, as in the
        # reference's TF float32 graph, 1-1e-7 rounds to 1-1.19e-7, which moves that one loss term by 1 %.)
        p[0, 0], p[1, 0] = 0.0, 0.9999
        t = (rs.uniform(size=(B, D)) > 0.5).astype(np.float32)
        pt = torch.as_tensor(p, dtype=torch.float64).requires_grad_(True)
        tt = torch.as_tensor(t, dtype=torch.float64)
        if kind == L_.LOSS_BCE:
            l = ko.binary_crossentropy(tt, pt)
        elif kind == L_.LOSS_MSE:
            l = ko.mean_squared_error(tt, pt)
        else:
            l = ko.chisquare_loss(0.7)(tt, pt)
        l.mean().backward()
        out = torch.zeros(2, device='cuda')
        dp = torch.empty(B, D, device='cuda')
        mk = 0 if (D == 1 or kind == L_.LOSS_BCE) else 1
        p_, t_ = cu(p), cu(t)
        L_.call('gn_loss_fwd_bwd_f32', L_.ptr(p_), L_.ptr(t_), L_.ptr(out), L_.ptr(dp), B, D, kind, 0.7, 1.0 / B, 0, mk,
                L_.stream())
        assert_close(out[0].item() / B, l.mean().item(), name + ' loss', 1e-5)
        assert_close(dp.cpu().numpy(), pt.grad.numpy(), name + ' dpred', 1e-5)
        acc = ko.accuracy(tt, pt.detach(), 'binary_crossentropy' if kind == L_.LOSS_BCE else 'mean_squared_error')
        assert abs(out[1].item() / B - acc.item()) < 1e-6


def test_adam_and_sgd_kernels():
    L_ = lib()
    rs = np.random.RandomState(8)
    n = 1000
    p0, g = rs.normal(size=n).astype(np.float32), rs.normal(size=n).astype(np.float32)
    p, m, v, g_ = cu(p0), cu(np.zeros(n)), cu(np.zeros(n)), cu(g)
    op = torch.tensor(p0, dtype=torch.float64)
    opt = ko.Adam(9e-5, beta_1=0.5)
    for t in range(3):
        opt.step([op], [torch.tensor(g, dtype=torch.float64)])
        lr_t = 9e-5 * math.sqrt(1 - 0.999 ** (t + 1)) / (1 - 0.5 ** (t + 1))
        L_.call('gn_adam_step_f32', L_.ptr(p), L_.ptr(g_), L_.ptr(m), L_.ptr(v), n, lr_t, 0.5, 0.999, 1e-7, 1.0, L_.stream())
    assert_close(p.cpu().numpy() - p0, op.numpy() - p0, 'adam 3 steps', 2e-3)
    q = cu(p0)
    L_.call('gn_sgd_step_f32', L_.ptr(q), L_.ptr(g_), n, 0.0425, 0.5, L_.stream())
    assert_close(q.cpu().numpy(), p0 - 0.0425 * 0.5 * g, 'sgd', 1e-6)


def test_philox_streams():
    L_ = lib()
    n = 1 << 20
    u = torch.empty(n, device='cuda')
    L_.call('gn_uniform_f32', L_.ptr(u), n, -1.0, 1.0, 42, 0, L_.stream())
    assert abs(u.mean().item()) < 5e-3 and abs(u.var().item() - 1.0 / 3.0) < 5e-3
    assert u.min().item() > -1.0 and u.max().item() < 1.0
    z = torch.empty(n, device='cuda')
    L_.call('gn_normal_f32', L_.ptr(z), n, 0.0, 1.0, 42, 0, L_.stream())
    assert abs(z.mean().item()) < 5e-3 and abs(z.std().item() - 1.0) < 5e-3
    # counter-based: a later window of the same stream equals the tail of the full draw
    z2 = torch.empty(n // 2, device='cuda')
    L_.call('gn_normal_f32', L_.ptr(z2), n // 2, 0.0, 1.0, 42, n // 2, L_.stream())
    assert torch.equal(z2, z[n // 2:])
    k = torch.empty(n, device='cuda')
    L_.call('gn_noise_draw_f32', L_.ptr(k), n, L_.NOISE_DROPOUT, 0.4, 7, 0, L_.stream())
    assert abs(k.mean().item() - 0.6) < 5e-3 and set(k.unique().tolist()) == {0.0, 1.0}
    # host re-implementation of Philox4x32-10 pins the generator
    def philox(c, key):
        M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
        c = list(c); key = list(key)
        for _ in range(10):
            p0, p1 = M0 * c[0], M1 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ key[0], p1 & 0xffffffff, (p0 >> 32) ^ c[3] ^ key[1], p0 & 0xffffffff]
            key = [(key[0] + W0) & 0xffffffff, (key[1] + W1) & 0xffffffff]
        return c
    r = philox([5, 0, 0, 1], [42, 0])
    exp = [((x >> 8) + 0.5) / 16777216.0 * 2.0 - 1.0 for x in r]
    assert np.allclose(u[20:24].cpu().numpy(), np.array(exp, dtype=np.float32), atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize('rows,C,act,noise', [(300, 64, 2, 0), (77, 912, 1, 1), (1000, 8, 4, -1), (64, 1024, 0, 0)])
def test_bf16_bn_act_dropout_chain(rows, C, act, noise):
    """gn_bn_stats_bf16 / gn_chain_{fwd,bwd_sums,bwd}_bf16 vs torch float64 autograd of drop(act(bn(x))) on the same
    bf16-rounded x, dy and a fed mask; and the Philox-mask variant is self-consistent between forward and backward."""
    from gennet_b200 import _lib as L_
    rs = np.random.RandomState(rows + C)
    bfl = torch.bfloat16
    x = torch.as_tensor(rs.normal(0.3, 1.5, size=(rows, C)).astype(np.float32)).cuda().to(bfl)
    dy = torch.as_tensor(rs.normal(size=(rows, C)).astype(np.float32)).cuda().to(bfl)
    gamma = torch.as_tensor(rs.uniform(0.5, 1.5, C).astype(np.float32)).cuda()
    beta = torch.as_tensor(rs.normal(size=C).astype(np.float32)).cuda()
    eps, rate = 1e-3, 0.2
    st = L_.stream()
    sums = torch.empty(2 * C, dtype=torch.float64, device='cuda')
    L_.call('gn_bn_stats_bf16', L_.ptr(x, bfl), rows, C, L_.ptr(sums, torch.float64), st)
    xd = x.double()
    assert torch.allclose(sums[:C], xd.sum(0), rtol=1e-6, atol=1e-6) and torch.allclose(sums[C:], (xd * xd).sum(0), rtol=1e-6)
    mean = xd.mean(0)
    var = xd.var(0, unbiased=False)
    invstd = 1.0 / torch.sqrt(var + eps)
    if noise == 0:
        r = torch.as_tensor((rs.uniform(size=(rows, C)) >= rate).astype(np.float32)).cuda()
        fac = r.double() / (1 - rate)
    elif noise == 1:
        r = torch.as_tensor(rs.normal(size=(rows, C)).astype(np.float32)).cuda()
        fac = 1 + r.double() * np.sqrt(rate / (1 - rate))
    else:
        r, fac = None, torch.ones(rows, C, dtype=torch.float64, device='cuda')
    xg = xd.clone().requires_grad_(True)
    gg = gamma.double().clone().requires_grad_(True)
    bg = beta.double().clone().requires_grad_(True)
    mu_b = xg.mean(0)
    h = gg * (xg - mu_b) / torch.sqrt(xg.var(0, unbiased=False) + eps) + bg
    a = {0: h, 1: torch.relu(h), 2: torch.tanh(h), 4: torch.where(h >= 0, h, 0.2 * h)}[act]
    yref = a * fac
    (yref * dy.double()).sum().backward()
    y = torch.empty_like(x)
    meanf, invf = mean.float().contiguous(), invstd.float().contiguous()
    rp = L_.ptr(r) if r is not None else None
    L_.call('gn_chain_fwd_bf16', L_.ptr(x, bfl), L_.ptr(y, bfl), L_.ptr(meanf), L_.ptr(invf), L_.ptr(gamma), L_.ptr(beta), 0, eps,
            act, 0.2, noise, rate, rp, 0, 0, rows, C, st)
    assert_close(y.float().cpu().numpy(), yref.detach().cpu().numpy(), 'chain fwd', 2 ** -7)
    L_.call('gn_chain_bwd_sums_bf16', L_.ptr(x, bfl), L_.ptr(dy, bfl), L_.ptr(meanf), L_.ptr(invf), L_.ptr(gamma), L_.ptr(beta), act,
            0.2, noise, rate, rp, 0, 0, rows, C, L_.ptr(sums, torch.float64), st)
    dx = torch.empty_like(x)
    dgamma = torch.empty(C, device='cuda')
    dbeta = torch.empty(C, device='cuda')
    L_.call('gn_chain_bwd_bf16', L_.ptr(x, bfl), L_.ptr(dy, bfl), L_.ptr(dx, bfl), L_.ptr(meanf), L_.ptr(invf), L_.ptr(gamma),
            L_.ptr(beta), L_.ptr(sums, torch.float64), float(rows), act, 0.2, noise, rate, rp, 0, 0, L_.ptr(dgamma), L_.ptr(dbeta),
            rows, C, st)
    assert_close(dbeta.cpu().numpy(), bg.grad.cpu().numpy(), 'chain dbeta', 1e-4)
    assert_close(dgamma.cpu().numpy(), gg.grad.cpu().numpy(), 'chain dgamma', 1e-4)
    assert_close(dx.float().cpu().numpy(), xg.grad.cpu().numpy(), 'chain dx', 2 ** -7)
    if noise >= 0:
        # Philox masks: identical to gn_noise_draw_f32 at the same (seed, offset), in forward and in backward
        rr = torch.empty(rows, C, device='cuda')
        L_.call('gn_noise_draw_f32', L_.ptr(rr), rows * C, noise, rate, 77, 1024, st)
        y1, y2 = torch.empty_like(x), torch.empty_like(x)
        L_.call('gn_chain_fwd_bf16', L_.ptr(x, bfl), L_.ptr(y1, bfl), None, None, None, None, 0, 0.0, 0, 0.0, noise, rate, None, 77,
                1024, rows, C, st)
        L_.call('gn_chain_fwd_bf16', L_.ptr(x, bfl), L_.ptr(y2, bfl), None, None, None, None, 0, 0.0, 0, 0.0, noise, rate, L_.ptr(rr),
                0, 0, rows, C, st)
        assert torch.equal(y1, y2)
        L_.call('gn_chain_bwd_bf16', L_.ptr(x, bfl), L_.ptr(dy, bfl), L_.ptr(y1, bfl), None, None, None, None, None, 1.0, 0, 0.0,
                noise, rate, None, 77, 1024, None, None, rows, C, st)
        L_.call('gn_chain_bwd_bf16', L_.ptr(x, bfl), L_.ptr(dy, bfl), L_.ptr(y2, bfl), None, None, None, None, None, 1.0, 0, 0.0,
                noise, rate, L_.ptr(rr), 0, 0, None, None, rows, C, st)
        assert torch.equal(y1, y2)
