"""GPU parity of the bandwidth-bound companions with FLOAT32 activations (the float32 / split-operand modes): first
convolutions (Cin <= 2), last convolution (Cout = 1), Dense heads with <= 4 outputs and the BatchNormalization ->
activation -> dropout chain, each against torch float64 on the same inputs (float32-level tolerances)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import keras_oracle as ko
from tests.parity_cases import assert_close, case_seed

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


@pytest.mark.parametrize('case', [(3, 200, 1, 64, 5, 2, 'same'), (2, 256, 1, 64, 5, 1, 'same'), (2, 100, 2, 128, 5, 2, 'same'),
                                  (2, 77, 1, 32, 3, 1, 'valid'), (3, 128, 2, 512, 5, 2, 'same'),
                                  # any filter count / up to 16 taps: Conv1D(50, 16), Conv1D(25, 5) on one input channel
                                  (5, 300, 1, 50, 16, 1, 'valid'), (3, 513, 1, 25, 5, 1, 'valid'), (2, 90, 2, 70, 16, 2, 'same'),
                                  (2, 64, 1, 64, 16, 1, 'same'), (4, 40, 2, 3, 7, 1, 'valid')])
def test_first_layer_kernels_f32(case):
    from gennet_b200 import _lib as L_
    B, L, Cin, Cout, k, s, padding = case
    rs = np.random.RandomState(case_seed(case))
    x, w, bias = dev(rs.normal(size=(B, L, Cin))), dev(rs.normal(size=(k, Cin, Cout)) / math.sqrt(k * Cin)), dev(rs.normal(size=Cout))
    xr = x.cpu().double().requires_grad_(True)
    wr = w.cpu().double().requires_grad_(True)
    br = bias.cpu().double().requires_grad_(True)
    xp = xr.permute(0, 2, 1)
    pad = 0
    if padding == 'same':
        pl, pr = ko.same_pad(L, k, s)
        xp = F.pad(xp, (pl, pr))
        pad = pl
    yr = F.conv1d(xp, wr.permute(2, 1, 0), br, stride=s).permute(0, 2, 1)
    Lout = yr.shape[1]
    st = L_.stream()
    y = torch.full((B, Lout, Cout), float('nan'), device='cuda')
    L_.call('gn_conv1d_smallcin_fwd_f32', L_.ptr(x), L_.ptr(w), L_.ptr(bias), L_.ptr(y), B, L, Cin, Lout, Cout, k, s, pad,
            L_.ACT_RELU, 0.0, st)
    assert_close(y.cpu().numpy(), torch.relu(yr).detach().numpy(), 'smallcin fwd f32', 2e-6)
    dy = dev(rs.normal(size=(B, Lout, Cout)))
    (yr * dy.cpu().double()).sum().backward()
    dw = torch.full((k, Cin, Cout), float('nan'), device='cuda')
    db = torch.full((Cout,), float('nan'), device='cuda')
    L_.call('gn_conv1d_smallcin_wgrad_f32', L_.ptr(x), L_.ptr(dy), L_.ptr(dw), L_.ptr(db), B, L, Cin, Lout, Cout, k, s, pad, st)
    assert_close(dw.cpu().numpy(), wr.grad.numpy(), 'smallcin wgrad f32', 1e-5)
    assert_close(db.cpu().numpy(), br.grad.numpy(), 'smallcin bias grad f32', 1e-5)
    dx = torch.full((B, L, Cin), float('nan'), device='cuda')
    L_.call('gn_conv1d_smallcin_dgrad_f32', L_.ptr(dy), L_.ptr(w), L_.ptr(dx), B, L, Cin, Lout, Cout, k, s, pad, st)
    assert_close(dx.cpu().numpy(), xr.grad.numpy(), 'smallcin dgrad f32', 1e-5)


@pytest.mark.parametrize('M,K,N', [(5, 64000, 1), (3, 4096, 2), (9, 1000 * 8, 4)])
def test_dense_small_f32(M, K, N):
    from gennet_b200 import _lib as L_
    rs = np.random.RandomState(M * 7 + N)
    x = dev(np.maximum(rs.normal(size=(M, K)), 0))            # post-ReLU features
    w, b, dy = dev(rs.normal(size=(K, N)) / math.sqrt(K)), dev(rs.normal(size=N)), dev(rs.normal(size=(M, N)))
    xr, wr = x.cpu().double(), w.cpu().double()
    st = L_.stream()
    y = torch.empty(M, N, device='cuda')
    L_.call('gn_dense_small_fwd_f32', L_.ptr(x), L_.ptr(w), L_.ptr(b), L_.ptr(y), M, K, N, L_.ACT_NONE, 0.0, st)
    assert_close(y.cpu().numpy(), (xr @ wr + b.cpu().double()).numpy(), 'dense small fwd f32', 1e-5)
    dx = torch.empty(M, K, device='cuda')
    C = 64
    cs = torch.full((C,), float('nan'), device='cuda')
    L_.call('gn_dense_small_dgrad_f32', L_.ptr(dy), L_.ptr(w), L_.ptr(x), L_.ptr(dx), L_.ptr(cs), C, M, K, N, L_.ACT_RELU, 0.0, st)
    ref = (dy.cpu().double() @ wr.t()) * (xr > 0)
    assert_close(dx.cpu().numpy(), ref.numpy(), 'dense small dgrad*mask f32', 2e-6)
    assert_close(cs.cpu().numpy(), ref.reshape(M, K // C, C).sum((0, 1)).numpy(), 'dense small dgrad column sums f32', 1e-5)
    L_.call('gn_dense_small_dgrad_f32', L_.ptr(dy), L_.ptr(w), None, L_.ptr(dx), None, 0, M, K, N, L_.ACT_NONE, 0.0, st)
    assert_close(dx.cpu().numpy(), (dy.cpu().double() @ wr.t()).numpy(), 'dense small dgrad f32', 2e-6)
    dw = torch.full((K, N), float('nan'), device='cuda')
    db = torch.full((N,), float('nan'), device='cuda')
    L_.call('gn_dense_small_wgrad_f32', L_.ptr(x), L_.ptr(dy), L_.ptr(dw), L_.ptr(db), M, K, N, st)
    assert_close(dw.cpu().numpy(), (xr.t() @ dy.cpu().double()).numpy(), 'dense small wgrad f32', 1e-5)
    assert_close(db.cpu().numpy(), dy.cpu().double().sum(0).numpy(), 'dense small bias grad f32', 1e-6)


@pytest.mark.parametrize('case', [(3, 300, 1024, 5, 'same'), (2, 97, 64, 5, 'valid'), (2, 64, 256, 3, 'same'),
                                  (40, 37, 256, 5, 'same'), (1500, 9, 512, 4, 'valid')])
def test_cout1_conv_kernels_f32(case):
    from gennet_b200 import _lib as L_
    B, L, Cin, k, padding = case
    rs = np.random.RandomState(case_seed(case))
    x, w, bias = dev(rs.normal(size=(B, L, Cin))), dev(rs.normal(size=(k, Cin, 1)) / math.sqrt(k * Cin)), dev(rs.normal(size=1))
    xr = x.cpu().double().requires_grad_(True)
    wr = w.cpu().double().requires_grad_(True)
    br = bias.cpu().double().requires_grad_(True)
    xp = xr.permute(0, 2, 1)
    pad = 0
    if padding == 'same':
        pl, pr = ko.same_pad(L, k, 1)
        xp = F.pad(xp, (pl, pr))
        pad = pl
    yr = F.conv1d(xp, wr.permute(2, 1, 0), br).permute(0, 2, 1)
    Lout = yr.shape[1]
    st = L_.stream()
    y = torch.full((B, Lout, 1), float('nan'), device='cuda')
    L_.call('gn_conv1d_cout1_fwd_f32', L_.ptr(x), L_.ptr(w), L_.ptr(bias), L_.ptr(y), B, L, Cin, Lout, k, pad, st)
    assert_close(y.cpu().numpy(), yr.detach().numpy(), 'cout1 fwd f32', 2e-6)
    dy = dev(rs.normal(size=(B, Lout, 1)))
    (yr * dy.cpu().double()).sum().backward()
    dx = torch.full((B, L, Cin), float('nan'), device='cuda')
    L_.call('gn_conv1d_cout1_dgrad_f32', L_.ptr(dy), L_.ptr(w), L_.ptr(dx), B, L, Cin, Lout, k, pad, st)
    assert_close(dx.cpu().numpy(), xr.grad.numpy(), 'cout1 dgrad f32', 2e-6)
    dw = torch.full((k, Cin, 1), float('nan'), device='cuda')
    db = torch.full((1,), float('nan'), device='cuda')
    L_.call('gn_conv1d_cout1_wgrad_f32', L_.ptr(x), L_.ptr(dy), L_.ptr(dw), L_.ptr(db), B, L, Cin, Lout, k, pad, st)
    assert_close(dw.cpu().numpy(), wr.grad.numpy(), 'cout1 wgrad f32', 1e-5)
    assert_close(db.cpu().numpy(), br.grad.numpy(), 'cout1 bias grad f32', 1e-5)


@pytest.mark.parametrize('rows,C,act,noise', [(300, 64, 2, 0), (77, 912, 1, 1), (1000, 8, 4, -1), (64, 1024, 0, 0), (4096, 256, 2, 0),
                                                (515, 128, 4, 0)])
def test_f32_bn_act_dropout_chain(rows, C, act, noise):
    """gn_bn_sums_f32 / gn_chain_{fwd,bwd_sums,bwd}_f32 vs torch float64 autograd of drop(act(bn(x))) with a fed mask;
    the Philox-mask variant is self-consistent between forward and backward and equals gn_noise_draw_f32."""
    from gennet_b200 import _lib as L_
    rs = np.random.RandomState(rows + C)
    x, dy = dev(rs.normal(0.3, 1.5, size=(rows, C))), dev(rs.normal(size=(rows, C)))
    gamma, beta = dev(rs.uniform(0.5, 1.5, C)), dev(rs.normal(size=C))
    eps, rate = 1e-3, 0.2
    st = L_.stream()
    f64 = torch.float64
    sums = torch.empty(2 * C, dtype=f64, device='cuda')
    L_.call('gn_bn_sums_f32', L_.ptr(x), rows, C, L_.ptr(sums, f64), st)
    xd = x.double()
    assert torch.allclose(sums[:C], xd.sum(0), rtol=1e-12, atol=1e-9) and torch.allclose(sums[C:], (xd * xd).sum(0), rtol=1e-12)
    mean = xd.mean(0)
    invstd = 1.0 / torch.sqrt(xd.var(0, unbiased=False) + eps)
    if noise == 0:
        r = dev((rs.uniform(size=(rows, C)) >= rate))
        fac = r.double() / (1 - rate)
    elif noise == 1:
        r = dev(rs.normal(size=(rows, C)))
        fac = 1 + r.double() * np.sqrt(rate / (1 - rate))
    else:
        r, fac = None, torch.ones(rows, C, dtype=f64, device='cuda')
    xg = xd.clone().requires_grad_(True)
    gg = gamma.double().clone().requires_grad_(True)
    bg = beta.double().clone().requires_grad_(True)
    h = gg * (xg - xg.mean(0)) / torch.sqrt(xg.var(0, unbiased=False) + eps) + bg
    a = {0: h, 1: torch.relu(h), 2: torch.tanh(h), 4: torch.where(h >= 0, h, 0.2 * h)}[act]
    yref = a * fac
    (yref * dy.double()).sum().backward()
    y = torch.empty_like(x)
    meanf, invf = mean.float().contiguous(), invstd.float().contiguous()
    rp = L_.ptr(r) if r is not None else None
    L_.call('gn_chain_fwd_f32', L_.ptr(x), L_.ptr(y), L_.ptr(meanf), L_.ptr(invf), L_.ptr(gamma), L_.ptr(beta), 0, eps, act, 0.2,
            noise, rate, rp, 0, 0, rows, C, st)
    assert_close(y.cpu().numpy(), yref.detach().cpu().numpy(), 'chain fwd f32', 2e-6)
    L_.call('gn_chain_bwd_sums_f32', L_.ptr(x), L_.ptr(dy), L_.ptr(meanf), L_.ptr(invf), L_.ptr(gamma), L_.ptr(beta), act, 0.2, noise,
            rate, rp, 0, 0, rows, C, L_.ptr(sums, f64), st)
    dx, dgamma, dbeta = torch.empty_like(x), torch.empty(C, device='cuda'), torch.empty(C, device='cuda')
    L_.call('gn_chain_bwd_f32', L_.ptr(x), L_.ptr(dy), L_.ptr(dx), L_.ptr(meanf), L_.ptr(invf), L_.ptr(gamma), L_.ptr(beta),
            L_.ptr(sums, f64), float(rows), act, 0.2, noise, rate, rp, 0, 0, L_.ptr(dgamma), L_.ptr(dbeta), rows, C, st)
    # ReLU-type kinks: an element within rounding distance of zero may take the other side than float64
    kink_tol = 1e-5 if act in (1, 4) else 5e-6
    assert_close(dbeta.cpu().numpy(), bg.grad.cpu().numpy(), 'chain dbeta f32', kink_tol)
    assert_close(dgamma.cpu().numpy(), gg.grad.cpu().numpy(), 'chain dgamma f32', kink_tol)
    assert_close(dx.cpu().numpy(), xg.grad.cpu().numpy(), 'chain dx f32', 1e-5)
    # the forms with the max |result| side output (scale source of the f16x2 operand split): same results, exact maximum
    y_a, dx_a = torch.full_like(x, float('nan')), torch.full_like(x, float('nan'))
    am_y, am_dx = torch.full((1,), float('nan'), device='cuda'), torch.full((1,), float('nan'), device='cuda')
    L_.call('gn_chain_fwd_amax_f32', L_.ptr(x), L_.ptr(y_a), L_.ptr(meanf), L_.ptr(invf), L_.ptr(gamma), L_.ptr(beta), 0, eps, act,
            0.2, noise, rate, rp, 0, 0, rows, C, L_.ptr(am_y), st)
    L_.call('gn_chain_bwd_amax_f32', L_.ptr(x), L_.ptr(dy), L_.ptr(dx_a), L_.ptr(meanf), L_.ptr(invf), L_.ptr(gamma), L_.ptr(beta),
            L_.ptr(sums, f64), float(rows), act, 0.2, noise, rate, rp, 0, 0, None, None, rows, C, L_.ptr(am_dx), st)
    assert torch.equal(y_a, y) and torch.equal(dx_a, dx)
    assert am_y.item() == y.abs().max().item() and am_dx.item() == dx.abs().max().item()
    if act == 2 and noise in (-1, 0):
        # bounded activation: the apply pass also writes the next convolution's operand planes (scaled fp16 pair) with the
        # scale of the a-priori bound 1 / (1 - rate); they equal the split of y under that scale
        bound = 1.0 / (1.0 - rate) if noise == 0 else 1.0
        H = torch.float16
        y_p = torch.full_like(x, float('nan'))
        planes = torch.full((2, rows, C), float('nan'), dtype=H, device='cuda')
        am_p = torch.full((1,), float('nan'), device='cuda')
        L_.call('gn_chain_fwd_planes_f32', L_.ptr(x), L_.ptr(y_p), L_.ptr(meanf), L_.ptr(invf), L_.ptr(gamma), L_.ptr(beta), 0, eps,
                act, 0.2, noise, rate, rp, 0, 0, rows, C, L_.ptr(planes, H), L_.ptr(am_p), bound, st)
        assert torch.equal(y_p, y) and am_p.item() == np.float32(bound)
        ref = torch.empty((2, rows, C), dtype=H, device='cuda')
        L_.call('gn_split_f32_f16x2', L_.ptr(y), L_.ptr(ref, H), L_.ptr(am_p), 1, y.numel(), st)
        assert torch.equal(planes, ref)
        assert y.abs().max().item() <= bound
    if noise >= 0:
        rr = torch.empty(rows, C, device='cuda')
        L_.call('gn_noise_draw_f32', L_.ptr(rr), rows * C, noise, rate, 77, 1024, st)
        y1, y2 = torch.empty_like(x), torch.empty_like(x)
        L_.call('gn_chain_fwd_f32', L_.ptr(x), L_.ptr(y1), None, None, None, None, 0, 0.0, 0, 0.0, noise, rate, None, 77, 1024, rows, C, st)
        L_.call('gn_chain_fwd_f32', L_.ptr(x), L_.ptr(y2), None, None, None, None, 0, 0.0, 0, 0.0, noise, rate, L_.ptr(rr), 0, 0, rows, C, st)
        assert torch.equal(y1, y2)
        if act in (1, 4):
            # dropout behind a convolution with a fused ReLU / LeakyReLU: the mask-only backward pass also applies the
            # activation derivative, taken from the activation's OUTPUT (same sign as its input)
            post = torch.relu(x) if act == 1 else torch.where(x >= 0, x, 0.2 * x)
            L_.call('gn_chain_bwd_f32', L_.ptr(post.contiguous()), L_.ptr(dy), L_.ptr(y1), None, None, None, None, None, 1.0, act, 0.2,
                    noise, rate, L_.ptr(rr), 0, 0, None, None, rows, C, st)
            fr = (rr >= rate).double() / (1 - rate) if noise == 0 else 1 + rr.double() * np.sqrt(rate / (1 - rate))
            dref = dy.double() * fr * torch.where(x > 0, torch.ones_like(fr), torch.full_like(fr, 0.0 if act == 1 else 0.2))
            assert_close(y1.cpu().numpy(), dref.cpu().numpy(), 'dropout backward with the producer activation mask', 1e-6)
        L_.call('gn_chain_bwd_f32', L_.ptr(x), L_.ptr(dy), L_.ptr(y1), None, None, None, None, None, 1.0, 0, 0.0, noise, rate, None, 77,
                1024, None, None, rows, C, st)
        L_.call('gn_chain_bwd_f32', L_.ptr(x), L_.ptr(dy), L_.ptr(y2), None, None, None, None, None, 1.0, 0, 0.0, noise, rate, L_.ptr(rr),
                0, 0, None, None, rows, C, st)
        assert torch.equal(y1, y2)
