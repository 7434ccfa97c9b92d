"""GPU parity of the split-operand tcgen05 kernels (conv1d_tc3.cu) with SCALED FP16 PAIR operands ("f16x2") against
float64 math on the same float32 inputs.  A tensor is carried as T0 = fp16(t s), T1 = fp16((t s - T0) 2^11) with the power
of two s taken from the tensor's max |t|: 22-23 bits relative to the tensor's scale, three plane products per K step (half
of bf16x3).  Tolerance: 8e-6 of the result's scale at test size (largest observed 4.4e-6: the fp32 tensor-memory accumulator
truncates, ~0.3 ulp per accumulation, up to 320 accumulations of the leading products at K = 5120), 1e-5 at BASELINE size."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import keras_oracle as ko
from tests.parity_cases import assert_close, case_seed
from tests.test_gpu_conv_tc3 import CASES, dev

pytestmark = pytest.mark.gpu
H = torch.float16
TOL = 8e-6


def split_h(x, have=None):
    from gennet_b200 import _lib as L_
    p = torch.empty((2,) + tuple(x.shape), dtype=H, device='cuda')
    amax = have if have is not None else torch.full((1,), float('nan'), device='cuda')
    L_.call('gn_split_f32_f16x2', L_.ptr(x), L_.ptr(p, H), L_.ptr(amax), 1 if have is not None else 0, x.numel(), L_.stream())
    return p, amax


def scale_of(amax):
    a = float(amax.item())
    return 2.0 ** (14 - math.floor(math.log2(a))) if a > 0 else 1.0


def unsplit(p, amax):
    return (p[0].double() + p[1].double() / 2048.0) / scale_of(amax)


@pytest.mark.parametrize('spread', [0.0, 6.0, 20.0])
def test_f16x2_planes_reconstruct(spread):
    """amax, the scale rule and the reconstruction error: <= 2^-22 of an element down to 2^-28 of the tensor's max,
    <= 2^-22 * 2^-28 of the tensor's max below that."""
    rs = np.random.RandomState(1)
    x = dev(rs.normal(size=(8192,)) * np.exp(rs.uniform(-spread, spread, 8192)) * 3e-7)
    p, amax = split_h(x)
    assert amax.item() == x.abs().max().item()
    assert 2 ** 14 <= p[0].abs().max().item() <= 2 ** 15
    err = (unsplit(p, amax) - x.double()).abs()
    bound = torch.maximum(x.double().abs(), torch.full_like(err, x.abs().max().item() * 2.0 ** -28)) * 2.0 ** -22
    assert (err <= bound).all(), (err / bound).max().item()
    # have_amax: the split takes the scalar it is given
    p2, _ = split_h(x, have=amax.clone())
    assert torch.equal(p, p2)
    a2 = torch.full((1,), float('nan'), device='cuda')
    from gennet_b200 import _lib as L_
    L_.call('gn_amax_f32', L_.ptr(x), x.numel(), L_.ptr(a2), L_.stream())
    assert a2.item() == amax.item()
    z, az = split_h(torch.zeros(64, device='cuda'))
    assert az.item() == 0.0 and (z == 0).all()


@pytest.mark.parametrize('rows,C', [(1000, 64), (333, 1024), (77, 8), (4096, 2048), (5, 512)])
def test_f16x2_gradient_split_with_column_sums(rows, C):
    """gn_split_colsum_f32_f16x2: the planes equal the plain split, colsum = per-channel sums of dy (the bias gradient)."""
    from gennet_b200 import _lib as L_
    rs = np.random.RandomState(rows + C)
    dy = dev(rs.normal(0.1, 1.0, size=(rows, C)) * 1e-3)
    ref, amax = split_h(dy)
    p = torch.full((2, rows, C), float('nan'), dtype=H, device='cuda')
    a2 = torch.full((1,), float('nan'), device='cuda')
    cs = torch.full((C,), float('nan'), device='cuda')
    L_.call('gn_split_colsum_f32_f16x2', L_.ptr(dy), L_.ptr(p, H), L_.ptr(a2), 0, rows, C, L_.ptr(cs), L_.stream())
    assert torch.equal(p, ref) and a2.item() == amax.item()
    assert_close(cs.cpu().numpy(), dy.double().sum(0).cpu().numpy(), 'column sums of the gradient split', 5e-6)
    L_.call('gn_split_colsum_f32_f16x2', L_.ptr(dy), L_.ptr(p, H), L_.ptr(amax), 1, rows, C, L_.ptr(cs), L_.stream())
    assert torch.equal(p, ref)


@pytest.mark.parametrize('scale_x,scale_w', [(1.0, 1.0), (3e-9, 5e3)])
@pytest.mark.parametrize('case', CASES)
def test_f16x2_conv_fwd_dgrad_wgrad(case, scale_x, scale_w):
    from gennet_b200 import _lib as L_
    B, L, Cin, Cout, k, s, padding = case
    if scale_x != 1.0 and B > 3:
        pytest.skip('rescaled operands: small cases only')
    rs = np.random.RandomState(case_seed(case))
    x = dev(rs.normal(size=(B, L, Cin)) * scale_x)
    w = dev(rs.normal(size=(k, Cin, Cout)) / math.sqrt(k * Cin) * scale_w)
    bias = dev(rs.normal(size=Cout) * scale_x * scale_w)
    st = L_.stream()
    wk = torch.empty(2, k, Cin, Cout, dtype=H, device='cuda')
    wt = torch.empty(2, k, Cout, Cin, dtype=H, device='cuda')
    w_amax = torch.full((1,), float('nan'), device='cuda')
    L_.call('gn_conv_w_split_f16x2', L_.ptr(w), L_.ptr(wk, H), L_.ptr(wt, H), L_.ptr(w_amax), k, Cin, Cout, st)
    assert torch.equal(wt, wk.permute(0, 1, 3, 2).contiguous())
    assert torch.equal(wk, split_h(w)[0]) and w_amax.item() == w.abs().max().item()
    big = B * L * Cin > 4e6
    rdev = 'cuda' if big else 'cpu'
    xr = x.to(rdev).double().requires_grad_(True)
    wr = w.to(rdev).double().requires_grad_(True)
    br = bias.to(rdev).double()
    xp = xr.permute(0, 2, 1)
    pad = 0
    if padding == 'same':
        pl, pr = ko.same_pad(L, k, s)
        xp = F.pad(xp, (pl, pr))
        pad = pl
    yr = F.conv1d(xp, wr.permute(2, 1, 0), br, stride=s).permute(0, 2, 1)
    Lout = yr.shape[1]
    xs, x_amax = split_h(x)
    y = torch.full((B, Lout, Cout), float('nan'), device='cuda')
    y_amax = torch.full((1,), float('nan'), device='cuda')
    L_.call('gn_conv1d_fwd_f16x2', L_.ptr(xs, H), L_.ptr(x_amax), L_.ptr(wt, H), L_.ptr(w_amax), L_.ptr(bias), L_.ptr(y),
            L_.ptr(y_amax), B, L, Cin, Lout, Cout, k, s, pad, L_.ACT_NONE, 0.0, st)
    torch.cuda.synchronize()
    assert_close(y.cpu().numpy(), yr.detach().cpu().numpy(), 'f16x2 conv fwd', TOL)
    assert y_amax.item() == y.abs().max().item()
    L_.call('gn_conv1d_fwd_f16x2', L_.ptr(xs, H), L_.ptr(x_amax), L_.ptr(wt, H), L_.ptr(w_amax), L_.ptr(bias), L_.ptr(y),
            None, B, L, Cin, Lout, Cout, k, s, pad, L_.ACT_RELU, 0.0, st)
    assert_close(y.cpu().numpy(), torch.relu(yr).detach().cpu().numpy(), 'f16x2 conv fwd+relu', TOL)
    # the form that also gathers the BatchNormalization statistics of its output in the epilogue
    if Cout <= 1024:
        y2 = torch.full((B, Lout, Cout), float('nan'), device='cuda')
        sums = torch.full((2 * Cout,), float('nan'), dtype=torch.float64, device='cuda')
        L_.call('gn_conv1d_fwd_stats_f16x2', L_.ptr(xs, H), L_.ptr(x_amax), L_.ptr(wt, H), L_.ptr(w_amax), L_.ptr(bias), L_.ptr(y2),
                L_.ptr(sums, torch.float64), B, L, Cin, Lout, Cout, k, s, pad, L_.ACT_RELU, 0.0, st)
        torch.cuda.synchronize()
        assert torch.equal(y2, y)
        yd = y.double().reshape(-1, Cout)
        assert torch.allclose(sums[:Cout], yd.sum(0), rtol=2e-6, atol=2e-6 * float(yd.abs().sum(0).max()))
        assert torch.allclose(sums[Cout:], (yd * yd).sum(0), rtol=2e-6)
    # backward
    dy = dev(rs.normal(size=(B, Lout, Cout)) * 1e-4)
    (yr * dy.to(rdev).double()).sum().backward()
    dys, dy_amax = split_h(dy)
    dx = torch.full((B, L, Cin), float('nan'), device='cuda')
    L_.call('gn_conv1d_dgrad_f16x2', L_.ptr(dys, H), L_.ptr(dy_amax), L_.ptr(wk, H), L_.ptr(w_amax), None, L_.ptr(dx), None,
            None, B, L, Cin, Lout, Cout, k, s, pad, L_.ACT_NONE, 0.0, st)
    torch.cuda.synchronize()
    assert_close(dx.cpu().numpy(), xr.grad.cpu().numpy(), 'f16x2 conv dgrad', TOL)
    cs = torch.full((Cin,), float('nan'), device='cuda')
    dx_amax = torch.full((1,), float('nan'), device='cuda')
    L_.call('gn_conv1d_dgrad_f16x2', L_.ptr(dys, H), L_.ptr(dy_amax), L_.ptr(wk, H), L_.ptr(w_amax), L_.ptr(x), L_.ptr(dx),
            L_.ptr(cs), L_.ptr(dx_amax), B, L, Cin, Lout, Cout, k, s, pad, L_.ACT_RELU, 0.0, st)
    ref = xr.grad * (x > 0).to(rdev).double()
    assert_close(dx.cpu().numpy(), ref.cpu().numpy(), 'f16x2 conv dgrad*relu mask', TOL)
    assert dx_amax.item() == dx.abs().max().item()
    assert_close(cs.cpu().numpy(), ref.sum((0, 1)).cpu().numpy(), 'f16x2 conv dgrad column sums', 1e-5)
    if Cin % 128 == 0 or (Cin == 64 and Cout % 128 == 0):
        dw = torch.full((k, Cin, Cout), float('nan'), device='cuda')
        db = torch.full((Cout,), float('nan'), device='cuda')
        L_.call('gn_conv1d_wgrad_f16x2', L_.ptr(xs, H), L_.ptr(x_amax), L_.ptr(dys, H), L_.ptr(dy_amax), L_.ptr(dy), L_.ptr(dw),
                L_.ptr(db), B, L, Cin, Lout, Cout, k, s, pad, st)
        torch.cuda.synchronize()
        assert_close(dw.cpu().numpy(), wr.grad.cpu().numpy(), 'f16x2 conv wgrad', TOL)
        assert_close(db.cpu().numpy(), dy.double().sum((0, 1)).cpu().numpy(), 'f16x2 conv bias grad', 1e-5)


def test_f16x2_accuracy_at_baseline_size(capsys):
    """Largest layer of the CNN point estimator at BASELINE batch (see test_tc3_accuracy_at_baseline_size): observed
    errors of the three kernels next to bf16x3's and the float32 SIMT kernel's on the same inputs."""
    from gennet_b200 import _lib as L_
    B, L, Cin, Cout, k, s = 64, 1018, 512, 1024, 5, 2
    rs = np.random.RandomState(5)
    x = torch.relu(dev(rs.normal(size=(B, L, Cin))))
    w = dev(rs.normal(size=(k, Cin, Cout)) / math.sqrt(k * Cin))
    st = L_.stream()
    wk = torch.empty(2, k, Cin, Cout, dtype=H, device='cuda')
    wt = torch.empty(2, k, Cout, Cin, dtype=H, device='cuda')
    w_amax = torch.empty(1, device='cuda')
    L_.call('gn_conv_w_split_f16x2', L_.ptr(w), L_.ptr(wk, H), L_.ptr(wt, H), L_.ptr(w_amax), k, Cin, Cout, st)
    xr = x.double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    yr = F.conv1d(xr.permute(0, 2, 1), wr.permute(2, 1, 0), None, stride=s).permute(0, 2, 1)
    Lout = yr.shape[1]
    xs, x_amax = split_h(x)
    y = torch.empty(B, Lout, Cout, device='cuda')
    L_.call('gn_conv1d_fwd_f16x2', L_.ptr(xs, H), L_.ptr(x_amax), L_.ptr(wt, H), L_.ptr(w_amax), None, L_.ptr(y), None, B, L,
            Cin, Lout, Cout, k, s, 0, L_.ACT_NONE, 0.0, st)
    dy = dev(rs.normal(size=(B, Lout, Cout)) + 0.3)
    (yr * dy.double()).sum().backward()
    dys, dy_amax = split_h(dy)
    dx = torch.empty(B, L, Cin, device='cuda')
    L_.call('gn_conv1d_dgrad_f16x2', L_.ptr(dys, H), L_.ptr(dy_amax), L_.ptr(wk, H), L_.ptr(w_amax), None, L_.ptr(dx), None,
            None, B, L, Cin, Lout, Cout, k, s, 0, L_.ACT_NONE, 0.0, st)
    dw = torch.empty(k, Cin, Cout, device='cuda')
    L_.call('gn_conv1d_wgrad_f16x2', L_.ptr(xs, H), L_.ptr(x_amax), L_.ptr(dys, H), L_.ptr(dy_amax), None, L_.ptr(dw), None,
            B, L, Cin, Lout, Cout, k, s, 0, st)
    torch.cuda.synchronize()

    def err(a, b):
        return ((a.double() - b).abs().max() / b.abs().max()).item()
    e = {'fwd': err(y, yr.detach()), 'dgrad': err(dx, xr.grad), 'wgrad': err(dw, wr.grad)}
    with capsys.disabled():
        print('\n[f16x2 accuracy, conv 512->1024 k5 s2, B=64] ' + ' '.join('%s=%.2e' % kv for kv in e.items()))
    assert e['fwd'] <= 1e-5 and e['dgrad'] <= 1e-5 and e['wgrad'] <= 1e-5, e


@pytest.mark.parametrize('M,K,N', [(16, 100, 512), (128, 100, 4096), (24, 16128, 1024), (200, 256, 128), (8, 64, 256)])
def test_f16x2_dense(M, K, N):
    from gennet_b200 import _lib as L_
    rs = np.random.RandomState(M + K + N)
    x, w, b = dev(rs.normal(size=(M, K)) * 40.0), dev(rs.normal(size=(K, N)) / math.sqrt(K)), dev(rs.normal(size=N))
    dy = dev(rs.normal(size=(M, N)) * 1e-5)
    Kp = 64 if (K <= 64 and N % 128 == 0) else -(-K // 128) * 128
    st = L_.stream()
    xs = torch.full((2, M, Kp), float('nan'), dtype=H, device='cuda')
    x_amax = torch.full((1,), float('nan'), device='cuda')
    L_.call('gn_split_pad_f32_f16x2', L_.ptr(x), L_.ptr(xs, H), L_.ptr(x_amax), M, K, Kp, st)
    assert torch.equal(xs[:, :, :K], split_h(x)[0]) and (xs[:, :, K:] == 0).all() and x_amax.item() == x.abs().max().item()
    wk = torch.full((2, Kp, N), float('nan'), dtype=H, device='cuda')
    wt = torch.full((2, N, Kp), float('nan'), dtype=H, device='cuda')
    w_amax = torch.full((1,), float('nan'), device='cuda')
    L_.call('gn_dense_w_split_f16x2', L_.ptr(w), L_.ptr(wk, H), L_.ptr(wt, H), L_.ptr(w_amax), K, Kp, N, st)
    assert torch.equal(wk[:, :K], split_h(w)[0]) and (wk[:, K:] == 0).all() and torch.equal(wt, wk.permute(0, 2, 1).contiguous())
    xr, wr = x.double(), w.double()
    y = torch.full((M, N), float('nan'), device='cuda')
    L_.call('gn_dense_fwd_f16x2', L_.ptr(xs, H), L_.ptr(x_amax), L_.ptr(wt, H), L_.ptr(w_amax), L_.ptr(b), L_.ptr(y), M, Kp, N,
            L_.ACT_NONE, 0.0, st)
    torch.cuda.synchronize()
    assert_close(y.cpu().numpy(), (xr @ wr + b.double()).cpu().numpy(), 'f16x2 dense fwd', TOL)
    dys, dy_amax = split_h(dy)
    dw = torch.full((K, N), float('nan'), device='cuda')
    db = torch.full((N,), float('nan'), device='cuda')
    guard = torch.full((64,), 7.0, device='cuda')
    L_.call('gn_dense_wgrad_f16x2', L_.ptr(xs, H), L_.ptr(x_amax), L_.ptr(dys, H), L_.ptr(dy_amax), L_.ptr(dy), L_.ptr(dw),
            L_.ptr(db), M, K, N, Kp, st)
    torch.cuda.synchronize()
    assert_close(dw.cpu().numpy(), (xr.t() @ dy.double()).cpu().numpy(), 'f16x2 dense wgrad', TOL)
    assert_close(db.cpu().numpy(), dy.double().sum(0).cpu().numpy(), 'f16x2 dense bias grad', 1e-5)
    assert (guard == 7.0).all()
    if K % 64 == 0:
        dx = torch.full((M, K), float('nan'), device='cuda')
        cs = torch.full((K,), float('nan'), device='cuda')
        xpos = torch.relu(x).contiguous()
        L_.call('gn_dense_dgrad_f16x2', L_.ptr(dys, H), L_.ptr(dy_amax), L_.ptr(wk, H), L_.ptr(w_amax), L_.ptr(xpos), L_.ptr(dx),
                L_.ptr(cs), M, K, N, L_.ACT_RELU, 0.0, st)
        ref = (dy.double() @ wr.t()) * (xpos > 0)
        assert_close(dx.cpu().numpy(), ref.cpu().numpy(), 'f16x2 dense dgrad*mask', TOL)
        assert_close(cs.cpu().numpy(), ref.sum(0).cpu().numpy(), 'f16x2 dense dgrad column sums', 1e-5)
