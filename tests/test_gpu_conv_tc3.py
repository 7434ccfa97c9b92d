"""GPU parity of the split-bf16 tcgen05 Conv1D kernels (conv1d_tc3.cu) against float64 math on the SAME float32
inputs.  With three planes the kernel is a float32-class convolution (six bf16 plane products per K step, exact in
the fp32 accumulator; the accumulator itself truncates, ~0.3 ulp per tcgen05.mma, which is why the kernel keeps the
large x0*w0 products in their own accumulator): tolerance 6e-6 of the tensor scale at test size and 1e-5 at BASELINE
size (K up to 5 * 512 forward, 64 * 507 positions in the weight gradient), far inside north_star's rtol 1e-4.
Two planes: 1e-4."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import keras_oracle as ko
from tests.parity_cases import assert_close, case_seed

pytestmark = pytest.mark.gpu
BF = torch.bfloat16

CASES = [
    # B, L, Cin, Cout, k, s, padding
    (2, 200, 64, 128, 5, 1, 'valid'),     # q tower conv2 geometry (ragged tail tile), wgrad SWAP form
    (2, 256, 64, 128, 5, 2, 'valid'),     # mc tower conv2 geometry (traversal stride 2)
    (3, 130, 128, 256, 5, 1, 'same'),     # 'same' padding: negative / overflowing TMA coordinates
    (2, 253, 256, 512, 5, 2, 'valid'),    # odd length, stride 2
    (2, 64, 512, 64, 5, 1, 'same'),       # BN = 64 path, many K blocks
    (1, 1018, 512, 1024, 5, 2, 'valid'),  # q tower conv5 geometry
    (300, 600, 64, 128, 5, 1, 'valid'),   # many tiles per persistent CTA
    (96, 515, 256, 512, 5, 2, 'same'),    # two n-tiles forward, two parity classes in the data gradient
]
TOL = {3: 6e-6, 2: 1e-4, 1: 2 ** -7}


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def split(x, nc):
    from gennet_b200 import _lib as L_
    p = torch.empty((nc,) + tuple(x.shape), dtype=BF, device='cuda')
    L_.call('gn_split_f32_bf16', L_.ptr(x), L_.ptr(p, BF), x.numel(), nc, L_.stream())
    return p


def test_split_planes_reconstruct():
    rs = np.random.RandomState(0)
    x = dev(rs.normal(size=(4096,)) * np.exp(rs.uniform(-20, 20, 4096)))
    for nc, tol in ((3, 2 ** -23), (2, 2 ** -16), (1, 2 ** -8)):
        p = split(x, nc)
        rec = p.double().sum(0)
        rel = ((rec - x.double()).abs() / x.double().abs()).max().item()
        assert rel <= tol, (nc, rel)
        assert torch.equal(p[0], x.to(BF))


@pytest.mark.parametrize('nc', [3, 2])
@pytest.mark.parametrize('case', CASES)
def test_tc3_conv_fwd_dgrad_wgrad(case, nc):
    from gennet_b200 import _lib as L_
    B, L, Cin, Cout, k, s, padding = case
    if nc == 2 and B > 3:
        pytest.skip('two-plane variant: small cases only')
    tol = TOL[nc]
    rs = np.random.RandomState(case_seed(case))
    x = dev(rs.normal(size=(B, L, Cin)))
    w = dev(rs.normal(size=(k, Cin, Cout)) / math.sqrt(k * Cin))
    bias = dev(rs.normal(size=Cout))
    st = L_.stream()
    wk = torch.empty(nc, k, Cin, Cout, dtype=BF, device='cuda')
    wt = torch.empty(nc, k, Cout, Cin, dtype=BF, device='cuda')
    L_.call('gn_conv_w_split_bf16', L_.ptr(w), L_.ptr(wk, BF), L_.ptr(wt, BF), k, Cin, Cout, nc, st)
    assert torch.equal(wt, wk.permute(0, 1, 3, 2).contiguous())
    assert torch.equal(wk, split(w, nc))
    big = B * L * Cin > 4e6
    rdev = 'cuda' if big else 'cpu'          # float64 reference (torch double; on the device for the large cases)
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
    xs = split(x, nc)
    y = torch.full((B, Lout, Cout), float('nan'), device='cuda')
    ys = torch.full((nc, B, Lout, Cout), float('nan'), dtype=BF, device='cuda')
    L_.call('gn_conv1d_fwd_bf16x3', L_.ptr(xs, BF), L_.ptr(wt, BF), L_.ptr(bias), L_.ptr(y), L_.ptr(ys, BF), B, L, Cin,
            Lout, Cout, k, s, pad, L_.ACT_NONE, 0.0, nc, st)
    torch.cuda.synchronize()
    assert_close(y.cpu().numpy(), yr.detach().cpu().numpy(), 'tc3 conv fwd', tol)
    assert torch.equal(ys, split(y, nc)), 'planes written by the epilogue != split of the float32 result'
    # fused ReLU epilogue, planes only
    L_.call('gn_conv1d_fwd_bf16x3', L_.ptr(xs, BF), L_.ptr(wt, BF), L_.ptr(bias), None, L_.ptr(ys, BF), B, L, Cin,
            Lout, Cout, k, s, pad, L_.ACT_RELU, 0.0, nc, st)
    assert_close(ys.double().sum(0).cpu().numpy(), torch.relu(yr).detach().cpu().numpy(), 'tc3 conv fwd+relu (planes)', tol)
    # backward
    dy = dev(rs.normal(size=(B, Lout, Cout)))
    (yr * dy.to(rdev).double()).sum().backward()
    dys = split(dy, nc)
    dx = torch.full((B, L, Cin), float('nan'), device='cuda')
    L_.call('gn_conv1d_dgrad_bf16x3', L_.ptr(dys, BF), L_.ptr(wk, BF), None, L_.ptr(dx), None, None, B, L, Cin, Lout, Cout,
            k, s, pad, L_.ACT_NONE, 0.0, nc, st)
    torch.cuda.synchronize()
    assert_close(dx.cpu().numpy(), xr.grad.cpu().numpy(), 'tc3 conv dgrad', tol)
    # fused ReLU mask of the conv input + column sums (bias gradient of the producer) + planes of dx
    cs = torch.full((Cin,), float('nan'), device='cuda')
    dxs = torch.full((nc, B, L, Cin), float('nan'), dtype=BF, device='cuda')
    L_.call('gn_conv1d_dgrad_bf16x3', L_.ptr(dys, BF), L_.ptr(wk, BF), L_.ptr(x), L_.ptr(dx), L_.ptr(dxs, BF), L_.ptr(cs), B, L,
            Cin, Lout, Cout, k, s, pad, L_.ACT_RELU, 0.0, nc, st)
    mask = (x > 0).to(rdev).double()
    ref = (xr.grad * mask)
    assert_close(dx.cpu().numpy(), ref.cpu().numpy(), 'tc3 conv dgrad*relu mask', tol)
    assert torch.equal(dxs, split(dx, nc))
    assert_close(cs.cpu().numpy(), ref.sum((0, 1)).cpu().numpy(), 'tc3 conv dgrad column sums', max(tol, 1e-5))
    if Cin % 128 == 0 or (Cin == 64 and Cout % 128 == 0):
        dw = torch.full((k, Cin, Cout), float('nan'), device='cuda')
        db = torch.full((Cout,), float('nan'), device='cuda')
        L_.call('gn_conv1d_wgrad_bf16x3', L_.ptr(xs, BF), L_.ptr(dys, BF), L_.ptr(dy), L_.ptr(dw), L_.ptr(db), B, L, Cin,
                Lout, Cout, k, s, pad, nc, st)
        torch.cuda.synchronize()
        assert_close(dw.cpu().numpy(), wr.grad.cpu().numpy(), 'tc3 conv wgrad', max(tol, 4e-6))
        assert_close(db.cpu().numpy(), dy.double().sum((0, 1)).cpu().numpy(), 'tc3 conv bias grad', 1e-5)


def test_tc3_accuracy_at_baseline_size(capsys):
    """Largest layer of the CNN point estimator at BASELINE batch (conv 512 -> 1024, k 5, stride 2, 64 of the 512
    samples): forward K = 2560, weight gradient over 64 * 507 positions per accumulator chunk.  Reports the observed
    errors (the evidence for the accumulation behaviour of the fp32 tensor-memory adder) and bounds them at 1e-5."""
    from gennet_b200 import _lib as L_
    B, L, Cin, Cout, k, s, nc = 64, 1018, 512, 1024, 5, 2, 3
    rs = np.random.RandomState(5)
    x = torch.relu(dev(rs.normal(size=(B, L, Cin))))            # post-ReLU activations: non-zero mean, like the real net
    w = dev(rs.normal(size=(k, Cin, Cout)) / math.sqrt(k * Cin))
    st = L_.stream()
    wk = torch.empty(nc, k, Cin, Cout, dtype=BF, device='cuda')
    wt = torch.empty(nc, k, Cout, Cin, dtype=BF, device='cuda')
    L_.call('gn_conv_w_split_bf16', L_.ptr(w), L_.ptr(wk, BF), L_.ptr(wt, BF), k, Cin, Cout, nc, st)
    xr = x.double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    yr = F.conv1d(xr.permute(0, 2, 1), wr.permute(2, 1, 0), None, stride=s).permute(0, 2, 1)
    Lout = yr.shape[1]
    xs = split(x, nc)
    y = torch.empty(B, Lout, Cout, device='cuda')
    L_.call('gn_conv1d_fwd_bf16x3', L_.ptr(xs, BF), L_.ptr(wt, BF), None, L_.ptr(y), None, B, L, Cin, Lout, Cout, k, s, 0,
            L_.ACT_NONE, 0.0, nc, st)
    dy = dev(rs.normal(size=(B, Lout, Cout)) + 0.3)
    (yr * dy.double()).sum().backward()
    dys = split(dy, nc)
    dx = torch.empty(B, L, Cin, device='cuda')
    L_.call('gn_conv1d_dgrad_bf16x3', L_.ptr(dys, BF), L_.ptr(wk, BF), None, L_.ptr(dx), None, None, B, L, Cin, Lout, Cout,
            k, s, 0, L_.ACT_NONE, 0.0, nc, st)
    dw = torch.empty(k, Cin, Cout, device='cuda')
    L_.call('gn_conv1d_wgrad_bf16x3', L_.ptr(xs, BF), L_.ptr(dys, BF), None, L_.ptr(dw), None, B, L, Cin, Lout, Cout, k, s, 0,
            nc, st)
    torch.cuda.synchronize()
    # the float32 SIMT kernels on the same inputs, for comparison
    y32 = torch.empty_like(y)
    L_.call('gn_conv1d_fwd_f32', L_.ptr(x), L_.ptr(w), None, L_.ptr(y32), B, L, Cin, Lout, Cout, k, s, 0, 1, L_.ACT_NONE, 0.0, st)
    dw32 = torch.empty_like(dw)
    L_.call('gn_conv1d_wgrad_f32', L_.ptr(x), L_.ptr(dy), L_.ptr(dw32), None, B, L, Cin, Lout, Cout, k, s, 0, 1, st)
    torch.cuda.synchronize()

    def err(a, b):
        return ((a.double() - b).abs().max() / b.abs().max()).item()
    e = {'fwd': err(y, yr.detach()), 'dgrad': err(dx, xr.grad), 'wgrad': err(dw, wr.grad),
         'fwd_simt_f32': err(y32, yr.detach()), 'wgrad_simt_f32': err(dw32, wr.grad)}
    with capsys.disabled():
        print('\n[tc3 accuracy, conv 512->1024 k5 s2, B=64] ' + ' '.join('%s=%.2e' % kv for kv in e.items()))
    assert e['fwd'] <= 1e-5 and e['dgrad'] <= 1e-5 and e['wgrad'] <= 1e-5, e


@pytest.mark.parametrize('M,K,N', [(16, 100, 512), (128, 100, 4096), (24, 16128, 1024), (200, 256, 128), (8, 64, 256)])
@pytest.mark.parametrize('nc', [3, 1])
def test_dense_on_split_tensor_core_kernels(M, K, N, nc):
    """Dense fwd / dgrad / wgrad on the split-operand tcgen05 kernels (one-tap convolution over the batch rows), with
    the feature count padded in the planes only (K = 100 -> Kp = 128): vs torch float64."""
    from gennet_b200 import _lib as L_
    tol = 6e-6 if nc == 3 else 2 ** -7
    rs = np.random.RandomState(M + K + N)
    x, w, b = dev(rs.normal(size=(M, K))), dev(rs.normal(size=(K, N)) / math.sqrt(K)), dev(rs.normal(size=N))
    dy = dev(rs.normal(size=(M, N)))
    Kp = 64 if (K <= 64 and N % 128 == 0) else -(-K // 128) * 128
    st = L_.stream()
    xs = torch.full((nc, M, Kp), float('nan'), dtype=BF, device='cuda')
    L_.call('gn_split_pad_f32_bf16', L_.ptr(x), L_.ptr(xs, BF), M, K, Kp, nc, st)
    assert torch.equal(xs[:, :, :K], split(x, nc)) and (xs[:, :, K:] == 0).all()
    wk = torch.full((nc, Kp, N), float('nan'), dtype=BF, device='cuda')
    wt = torch.full((nc, N, Kp), float('nan'), dtype=BF, device='cuda')
    L_.call('gn_dense_w_split_bf16', L_.ptr(w), L_.ptr(wk, BF), L_.ptr(wt, BF), K, Kp, N, nc, st)
    assert torch.equal(wk[:, :K], split(w, nc)) and (wk[:, K:] == 0).all() and torch.equal(wt, wk.permute(0, 2, 1).contiguous())
    xr, wr = (xs.double().sum(0)[:, :K] if nc == 1 else x.double()), (wk.double().sum(0)[:K] if nc == 1 else w.double())
    y = torch.full((M, N), float('nan'), device='cuda')
    L_.call('gn_dense_fwd_bf16x3', L_.ptr(xs, BF), L_.ptr(wt, BF), L_.ptr(b), L_.ptr(y), None, M, Kp, N, L_.ACT_TANH, 0.0, nc, st)
    torch.cuda.synchronize()
    assert_close(y.cpu().numpy(), torch.tanh(xr @ wr + b.double()).cpu().numpy(), 'dense tc fwd', tol)
    dys = split(dy, nc)
    dyr = dys.double().sum(0) if nc == 1 else dy.double()
    dw = torch.full((K, N), float('nan'), device='cuda')
    db = torch.full((N,), float('nan'), device='cuda')
    guard = torch.full((64,), 7.0, device='cuda')                 # allocated right after dw: padded rows must not be written
    L_.call('gn_dense_wgrad_bf16x3', L_.ptr(xs, BF), L_.ptr(dys, BF), L_.ptr(dy), L_.ptr(dw), L_.ptr(db), M, K, N, Kp, nc, st)
    torch.cuda.synchronize()
    assert_close(dw.cpu().numpy(), (xr.t() @ dyr).cpu().numpy(), 'dense tc wgrad', tol)
    assert_close(db.cpu().numpy(), dy.double().sum(0).cpu().numpy(), 'dense tc bias grad', 1e-5)
    assert (guard == 7.0).all()
    if K % 64 == 0:
        dx = torch.full((M, K), float('nan'), device='cuda')
        cs = torch.full((K,), float('nan'), device='cuda')
        xpos = torch.relu(x).contiguous()
        L_.call('gn_dense_dgrad_bf16x3', L_.ptr(dys, BF), L_.ptr(wk, BF), L_.ptr(xpos), L_.ptr(dx), L_.ptr(cs), M, K, N,
                L_.ACT_RELU, 0.0, nc, st)
        ref = (dyr @ wr.t()) * (xpos > 0)
        assert_close(dx.cpu().numpy(), ref.cpu().numpy(), 'dense tc dgrad*mask', tol)
        assert_close(cs.cpu().numpy(), ref.sum(0).cpu().numpy(), 'dense tc dgrad column sums', max(tol, 1e-5))
