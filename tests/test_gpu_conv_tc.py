"""GPU parity of the tcgen05/TMA Conv1D kernels (bf16 in, fp32 accumulate) against float64 math on the SAME
bf16-rounded inputs: products of bf16 values are exact in fp32, so only the accumulation order and the bf16
rounding of the stored outputs differ (tolerance 2^-8 of the tensor scale for bf16 outputs, 1e-4 for the fp32
weight gradient)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import keras_oracle as ko
from tests.parity_cases import assert_close, case_seed

pytestmark = pytest.mark.gpu


def bf(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda().to(torch.bfloat16).contiguous()


CASES = [
    # B, L, Cin, Cout, k, s, padding
    (2, 200, 64, 128, 5, 1, 'valid'),     # q tower conv2 geometry (ragged tail tile)
    (2, 256, 64, 128, 5, 2, 'valid'),     # mc tower conv2 geometry (traversal stride 2)
    (3, 130, 128, 256, 5, 1, 'same'),     # 'same' padding: negative / overflowing TMA coordinates
    (2, 253, 256, 512, 5, 2, 'valid'),    # odd length, stride 2
    (2, 64, 512, 64, 5, 1, 'same'),       # BN = 64 path, many K blocks
    (1, 1018, 512, 1024, 5, 2, 'valid'),  # q tower conv5 geometry
    # many tiles per persistent CTA: the division-free tile walker carries through every digit (n-tile, m-tile,
    # parity class, sample) and both TMEM buffers / the slab ring wrap many times
    (300, 600, 64, 128, 5, 1, 'valid'),   # 1500 forward tiles, BN = 128 / 64
    (96, 515, 256, 512, 5, 2, 'same'),    # 576 tiles: two n-tiles forward, two parity classes in the data gradient
]


@pytest.mark.parametrize('case', CASES)
def test_tc_conv_fwd_dgrad_wgrad(case):
    from gennet_b200 import _lib as L_
    B, L, Cin, Cout, k, s, padding = case
    rs = np.random.RandomState(case_seed(case))
    x = bf(rs.normal(size=(B, L, Cin)))
    w32 = (rs.normal(size=(k, Cin, Cout)) / math.sqrt(k * Cin)).astype(np.float32)
    bias = torch.as_tensor(rs.normal(size=Cout).astype(np.float32)).cuda()
    w_dev = torch.as_tensor(w32).cuda()
    wk = torch.empty(k, Cin, Cout, dtype=torch.bfloat16, device='cuda')
    wt = torch.empty(k, Cout, Cin, dtype=torch.bfloat16, device='cuda')
    st = L_.stream()
    L_.call('gn_conv_w_to_bf16', L_.ptr(w_dev), L_.ptr(wk, torch.bfloat16), L_.ptr(wt, torch.bfloat16), k, Cin, Cout, st)
    assert torch.equal(wt, wk.permute(0, 2, 1).contiguous())
    # float64 reference on the bf16-rounded operands
    xr = x.float().cpu().double().requires_grad_(True)
    wr = wk.float().cpu().double().requires_grad_(True)
    br = bias.cpu().double()
    xp = xr.permute(0, 2, 1)
    pad = 0
    if padding == 'same':
        pl, pr = ko.same_pad(L, k, s)
        xp = F.pad(xp, (pl, pr))
        pad = pl
    yr = F.conv1d(xp, wr.permute(2, 1, 0), br, stride=s).permute(0, 2, 1)
    Lout = yr.shape[1]
    y = torch.empty(B, Lout, Cout, dtype=torch.bfloat16, device='cuda')
    L_.call('gn_conv1d_fwd_bf16', L_.ptr(x, torch.bfloat16), L_.ptr(wt, torch.bfloat16), L_.ptr(bias), L_.ptr(y, torch.bfloat16),
            B, L, Cin, Lout, Cout, k, s, pad, L_.ACT_NONE, 0.0, st)
    torch.cuda.synchronize()
    assert_close(y.float().cpu().numpy(), yr.detach().numpy(), 'tc conv fwd', 2 ** -8)
    # fused ReLU epilogue
    L_.call('gn_conv1d_fwd_bf16', L_.ptr(x, torch.bfloat16), L_.ptr(wt, torch.bfloat16), L_.ptr(bias), L_.ptr(y, torch.bfloat16),
            B, L, Cin, Lout, Cout, k, s, pad, L_.ACT_RELU, 0.0, st)
    assert_close(y.float().cpu().numpy(), torch.relu(yr).detach().numpy(), 'tc conv fwd+relu', 2 ** -8)
    # backward
    dy = bf(rs.normal(size=(B, Lout, Cout)))
    (yr * dy.float().cpu().double()).sum().backward()
    dx = torch.full((B, L, Cin), float('nan'), dtype=torch.bfloat16, device='cuda')
    L_.call('gn_conv1d_dgrad_bf16', L_.ptr(dy, torch.bfloat16), L_.ptr(wk, torch.bfloat16), None, L_.ptr(dx, torch.bfloat16),
            None, B, L, Cin, Lout, Cout, k, s, pad, L_.ACT_NONE, 0.0, st)
    torch.cuda.synchronize()
    assert_close(dx.float().cpu().numpy(), xr.grad.numpy(), 'tc conv dgrad', 2 ** -8)
    # dgrad with the fused ReLU mask of the conv input, and the fused per-channel sum of dx (= bias gradient of the
    # layer that produced x): must equal the sum of the bf16 values actually stored
    cs = torch.full((Cin,), float('nan'), device='cuda')
    L_.call('gn_conv1d_dgrad_bf16', L_.ptr(dy, torch.bfloat16), L_.ptr(wk, torch.bfloat16), L_.ptr(x, torch.bfloat16),
            L_.ptr(dx, torch.bfloat16), L_.ptr(cs), B, L, Cin, Lout, Cout, k, s, pad, L_.ACT_RELU, 0.0, st)
    mask = (x.float().cpu().numpy() > 0)
    assert_close(dx.float().cpu().numpy(), xr.grad.numpy() * mask, 'tc conv dgrad*relu mask', 2 ** -8)
    assert_close(cs.cpu().numpy(), dx.float().cpu().double().sum((0, 1)).numpy(), 'tc conv dgrad column sums', 1e-5)
    # column sums without a mask (rows in the right padding of a SAME convolution must not leak in)
    L_.call('gn_conv1d_dgrad_bf16', L_.ptr(dy, torch.bfloat16), L_.ptr(wk, torch.bfloat16), None, L_.ptr(dx, torch.bfloat16),
            L_.ptr(cs), B, L, Cin, Lout, Cout, k, s, pad, L_.ACT_NONE, 0.0, st)
    assert_close(cs.cpu().numpy(), dx.float().cpu().double().sum((0, 1)).numpy(), 'tc conv dgrad column sums (no mask)', 1e-5)
    dw = torch.empty(k, Cin, Cout, device='cuda')
    db = torch.empty(Cout, device='cuda')
    L_.call('gn_conv1d_wgrad_bf16', L_.ptr(x, torch.bfloat16), L_.ptr(dy, torch.bfloat16), L_.ptr(dw), L_.ptr(db), B, L, Cin, Lout,
            Cout, k, s, pad, st)
    torch.cuda.synchronize()
    assert_close(dw.cpu().numpy(), wr.grad.numpy(), 'tc conv wgrad', 1e-4)
    assert_close(db.cpu().numpy(), dy.float().cpu().double().sum((0, 1)).numpy(), 'tc conv bias grad', 1e-4)


@pytest.mark.parametrize('case', [(3, 200, 1, 64, 5, 2, 'same'), (2, 256, 1, 64, 5, 1, 'same'), (2, 100, 2, 128, 5, 2, 'same'),
                                  (2, 77, 1, 32, 3, 1, 'valid'), (3, 128, 2, 512, 5, 2, 'same')])
def test_first_layer_kernels(case):
    from gennet_b200 import _lib as L_
    B, L, Cin, Cout, k, s, padding = case
    rs = np.random.RandomState(case_seed(case))
    x = torch.as_tensor(rs.normal(size=(B, L, Cin)).astype(np.float32)).cuda()
    w = torch.as_tensor((rs.normal(size=(k, Cin, Cout)) / math.sqrt(k * Cin)).astype(np.float32)).cuda()
    bias = torch.as_tensor(rs.normal(size=Cout).astype(np.float32)).cuda()
    xr = x.cpu().double()
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
    y = torch.empty(B, Lout, Cout, dtype=torch.bfloat16, device='cuda')
    st = L_.stream()
    L_.call('gn_conv1d_smallcin_fwd_bf16', L_.ptr(x), L_.ptr(w), L_.ptr(bias), L_.ptr(y, torch.bfloat16), B, L, Cin, Lout, Cout,
            k, s, pad, L_.ACT_RELU, 0.0, st)
    assert_close(y.float().cpu().numpy(), torch.relu(yr).detach().numpy(), 'smallcin fwd', 2 ** -8)
    dy = bf(rs.normal(size=(B, Lout, Cout)))
    (yr * dy.float().cpu().double()).sum().backward()
    dw = torch.full((k, Cin, Cout), float('nan'), device='cuda')
    db = torch.full((Cout,), float('nan'), device='cuda')
    L_.call('gn_conv1d_smallcin_wgrad_bf16', L_.ptr(x), L_.ptr(dy, torch.bfloat16), L_.ptr(dw), L_.ptr(db), B, L, Cin, Lout, Cout,
            k, s, pad, st)
    assert_close(dw.cpu().numpy(), wr.grad.numpy(), 'smallcin wgrad', 1e-4)
    assert_close(db.cpu().numpy(), br.grad.numpy(), 'smallcin bias grad', 1e-4)
    # data gradient through the small-Cin convolution (generator step through the frozen discriminator)
    xg = x.cpu().double().requires_grad_(True)
    xpg = xg.permute(0, 2, 1)
    if padding == 'same':
        xpg = F.pad(xpg, (pl, pr))
    yg = F.conv1d(xpg, wr.detach().permute(2, 1, 0), br.detach(), stride=s).permute(0, 2, 1)
    (yg * dy.float().cpu().double()).sum().backward()
    dx = torch.full((B, L, Cin), float('nan'), device='cuda')
    L_.call('gn_conv1d_smallcin_dgrad_bf16', L_.ptr(dy, torch.bfloat16), L_.ptr(w), L_.ptr(dx), B, L, Cin, Lout, Cout, k, s,
            pad, st)
    assert_close(dx.cpu().numpy(), xg.grad.numpy(), 'smallcin dgrad', 1e-5)


@pytest.mark.parametrize('M,K,N', [(5, 64000, 1), (3, 4096, 2), (9, 1000 * 8, 4)])
def test_dense_small_bf16(M, K, N):
    from gennet_b200 import _lib as L_
    rs = np.random.RandomState(M * 7 + N)
    x = bf(np.maximum(rs.normal(size=(M, K)), 0))            # post-ReLU features
    w = torch.as_tensor((rs.normal(size=(K, N)) / math.sqrt(K)).astype(np.float32)).cuda()
    b = torch.as_tensor(rs.normal(size=N).astype(np.float32)).cuda()
    dy = torch.as_tensor(rs.normal(size=(M, N)).astype(np.float32)).cuda()
    xr, wr = x.float().cpu().double(), w.cpu().double()
    st = L_.stream()
    y = torch.empty(M, N, device='cuda')
    L_.call('gn_dense_small_fwd_bf16', L_.ptr(x, torch.bfloat16), L_.ptr(w), L_.ptr(b), L_.ptr(y), M, K, N, L_.ACT_NONE, 0.0, st)
    assert_close(y.cpu().numpy(), (xr @ wr + b.cpu().double()).numpy(), 'dense small fwd', 1e-5)
    dx = torch.empty(M, K, dtype=torch.bfloat16, device='cuda')
    C = 64
    cs = torch.full((C,), float('nan'), device='cuda')
    L_.call('gn_dense_small_dgrad_bf16', L_.ptr(dy), L_.ptr(w), L_.ptr(x, torch.bfloat16), L_.ptr(dx, torch.bfloat16),
            L_.ptr(cs), C, M, K, N, L_.ACT_RELU, 0.0, st)
    ref = (dy.cpu().double() @ wr.t()) * (xr > 0)
    assert_close(dx.float().cpu().numpy(), ref.numpy(), 'dense small dgrad*mask', 2 ** -8)
    assert_close(cs.cpu().numpy(), dx.float().cpu().double().reshape(M, K // C, C).sum((0, 1)).numpy(),
                 'dense small dgrad column sums', 1e-5)
    L_.call('gn_dense_small_dgrad_bf16', L_.ptr(dy), L_.ptr(w), None, L_.ptr(dx, torch.bfloat16), None, 0, M, K, N,
            L_.ACT_NONE, 0.0, st)
    assert_close(dx.float().cpu().numpy(), (dy.cpu().double() @ wr.t()).numpy(), 'dense small dgrad', 2 ** -8)
    dw = torch.full((K, N), float('nan'), device='cuda')
    db = torch.full((N,), float('nan'), device='cuda')
    L_.call('gn_dense_small_wgrad_bf16', L_.ptr(x, torch.bfloat16), L_.ptr(dy), L_.ptr(dw), L_.ptr(db), M, K, N, st)
    assert_close(dw.cpu().numpy(), (xr.t() @ dy.cpu().double()).numpy(), 'dense small wgrad', 1e-5)
    assert_close(db.cpu().numpy(), dy.cpu().double().sum(0).numpy(), 'dense small bias grad', 1e-6)


@pytest.mark.parametrize('case', [(3, 300, 1024, 5, 'same'), (2, 97, 64, 5, 'valid'), (2, 64, 256, 3, 'same')])
def test_cout1_conv_kernels(case):
    """Last generator convolution (Cout = 1, stride 1): fwd / dgrad / wgrad vs torch float64 on bf16-rounded inputs."""
    from gennet_b200 import _lib as L_
    B, L, Cin, k, padding = case
    rs = np.random.RandomState(case_seed(case))
    x = bf(rs.normal(size=(B, L, Cin)))
    w = torch.as_tensor((rs.normal(size=(k, Cin, 1)) / math.sqrt(k * Cin)).astype(np.float32)).cuda()
    bias = torch.as_tensor(rs.normal(size=1).astype(np.float32)).cuda()
    xr = x.float().cpu().double().requires_grad_(True)
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
    L_.call('gn_conv1d_cout1_fwd_bf16', L_.ptr(x, torch.bfloat16), L_.ptr(w), L_.ptr(bias), L_.ptr(y), B, L, Cin, Lout, k, pad, st)
    assert_close(y.cpu().numpy(), yr.detach().numpy(), 'cout1 fwd', 1e-5)
    dy = torch.as_tensor(rs.normal(size=(B, Lout, 1)).astype(np.float32)).cuda()
    (yr * dy.cpu().double()).sum().backward()
    dx = torch.full((B, L, Cin), float('nan'), dtype=torch.bfloat16, device='cuda')
    L_.call('gn_conv1d_cout1_dgrad_bf16', L_.ptr(dy), L_.ptr(w), L_.ptr(dx, torch.bfloat16), B, L, Cin, Lout, k, pad, st)
    assert_close(dx.float().cpu().numpy(), xr.grad.numpy(), 'cout1 dgrad', 2 ** -8)
    dw = torch.full((k, Cin, 1), float('nan'), device='cuda')
    db = torch.full((1,), float('nan'), device='cuda')
    L_.call('gn_conv1d_cout1_wgrad_bf16', L_.ptr(x, torch.bfloat16), L_.ptr(dy), L_.ptr(dw), L_.ptr(db), B, L, Cin, Lout, k, pad, st)
    assert_close(dw.cpu().numpy(), wr.grad.numpy(), 'cout1 wgrad', 1e-5)
    assert_close(db.cpu().numpy(), br.grad.numpy(), 'cout1 bias grad', 1e-5)


def test_upsample_bf16_kernels():
    from gennet_b200 import _lib as L_
    rs = np.random.RandomState(0)
    B, L, C = 3, 37, 64
    x = bf(rs.normal(size=(B, L, C)))
    y = torch.empty(B, 2 * L, C, dtype=torch.bfloat16, device='cuda')
    st = L_.stream()
    L_.call('gn_upsample1d_fwd_bf16', L_.ptr(x, torch.bfloat16), L_.ptr(y, torch.bfloat16), B, L, C, 2, st)
    assert torch.equal(y, x.repeat_interleave(2, dim=1))
    dy = bf(rs.normal(size=(B, 2 * L, C)))
    dx = torch.empty(B, L, C, dtype=torch.bfloat16, device='cuda')
    L_.call('gn_upsample1d_bwd_bf16', L_.ptr(dy, torch.bfloat16), L_.ptr(dx, torch.bfloat16), B, L, C, 2, st)
    ref = dy.float().reshape(B, L, 2, C).sum(2)
    assert_close(dx.float().cpu().numpy(), ref.cpu().numpy(), 'upsample bwd', 2 ** -8)
