import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from tests import parity_cases as pc
from gennet_b200 import nn, _lib as L_
BF = torch.bfloat16
nn.set_compute_dtype('bf16x3')
(g, d, dg, sub_g), (og, od, ocomp, osub), z, sX, sy, ny = pc.burst_case(512, 8)
cap = {}
conv = [l for l in g.all_layers() if isinstance(l, nn.Conv1D) and l.params[0].shape == (5, 256, 512)][0]
orig = conv._backward_tc3
def hook(dy, ctx, need_dx, db_done):
    cap['dy'] = dy.detach().clone(); cap['x'] = conv._x.detach().clone(); cap['xs'] = conv._xs.detach().clone()
    dx = orig(dy, ctx, need_dx, db_done)
    cap['dx'] = dx.detach().clone()
    return dx
conv._backward_tc3 = hook
noise = pc.draw_noise(ocomp, z, 0)
dg.train_on_batch(z, [1] * 8, _noise=pc.map_noise(noise, ocomp, dg))
dy, dx = cap['dy'], cap['dx']
print('dy shape', tuple(dy.shape), 'absmax %.3e' % dy.abs().max().item(), 'min nonzero %.3e' % dy[dy != 0].abs().min().item(),
      'frac zero %.3f' % (dy == 0).float().mean().item(), 'contig', dy.is_contiguous(), 'ptr %% 16 = %d' % (dy.data_ptr() % 16))
w = conv.params[0].data.detach().clone()     # NOTE: after the Adam update; recompute with pre-update weights below
# re-run the kernel standalone on the captured dy with the CURRENT weights and compare with float64
k, Cin, Cout = w.shape
B, L = dy.shape[0], 512
nc = 3
wk = torch.empty(nc, k, Cin, Cout, dtype=BF, device='cuda'); wt = torch.empty(nc, k, Cout, Cin, dtype=BF, device='cuda')
st = L_.stream()
L_.call('gn_conv_w_split_bf16', L_.ptr(w), L_.ptr(wk, BF), L_.ptr(wt, BF), k, Cin, Cout, nc, st)
dys = torch.empty(nc, *dy.shape, dtype=BF, device='cuda')
L_.call('gn_split_f32_bf16', L_.ptr(dy), L_.ptr(dys, BF), dy.numel(), nc, st)
print('split reconstruct err %.3e' % ((dys.double().sum(0) - dy.double()).abs().max() / dy.abs().max()).item())
dx2 = torch.empty(B, L, Cin, device='cuda')
L_.call('gn_conv1d_dgrad_bf16x3', L_.ptr(dys, BF), L_.ptr(wk, BF), None, L_.ptr(dx2), None, None, B, L, Cin, 512, Cout, k, 1, 2, 0, 0.0, nc, st)
def ref(dyv, wv):
    # dx[b,j,ci] = sum_t,co dy[b, j+p-t, co] w[t,ci,co]  (stride 1, p=2) == conv_transpose
    return F.conv_transpose1d(dyv.double().permute(0, 2, 1), wv.double().permute(2, 1, 0), stride=1, padding=2).permute(0, 2, 1)
r = ref(dy, w)
print('standalone dgrad vs f64: %.3e' % ((dx2.double() - r).abs().max() / r.abs().max()).item())
r0 = ref(dys[0].float(), w)
print('f64 with dy plane0 only vs full: %.3e' % ((r0 - r).abs().max() / r.abs().max()).item())
rw0 = ref(dy, wk[0].float())
print('f64 with w plane0 only vs full: %.3e' % ((rw0 - r).abs().max() / r.abs().max()).item())
# scale test: same dy scaled up to O(1)
s = 1.0 / dy.abs().max().item()
dy_s = (dy * s).contiguous()
L_.call('gn_split_f32_bf16', L_.ptr(dy_s), L_.ptr(dys, BF), dy.numel(), nc, st)
L_.call('gn_conv1d_dgrad_bf16x3', L_.ptr(dys, BF), L_.ptr(wk, BF), None, L_.ptr(dx2), None, None, B, L, Cin, 512, Cout, k, 1, 2, 0, 0.0, nc, st)
rs_ = ref(dy_s, w)
print('scaled dy (max 1): dgrad vs f64 %.3e' % ((dx2.double() - rs_).abs().max() / rs_.abs().max()).item())
# random dy of the same shape
dy_r = torch.randn_like(dy) * dy.abs().max()
L_.call('gn_split_f32_bf16', L_.ptr(dy_r), L_.ptr(dys, BF), dy.numel(), nc, st)
L_.call('gn_conv1d_dgrad_bf16x3', L_.ptr(dys, BF), L_.ptr(wk, BF), None, L_.ptr(dx2), None, None, B, L, Cin, 512, Cout, k, 1, 2, 0, 0.0, nc, st)
rr = ref(dy_r, w)
print('random dy same scale: dgrad vs f64 %.3e' % ((dx2.double() - rr).abs().max() / rr.abs().max()).item())
