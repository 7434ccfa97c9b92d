import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import parity_cases as pc
from gennet_b200 import nn, _lib as L_
BF = torch.bfloat16
nn.set_compute_dtype('bf16x3')
(g, d, dg, sub_g), (og, od, ocomp, osub), z, sX, sy, ny = pc.burst_case(512, 8)
conv = [l for l in g.all_layers() if isinstance(l, nn.Conv1D) and l.params[0].shape == (5, 256, 512)][0]
orig = conv._backward_tc3
orig_call = nn.call
def hook(dy, ctx, need_dx, db_done):
    st = L_.stream()
    w = conv.params[0].data
    k, Cin, Cout = w.shape
    nc = 3
    B, L = dy.shape[0], 512
    wk = torch.empty(nc, k, Cin, Cout, dtype=BF, device='cuda'); wt = torch.empty(nc, k, Cout, Cin, dtype=BF, device='cuda')
    L_.call('gn_conv_w_split_bf16', L_.ptr(w), L_.ptr(wk, BF), L_.ptr(wt, BF), k, Cin, Cout, nc, st)
    cwk, cwt = conv._split_weights(nc)
    print('cached wk == fresh wk:', torch.equal(cwk, wk), ' wt:', torch.equal(cwt, wt), 'in_act', conv.in_act, 'fused_up', conv.fused_up,
          'bias_src', conv.bias_src, 'pad', conv.pad, 's', conv.s, 'Lout', conv.Lout)
    dys = torch.empty(nc, *dy.shape, dtype=BF, device='cuda')
    L_.call('gn_split_f32_bf16', L_.ptr(dy.contiguous()), L_.ptr(dys, BF), dy.numel(), nc, st)
    dx2 = torch.empty(B, L, Cin, device='cuda')
    L_.call('gn_conv1d_dgrad_bf16x3', L_.ptr(dys, BF), L_.ptr(wk, BF), None, L_.ptr(dx2), None, None, B, L, Cin, 512, Cout, k, 1, 2, 0, 0.0, nc, st)
    calls = []
    def spy(name, *a, **kw):
        calls.append((name, a))
        return orig_call(name, *a, **kw)
    nn.call = spy
    dx = orig(dy, ctx, need_dx, db_done)
    nn.call = orig_call
    for name, a in calls:
        print('   call', name, [x for x in a if not isinstance(x, int) or x < 10**6][-14:])
    print('in-flow dx vs standalone dx: %.3e' % ((dx.double() - dx2.double()).abs().max() / dx2.abs().max()).item())
    return dx
conv._backward_tc3 = hook
noise = pc.draw_noise(ocomp, z, 0)
dg.train_on_batch(z, [1] * 8, _noise=pc.map_noise(noise, ocomp, dg))
