import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import parity_cases as pc
from gennet_b200 import nn
from oracle import keras_oracle as ko
(g, d, dg, sub_g), (og, od, ocomp, osub), z, sX, sy, ny = pc.burst_case(64, 8)
# record product dy per layer
rec = {}
for l in dg.all_layers():
    orig = l.backward
    def mk(l, orig):
        def bw(dy, ctx, need_dx=True):
            rec[l.name] = dy.detach().cpu().numpy().copy()
            return orig(dy, ctx, need_dx)
        return bw
    l.backward = mk(l, l.backward)
# oracle: capture outputs with retain_grad
outs = {}
for l in ocomp.all_layers():
    of = l.forward
    def mk2(l, of):
        def fw(x, training, noise):
            y = of(x, training, noise)
            if y.requires_grad:
                y.retain_grad()
            outs[l.name] = y
            return y
        return fw
    l.forward = mk2(l, l.forward)
noise = pc.draw_noise(ocomp, z, 0)
pn = pc.map_noise(noise, ocomp, dg)
outs.clear()
# need grads wrt intermediate: recompute with autograd
xin = torch.as_tensor(z, dtype=torch.float64)
outs.clear()
o = ocomp.forward(xin, True, dict(noise))
yt = torch.ones(8, 1, dtype=torch.float64)
loss = ko.binary_crossentropy(yt, o).mean()
loss.backward()
rp = dg.train_on_batch(z, [1]*8, _noise=pn)
ol = ocomp.all_layers(); plr = dg.all_layers()
for a, b in zip(ol, plr):
    if b.name in rec and a.name in outs and outs[a.name].grad is not None:
        ref = outs[a.name].grad.numpy(); got = rec[b.name]
        if got.size != ref.size:
            print(a.name, 'size mismatch (fused)'); continue
        got = got.reshape(ref.shape)
        err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30)
        print('%-22s %-22s dy err %.2e scale %.2e' % (a.name, b.name, err, np.abs(ref).max()))
