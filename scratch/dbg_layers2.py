"""After the resync flow: DG step backward outputs per layer in bf16x3 vs float32 mode on the SAME model objects."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import parity_cases as pc
from gennet_b200 import nn

nn.set_compute_dtype('bf16x3')
(g, d, dg, sub_g), (og, od, ocomp, osub), z, sX, sy, ny = pc.burst_case(512, 8)
d.train_on_batch(sX, sy)
od.train_on_batch(sX, sy)
if os.environ.get('RESYNC', '1') == '1':
    pc.resync([(d, od)])
rec = {}
for l in dg.all_layers():
    ob, of = l.backward, l.forward
    def bw(dy, ctx, need_dx=True, _l=l, _ob=ob):
        dx = _ob(dy, ctx, need_dx)
        rec.setdefault(nn.compute_dtype(), {})['bwd:' + _l.name] = None if dx is None else dx.detach().float().cpu().numpy().copy()
        return dx
    def fw(x, ctx, _l=l, _of=of):
        y = _of(x, ctx)
        rec.setdefault(nn.compute_dtype(), {})['fwd:' + _l.name] = y.detach().float().cpu().numpy().copy()
        return y
    l.backward, l.forward = bw, fw
noise = pc.draw_noise(ocomp, z, 0)
pn = pc.map_noise(noise, ocomp, dg)
w0 = [w.copy() for w in g.get_weights()]
for mode in ('bf16x3', 'float32'):
    nn.set_compute_dtype(mode)
    g.set_weights(w0)
    dg.train_on_batch(z, [1] * 8, _noise=pn)
names = [l.name for l in dg.all_layers()]
def rel(a, b):
    if a is None or b is None: return float('nan')
    return np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-30)
A, B = rec['bf16x3'], rec['float32']
for n in names:
    print('fwd %-22s %.2e' % (n, rel(A.get('fwd:' + n), B.get('fwd:' + n))))
for n in reversed(names):
    if 'bwd:' + n in A:
        print('bwd %-22s dx %.2e %s' % (n, rel(A['bwd:' + n], B['bwd:' + n]), None if A['bwd:' + n] is None else A['bwd:' + n].shape))
