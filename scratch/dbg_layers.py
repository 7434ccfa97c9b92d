"""Layer-by-layer comparison of backward outputs between compute dtypes on the burst DG step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import parity_cases as pc
from gennet_b200 import nn

def run(mode):
    nn.set_compute_dtype(mode)
    (g, d, dg, sub_g), (og, od, ocomp, osub), z, sX, sy, ny = pc.burst_case(512, 8)
    rec = {}
    for l in dg.all_layers():
        ob, of = l.backward, l.forward
        def bw(dy, ctx, need_dx=True, _l=l, _ob=ob):
            dx = _ob(dy, ctx, need_dx)
            rec['bwd:' + _l.name] = (None if dx is None else dx.detach().float().cpu().numpy().copy(),
                                     dy.detach().float().cpu().numpy().copy() if torch.is_tensor(dy) else None)
            return dx
        def fw(x, ctx, _l=l, _of=of):
            y = _of(x, ctx)
            rec['fwd:' + _l.name] = y.detach().float().cpu().numpy().copy() if torch.is_tensor(y) else None
            return y
        l.backward, l.forward = bw, fw
    noise = pc.draw_noise(ocomp, z, 0)
    r = dg.train_on_batch(z, [1] * 8, _noise=pc.map_noise(noise, ocomp, dg))
    return rec, [l.name for l in dg.all_layers()], r

ra, names, r1 = run('float32')
rb, _, r2 = run('bf16x3')
print('loss', r1, r2)
def rel(a, b):
    if a is None or b is None: return float('nan')
    return np.abs(a.astype(np.float64) - b).max() / max(np.abs(a).max(), 1e-30)
for n in names:
    fa, fb = ra.get('fwd:' + n), rb.get('fwd:' + n)
    print('fwd %-22s %.2e' % (n, rel(fa, fb)))
for n in reversed(names):
    if 'bwd:' + n in ra:
        (dxa, dya), (dxb, dyb) = ra['bwd:' + n], rb['bwd:' + n]
        print('bwd %-22s dy %.2e  dx %.2e  shape %s' % (n, rel(dya, dyb), rel(dxa, dxb), None if dxa is None else dxa.shape))
