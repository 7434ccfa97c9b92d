"""Per-gradient errors of whole-step parity cases in a given compute dtype (no asserts)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import parity_cases as pc
from gennet_b200 import nn

def report(tag, prod, orc, x, y, seed=0):
    noise = pc.draw_noise(orc, x, seed)
    pnoise = pc.map_noise(noise, orc, prod)
    rec = pc.record_kinks(prod)
    rp = prod.train_on_batch(x, y, _noise=pnoise)
    on = dict(noise); on['__kinks__'] = pc.map_kinks(rec, orc, prod)
    t = time.time()
    try:
        ro = orc.train_on_batch(x, y, noise=on)
    except AssertionError as e:
        print(tag, 'ORACLE ASSERT', e); return
    gp = prod.get_gradients()
    gf = 1e-3 * max(np.abs(b).max() for b in orc.last_grads)
    errs = [np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), gf) for a, b in zip(gp, orc.last_grads)]
    print(tag, 'loss', rp[0], ro[0], 'oracle %.1fs' % (time.time() - t))
    print('   ', ' '.join('%s:%.1e' % (tuple(b.shape), e) for e, b in zip(errs, orc.last_grads)), flush=True)

which = sys.argv[1]
for mode in sys.argv[2:]:
    nn.set_compute_dtype(mode)
    if which == 'burst':
        (g, d, dg, sub_g), (og, od, ocomp, osub), z, sX, sy, ny = pc.burst_case(int(os.environ.get('NPIX', 512)), 8)
        report(mode + ' burst D', d, od, sX, sy)
        report(mode + ' burst subG', sub_g, osub, z, ny)
        pc.resync([(g, og)])
        report(mode + ' burst DG', dg, ocomp, z, [1] * 8)
    elif which == 'gan':
        (g, d, dg), (og, od, ocomp), z, sX, sy = pc.gan_case(int(os.environ.get('NPIX', 2048)), 8)
        a, b = g.predict(z), og.predict(z)
        print(mode, 'G.predict err', np.abs(a - b).max() / np.abs(b).max())
        report(mode + ' gan D', d, od, sX, sy)
        report(mode + ' gan DG', dg, ocomp, z, [1] * 8)
    elif which == 'pe':
        prod, orc, x, y = pc.pe_case(int(os.environ.get('NPIX', 2048)), 8)
        report(mode + ' pe', prod, orc, x, y)
if which == 'burst_flow':
    for mode in sys.argv[2:]:
        nn.set_compute_dtype(mode)
        (g, d, dg, sub_g), (og, od, ocomp, osub), z, sX, sy, ny = pc.burst_case(int(os.environ.get('NPIX', 512)), 8)
        variant = os.environ.get('VARIANT', 'full')
        if variant in ('full', 'predict'):
            po, oo = d.predict(sX), od.predict(sX)
            print('D.predict err', np.abs(po - oo).max() / np.abs(oo).max())
        report(mode + ' D', d, od, sX, sy)
        if variant in ('full', 'resync'):
            pc.resync([(d, od)])
        report(mode + ' subG', sub_g, osub, z, ny)
        pc.resync([(g, og)])
        report(mode + ' DG', dg, ocomp, z, [1] * 8)
