"""iteration times of the other BASELINE configs (1: burst GAN, 4: train_on_wvf GAN, 5: two-model GAN)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gennet_b200 import nn, _lib
from tests import parity_cases as pc
def timeit(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def breakdown(f):
    _lib.PROFILE = []
    f(); torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    tot = {}
    for name, tag, a, b, _args in prof:
        d = tot.setdefault(name, [0.0, 0]); d[0] += a.elapsed_time(b); d[1] += 1
    return ', '.join('%s x%d %.2f' % (k.replace('gn_', ''), v[1], v[0]) for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:6])
for mode in (sys.argv[1:] or ('float32', 'f16x2', 'bfloat16')):
    nn.set_compute_dtype(mode)
    # config 1: burst GAN, n_pix 512, batch 64: D step, residual-moments step, G step
    (g, d, dg, sub_g), _, z, sX, sy, ny = pc.burst_case(512, 64)
    def it1():
        d.train_on_batch(sX, sy); sub_g.train_on_batch(z, ny); dg.train_on_batch(z, [1] * 64)
    t = timeit(it1)
    print('%s config1 burst GAN  B=64 L=512 : %.2f ms/iter -> %.0f samples/s | %s' % (mode, t, 64 / t * 1e3, breakdown(it1)))
    # config 4: train_on_wvf GAN out_dim 8192, batch 256
    (G, D, GAN), _, X, y, z, yz = pc.wvf_case(8192, 256)
    def it4():
        D.train_on_batch(X, y); GAN.train_on_batch(z, yz)
    t = timeit(it4)
    print('%s config4 wvf GAN    B=256 out=8192: %.2f ms/iter -> %.0f samples/s | %s' % (mode, t, 256 / t * 1e3, breakdown(it4)))
    # config 5: two-model GAN out_dim 50 (as shipped), batch 1024
    (G, D, GAN), _, X, y, z, yz = pc.two_model_case(1024)
    def it5():
        D.train_on_batch(X, y); GAN.train_on_batch(z, yz)
    t = timeit(it5)
    print('%s config5 2-model    B=1024 out=50: %.2f ms/iter -> %.0f samples/s | %s' % (mode, t, 1024 / t * 1e3, breakdown(it5)))
nn.set_compute_dtype('float32')
