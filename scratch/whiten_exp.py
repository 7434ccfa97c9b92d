"""Time gn_whiten_td_f32 (N = 8192) for experimental builds of the library: python scratch/whiten_exp.py name1 name2 ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
sys.path.insert(0, %r)
import numpy as np, torch
from gennet_b200 import synth
fs, T = 2048, 4
N = fs * T
s = synth.Synthesizer(fs, T, synth.analytic_psd(fs, T))
win = s.window.double().cpu().numpy(); wts = s.weights.double().cpu().numpy()
rng = np.random.default_rng(1)
xs = (rng.standard_normal((700, N)) * 1e-21).astype(np.float32)
ref = np.fft.irfft(np.fft.rfft(xs.astype(np.float64) * win, axis=1) * wts, N, axis=1)
got = s.whiten_td(torch.as_tensor(xs).cuda()).cpu().numpy()
err = np.abs(got - ref).max() / np.abs(ref).max()
res = []
for B in (8192, 32768):
    x = torch.randn(B, N, device='cuda') * 1e-21
    for _ in range(3): s.whiten_td(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): s.whiten_td(x)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20
    res.append('B=%%d %%.1f us %%.0f GB/s (%%.3f)' %% (B, t * 1e3, B * 8 * N / t / 1e6, B * 8 * N / t / 1e6 / 6540.8))
    del x
print('err %%.2e | ' %% err + ' | '.join(res))
''' % ROOT
for name in sys.argv[1:]:
    env = dict(os.environ)
    if name != 'main':
        env['GENNET_B200_LIB'] = os.path.join(ROOT, 'scratch', 'exp', 'libgennet_%s.so' % name)
    r = subprocess.run([sys.executable, '-c', CHILD], env=env, capture_output=True, text=True)
    print('%-12s' % name, r.stdout.strip() or r.stderr[-1500:], flush=True)
