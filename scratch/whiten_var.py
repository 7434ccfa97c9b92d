"""Timings and float64 error of the synthesis kernel organisations (GN_SYNTH_VAR = 0, 1; GN_WHITEN_RADIX = 64) on one GPU.
Each variant runs in its own process (the selection is read once per process)."""
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
xs = (rng.standard_normal((6, N)) * 1e-21).astype(np.float32)
ref = np.fft.irfft(np.fft.rfft(xs.astype(np.float64) * win, axis=1) * wts, N, axis=1)
got = s.whiten_td(torch.as_tensor(xs).cuda()).cpu().numpy()
err = np.abs(got - ref).max() / np.abs(ref).max()
B = 8192
x = torch.randn(B, N, device='cuda') * 1e-21
templ = torch.randn(1024, N, device='cuda') * 1e-22
idx = torch.randint(0, 1024, (B,), device='cuda', dtype=torch.int32)
out = torch.empty(B, fs, device='cuda')
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t1 = timeit(lambda: s.whiten_td(x))
t2 = timeit(lambda: s.whiten_td(x, crop=True))
t3 = timeit(lambda: s.synth(B, templates=templ, tidx=idx, out=out))
xs512 = x[:512]
t4 = timeit(lambda: s.whiten_td(xs512, crop=True), 50)
print('VAR=%%s err=%%.2e | whiten full %%.1f us %%.0f GB/s | crop %%.1f us %%.0f GB/s | philox synth %%.1f us %%.0f GB/s | B=512 crop %%.1f us' %% (
    os.environ.get('GN_SYNTH_VAR'), err, t1 * 1e3, B * 8 * N / t1 / 1e6, t2 * 1e3, B * (4 * N + 4 * fs) / t2 / 1e6,
    t3 * 1e3, B * (4 * N + 4 * fs) / t3 / 1e6, t4 * 1e3))
''' % ROOT
# arguments: comma-separated KEY=VALUE settings per run, e.g.  GN_SYNTH_VAR=2,GN_SYNTH_WAVES=1  GN_WHITEN_RADIX=64
for spec in (sys.argv[1:] or ['GN_SYNTH_VAR=0', 'GN_SYNTH_VAR=1', 'GN_WHITEN_RADIX=64']):
    env = dict(os.environ)
    env.update(kv.split('=') for kv in spec.split(','))
    r = subprocess.run([sys.executable, '-c', CHILD], env=env, capture_output=True, text=True)
    print(spec, '|', r.stdout.strip() or r.stderr[-2000:], flush=True)
