import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gennet_b200 import synth
fs, T = 2048, 4
N = fs * T
psd = synth.analytic_psd(fs, T)
s = synth.Synthesizer(fs, T, psd)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
x = torch.randn(B, N, device='cuda') * 1e-21
templ = torch.randn(1024, N, device='cuda') * 1e-22
idx = torch.randint(0, 1024, (B,), device='cuda', dtype=torch.int32)
normals = torch.randn(B, 2, N // 2 + 1, device='cuda')
out = torch.empty(B, fs, device='cuda')
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t = timeit(lambda: s.whiten_td(x))
print('whiten_td full  B=%d: %.3f ms  %.1f GB/s (8N B/series)' % (B, t, B * 8 * N / t / 1e6))
t = timeit(lambda: s.whiten_td(x, crop=True))
print('whiten_td crop  B=%d: %.3f ms  %.1f GB/s (4N+4L)' % (B, t, B * (4 * N + 4 * fs) / t / 1e6))
t = timeit(lambda: s.synth(B, templates=templ, tidx=idx, normals=normals, out=out))
print('synth fed       B=%d: %.3f ms  %.1f GB/s (8N+4L)' % (B, t, B * (8 * N + 4 * fs + 8) / t / 1e6))
t = timeit(lambda: s.synth(B, templates=templ, tidx=idx, out=out))
print('synth philox    B=%d: %.3f ms  %.1f GB/s (4N+4L)' % (B, t, B * (4 * N + 4 * fs) / t / 1e6))
