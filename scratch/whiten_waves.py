import sys, os
sys.path.insert(0, '/root/repo')
import torch
from gennet_b200 import synth
fs, T = 2048, 4
N = fs * T
s = synth.Synthesizer(fs, T, synth.analytic_psd(fs, T))
for B in (8192, 32768):
    x = torch.randn(B, N, device='cuda') * 1e-21
    for _ in range(3): s.whiten_td(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): s.whiten_td(x)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20
    print(os.environ.get('GENNET_B200_LIB', 'main')[-10:], 'B=%6d  %.1f us  %.0f GB/s' % (B, t * 1e3, B * 8 * N / t / 1e6), flush=True)
    del x
