"""gn_whiten_td_f32 throughput against the batch size (tail effect of the persistent grid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gennet_b200 import synth
fs, T = 2048, 4
N = fs * T
s = synth.Synthesizer(fs, T, synth.analytic_psd(fs, T))
for B in (4096, 8192, 8288, 11840, 16384, 16576):
    x = torch.randn(B, N, device='cuda') * 1e-21
    for _ in range(3): s.whiten_td(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): s.whiten_td(x)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20
    print('B=%6d  %.1f us  %.0f GB/s' % (B, t * 1e3, B * 8 * N / t / 1e6), flush=True)
    del x
