import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gennet_b200 import synth
fs, T = 2048, 4
s = synth.Synthesizer(fs, T, synth.analytic_psd(fs, T))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
x = torch.randn(B, fs * T, device='cuda') * 1e-21
for _ in range(3):
    y = s.whiten_td(x)
torch.cuda.synchronize()
print('ok')
