"""per-launch device time of one PE training step (bf16 path), in call order"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from gennet_b200 import nn, bbh, _lib
torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
nn.set_seed(1); nn.set_compute_dtype('bfloat16'); bbh.n_pix = bench.FS
pe = bbh.signal_pe_model()
pe.compile(loss='mean_squared_error', optimizer=nn.Adam(lr=9e-5, beta_1=0.5), metrics=['accuracy'])
B, L = bench.BATCH, bench.FS
x = torch.randn(B, L, 1, device=dev); y = [torch.rand(B, device=dev), torch.rand(B, device=dev)]
for _ in range(3): pe.train_on_batch(x, y, _return_device=True)
torch.cuda.synchronize()
_lib.PROFILE = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); pe.train_on_batch(x, y, _return_device=True); e1.record()
torch.cuda.synchronize()
prof, _lib.PROFILE = _lib.PROFILE, None
tot = 0
for name, tag, a, b, _args in prof:
    t = a.elapsed_time(b); tot += t
    print('%-34s %8.3f ms  %s' % (name, t, tag or ''))
print('sum of launches %.3f ms; step wall (events) %.3f ms; %d launches' % (tot, e0.elapsed_time(e1), len(prof)))
