"""per-call list (name, integer args, ms) of one GAN iteration"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gennet_b200 import nn, bbh, _lib
mode = sys.argv[1] if len(sys.argv) > 1 else 'bfloat16'
B = 128
nn.set_seed(1); nn.set_compute_dtype(mode); bbh.n_pix = 2048
L = 2048
noise_signal = np.random.RandomState(0).normal(size=(L, 1)).astype(np.float32)
G, D, DG, _ = bbh.build_gan(noise_signal)
ns = torch.as_tensor(noise_signal.reshape(-1)).cuda()
g = torch.Generator(device='cuda').manual_seed(0)
real = torch.randn(B, L, device='cuda', generator=g)
z1 = torch.rand(B, 100, device='cuda', generator=g) * 2 - 1
z2 = torch.rand(B, 100, device='cuda', generator=g) * 2 - 1
rn = torch.randn(B, L, device='cuda', generator=g)
for _ in range(2):
    bbh.gan_train_step(G, D, DG, ns, real, z1, rn, z2)
torch.cuda.synchronize()
rec = []
orig = _lib.call
def call(name, *args, tag=None):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); orig(name, *args); e1.record()
    rec.append((name, [a for a in args if isinstance(a, int) and not isinstance(a, bool) and abs(a) < 10**7], e0, e1))
for mod in (nn, bbh):
    mod.call = call
bbh.gan_train_step(G, D, DG, ns, real, z1, rn, z2)
torch.cuda.synchronize()
for name, ints, a, b in rec:
    t = a.elapsed_time(b)
    if t > 0.15:
        print('%-30s %8.3f ms  %s' % (name, t, ints))
