"""GAN iteration (BASELINE config 3: bbhMahoGANy.py generator + discriminator, n_pix 2048, batch 128/GPU): time + breakdown"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gennet_b200 import nn, bbh, _lib
mode = sys.argv[1] if len(sys.argv) > 1 else 'bfloat16'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
torch.cuda.set_device(0)
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
for _ in range(3):
    bbh.gan_train_step(G, D, DG, ns, real, z1, rn, z2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    r = bbh.gan_train_step(G, D, DG, ns, real, z1, rn, z2)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print('%s B=%d: %.3f ms / GAN iteration -> %.0f samples/s ; losses %s' % (mode, B, ms, B / ms * 1e3, r))
_lib.PROFILE = []
bbh.gan_train_step(G, D, DG, ns, real, z1, rn, z2)
torch.cuda.synchronize()
prof, _lib.PROFILE = _lib.PROFILE, None
tot = {}
for name, tag, a, b, _args in prof:
    d = tot.setdefault(name, [0.0, 0]); d[0] += a.elapsed_time(b); d[1] += 1
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:25]:
    print('  %-34s n=%3d %8.3f ms' % (k, v[1], v[0]))
print('  sum %.3f ms over %d launches' % (sum(v[0] for v in tot.values()), sum(v[1] for v in tot.values())))
