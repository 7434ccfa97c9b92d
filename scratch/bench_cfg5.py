"""BASELINE config 5: the 2_model_version subtract stage at out_dim 16384 (4 s @ 4096 Hz): one iteration of
subtract_model.train (D step on noise + residual rows, G step through the frozen D) per compute mode."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gennet_b200 import nn, _lib
from gennet_b200.two_model import subtract_model as sm
OUT = int(os.environ.get('OUT', 16384)); B = int(os.environ.get('B', 128))
for mode in sys.argv[1:] or ['bf16x3', 'bfloat16', 'float32']:
    nn.clear_session(); nn.set_seed(1); nn.set_compute_dtype(mode)
    sm.hyperparams.noise_dim, sm.hyperparams.outdim = 10, OUT
    G, _ = sm.get_generative(nn.Input(shape=(1, 10)), out_dim=OUT, lr=1e-4)
    D, _ = sm.get_discriminative(nn.Input(shape=(OUT,)), lr=1e-4)
    GAN, _ = sm.make_gan(nn.Input((1, 10)), G, D)
    g = torch.Generator(device='cuda').manual_seed(0)
    X = torch.randn(2 * B, OUT, device='cuda', generator=g) * 5
    y = torch.ones(2 * B, 2, device='cuda')
    z = torch.randn(B, 1, 10, device='cuda', generator=g)
    yz = torch.ones(B, 2, device='cuda')
    def it():
        nn.set_trainability(D, True)
        a = D.train_on_batch(X, y, _return_device=True)
        nn.set_trainability(D, False)
        b = GAN.train_on_batch(z, yz, _return_device=True)
        return a, b
    for _ in range(3): it()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n): r = it()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print('%s out_dim %d B=%d: %.3f ms / iteration -> %.0f samples/s; G params %d; losses %s' % (
        mode, OUT, B, ms, B / ms * 1e3, G.count_params(), [float(v) for v in torch.cat(r).cpu()]), flush=True)
    _lib.PROFILE = []
    it(); torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    tot = {}
    for name, tag, a, b, _args in prof:
        d = tot.setdefault(name, [0.0, 0]); d[0] += a.elapsed_time(b); d[1] += 1
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:8]:
        print('    %-30s n=%3d %8.3f ms' % (k, v[1], v[0]))
    del G, D, GAN
    torch.cuda.empty_cache()
