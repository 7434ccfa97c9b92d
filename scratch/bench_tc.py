import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gennet_b200 import _lib as L_
B = 512
layers = [  # L, Cin, Cout, s  (q tower, then mc tower)
    (2048, 64, 128, 1), (2044, 128, 256, 1), (2040, 256, 512, 2), (1018, 512, 1024, 2),
    (1024, 64, 128, 2), (510, 128, 256, 2), (253, 256, 512, 2)]
def timeit(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
tot = {'fwd': 0, 'dgrad': 0, 'wgrad': 0}; totf = 0
for (L, Cin, Cout, s) in layers:
    k = 5
    Lout = (L - k) // s + 1
    x = (torch.randn(B, L, Cin, device='cuda')).to(torch.bfloat16)
    dy = (torch.randn(B, Lout, Cout, device='cuda')).to(torch.bfloat16)
    wk = (torch.randn(k, Cin, Cout, device='cuda') * 0.05).to(torch.bfloat16)
    wt = wk.permute(0, 2, 1).contiguous()
    bias = torch.zeros(Cout, device='cuda')
    y = torch.empty(B, Lout, Cout, dtype=torch.bfloat16, device='cuda')
    dx = torch.empty(B, L, Cin, dtype=torch.bfloat16, device='cuda')
    dw = torch.empty(k, Cin, Cout, device='cuda'); db = torch.empty(Cout, device='cuda'); db0 = torch.empty(Cin, device='cuda')
    st = L_.stream()
    bf = torch.bfloat16
    flops = 2.0 * B * Lout * k * Cin * Cout
    t1 = timeit(lambda: L_.call('gn_conv1d_fwd_bf16', L_.ptr(x, bf), L_.ptr(wt, bf), L_.ptr(bias), L_.ptr(y, bf), B, L, Cin, Lout, Cout, k, s, 0, 1, 0.0, st))
    t1b = timeit(lambda: L_.call('gn_conv1d_fwd_bf16', L_.ptr(x, bf), L_.ptr(wt, bf), None, L_.ptr(y, bf), B, L, Cin, Lout, Cout, k, s, 0, 0, 0.0, st))
    print('   fwd no-bias/no-act %.3f ms' % t1b)
    t2 = timeit(lambda: L_.call('gn_conv1d_dgrad_bf16', L_.ptr(dy, bf), L_.ptr(wk, bf), L_.ptr(x, bf), L_.ptr(dx, bf), L_.ptr(db0), B, L, Cin, Lout, Cout, k, s, 0, 1, 0.0, st))
    t2b = timeit(lambda: L_.call('gn_conv1d_dgrad_bf16', L_.ptr(dy, bf), L_.ptr(wk, bf), None, L_.ptr(dx, bf), None, B, L, Cin, Lout, Cout, k, s, 0, 0, 0.0, st))
    print('   dgrad no-mask %.3f ms' % t2b)
    t3 = timeit(lambda: L_.call('gn_conv1d_wgrad_bf16', L_.ptr(x, bf), L_.ptr(dy, bf), L_.ptr(dw), L_.ptr(db), B, L, Cin, Lout, Cout, k, s, 0, st))
    print('L=%4d %4d->%4d s%d  GF=%7.1f  fwd %.3f ms %6.1f TF | dgrad %.3f ms %6.1f TF | wgrad %.3f ms %6.1f TF' % (
        L, Cin, Cout, s, flops / 1e9, t1, flops / t1 / 1e9, t2, flops / t2 / 1e9, t3, flops / t3 / 1e9))
    tot['fwd'] += t1; tot['dgrad'] += t2; tot['wgrad'] += t3; totf += flops
print('total ms', tot, 'sum %.2f ms; flops/step %.2f TF -> %.1f TF/s' % (sum(tot.values()), 3 * totf / 1e12, 3 * totf / sum(tot.values()) / 1e9))
