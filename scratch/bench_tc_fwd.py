"""forward conv timing of the four q-tower layers under GN_TC_DEBUG experiments (results are INVALID under debug flags)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gennet_b200 import _lib as L_
B = 512
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = []
for (L, Cin, Cout, s) in [(2048, 64, 128, 1), (2044, 128, 256, 1), (2040, 256, 512, 2), (1018, 512, 1024, 2)]:
    k = 5; Lout = (L - k) // s + 1; bf = torch.bfloat16
    x = torch.randn(B, L, Cin, device='cuda').to(bf)
    wt = (torch.randn(k, Cout, Cin, device='cuda') * 0.05).to(bf)
    bias = torch.zeros(Cout, device='cuda'); y = torch.empty(B, Lout, Cout, dtype=bf, device='cuda')
    st = L_.stream()
    t = timeit(lambda: L_.call('gn_conv1d_fwd_bf16', L_.ptr(x, bf), L_.ptr(wt, bf), L_.ptr(bias), L_.ptr(y, bf), B, L, Cin, Lout, Cout, k, s, 0, 1, 0.0, st))
    out.append('%d->%d: %.3f' % (Cin, Cout, t))
print('GN_TC_DEBUG=%s fwd ms: ' % os.environ.get('GN_TC_DEBUG', '0') + ' | '.join(out))
