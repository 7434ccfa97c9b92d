"""Per-layer timing of the split-bf16 tensor-core Conv1D kernels on the CNN point estimator's layers (batch 512)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gennet_b200 import _lib as L_
B = int(os.environ.get('B', 512))
NC = int(os.environ.get('NC', 3))
layers = [  # L, Cin, Cout, s  (q tower, then mc tower)
    (2048, 64, 128, 1), (2044, 128, 256, 1), (2040, 256, 512, 2), (1018, 512, 1024, 2),
    (1024, 64, 128, 2), (510, 128, 256, 2), (253, 256, 512, 2)]
def timeit(f, n=3):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
bf = torch.bfloat16
tot = {'fwd': 0, 'dgrad': 0, 'wgrad': 0, 'split': 0}; totf = 0
for (L, Cin, Cout, s) in layers:
    k = 5
    Lout = (L - k) // s + 1
    x = torch.randn(B, L, Cin, device='cuda')
    dy = torch.randn(B, Lout, Cout, device='cuda')
    w = torch.randn(k, Cin, Cout, device='cuda') * 0.05
    st = L_.stream()
    xs = torch.empty(NC, B, L, Cin, dtype=bf, device='cuda')
    dys = torch.empty(NC, B, Lout, Cout, dtype=bf, device='cuda')
    wk = torch.empty(NC, k, Cin, Cout, dtype=bf, device='cuda'); wt = torch.empty(NC, k, Cout, Cin, dtype=bf, device='cuda')
    L_.call('gn_conv_w_split_bf16', L_.ptr(w), L_.ptr(wk, bf), L_.ptr(wt, bf), k, Cin, Cout, NC, st)
    ts = timeit(lambda: L_.call('gn_split_f32_bf16', L_.ptr(x), L_.ptr(xs, bf), x.numel(), NC, st))
    ts += timeit(lambda: L_.call('gn_split_f32_bf16', L_.ptr(dy), L_.ptr(dys, bf), dy.numel(), NC, st))
    bias = torch.zeros(Cout, device='cuda')
    y = torch.empty(B, Lout, Cout, device='cuda'); ys = torch.empty(NC, B, Lout, Cout, dtype=bf, device='cuda')
    dx = torch.empty(B, L, Cin, device='cuda'); dxs = torch.empty(NC, B, L, Cin, dtype=bf, device='cuda')
    dw = torch.empty(k, Cin, Cout, device='cuda'); db = torch.empty(Cout, device='cuda'); db0 = torch.empty(Cin, device='cuda')
    flops = 2.0 * B * Lout * k * Cin * Cout
    t1 = timeit(lambda: L_.call('gn_conv1d_fwd_bf16x3', L_.ptr(xs, bf), L_.ptr(wt, bf), L_.ptr(bias), L_.ptr(y), L_.ptr(ys, bf), B, L, Cin, Lout, Cout, k, s, 0, 1, 0.0, NC, st))
    t1b = timeit(lambda: L_.call('gn_conv1d_fwd_bf16x3', L_.ptr(xs, bf), L_.ptr(wt, bf), None, L_.ptr(y), None, B, L, Cin, Lout, Cout, k, s, 0, 0, 0.0, NC, st))
    t2 = timeit(lambda: L_.call('gn_conv1d_dgrad_bf16x3', L_.ptr(dys, bf), L_.ptr(wk, bf), L_.ptr(x), L_.ptr(dx), L_.ptr(dxs, bf), L_.ptr(db0), B, L, Cin, Lout, Cout, k, s, 0, 1, 0.0, NC, st))
    t2b = timeit(lambda: L_.call('gn_conv1d_dgrad_bf16x3', L_.ptr(dys, bf), L_.ptr(wk, bf), None, L_.ptr(dx), None, None, B, L, Cin, Lout, Cout, k, s, 0, 0, 0.0, NC, st))
    t3 = timeit(lambda: L_.call('gn_conv1d_wgrad_bf16x3', L_.ptr(xs, bf), L_.ptr(dys, bf), L_.ptr(dy), L_.ptr(dw), L_.ptr(db), B, L, Cin, Lout, Cout, k, s, 0, NC, st))
    print('L=%4d %4d->%4d s%d GF=%7.1f split %.3f | fwd %.3f (plain %.3f) ms %6.1f TF | dgrad %.3f (plain %.3f) ms %6.1f TF | wgrad %.3f ms %6.1f TF' % (
        L, Cin, Cout, s, flops / 1e9, ts, t1, t1b, flops / t1 / 1e9, t2, t2b, flops / t2 / 1e9, t3, flops / t3 / 1e9), flush=True)
    tot['fwd'] += t1; tot['dgrad'] += t2; tot['wgrad'] += t3; tot['split'] += ts; totf += flops
mm = tot['fwd'] + tot['dgrad'] + tot['wgrad']
print('NC=%d total ms' % NC, tot, 'mma sum %.2f ms; flops/step %.2f TF -> %.1f TF/s (effective fp32 flops)' % (mm, 3 * totf / 1e12, 3 * totf / mm / 1e9))
