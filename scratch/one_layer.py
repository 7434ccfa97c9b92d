"""one launch each of the tcgen05 Conv1D fwd / dgrad / wgrad kernels on the largest PE layer (for ncu)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gennet_b200 import _lib as L_
B, L, Cin, Cout, s, k = 512, 1018, 512, 1024, 2, 5
if len(sys.argv) > 1:
    L, Cin, Cout, s = [int(v) for v in sys.argv[1:5]]
Lout = (L - k) // s + 1
bf = torch.bfloat16
x = torch.randn(B, L, Cin, device='cuda').to(bf)
dy = torch.randn(B, Lout, Cout, device='cuda').to(bf)
wk = (torch.randn(k, Cin, Cout, device='cuda') * 0.05).to(bf)
wt = wk.permute(0, 2, 1).contiguous()
bias = torch.zeros(Cout, device='cuda')
y = torch.empty(B, Lout, Cout, dtype=bf, device='cuda')
dx = torch.empty(B, L, Cin, dtype=bf, device='cuda')
dw = torch.empty(k, Cin, Cout, device='cuda'); db = torch.empty(Cout, device='cuda'); db0 = torch.empty(Cin, device='cuda')
st = L_.stream()
for _ in range(2):
    L_.call('gn_conv1d_fwd_bf16', L_.ptr(x, bf), L_.ptr(wt, bf), L_.ptr(bias), L_.ptr(y, bf), B, L, Cin, Lout, Cout, k, s, 0, 1, 0.0, st)
    L_.call('gn_conv1d_dgrad_bf16', L_.ptr(dy, bf), L_.ptr(wk, bf), L_.ptr(x, bf), L_.ptr(dx, bf), L_.ptr(db0), B, L, Cin, Lout, Cout, k, s, 0, 1, 0.0, st)
    L_.call('gn_conv1d_wgrad_bf16', L_.ptr(x, bf), L_.ptr(dy, bf), L_.ptr(dw), L_.ptr(db), B, L, Cin, Lout, Cout, k, s, 0, st)
torch.cuda.synchronize()
print('ok')
