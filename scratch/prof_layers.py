"""Per-launch device time of every GEMM-shaped entry point in one step: geometry, ms, algorithmic TF/s.
usage: python scratch/prof_layers.py {pe|gan} [mode]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gennet_b200 import nn, _lib
cfg = sys.argv[1] if len(sys.argv) > 1 else 'gan'
mode = sys.argv[2] if len(sys.argv) > 2 else 'f16x2'
torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
nn.set_seed(1)
nn.set_compute_dtype(bench.MODES[mode])
s, t, l = bench.make_inputs(7, dev)
w = {'pe': bench.PEWorkload, 'gan': bench.GANWorkload}[cfg](dev, 0, 1, s, t, l)
for it in range(4):
    w.step(it)
torch.cuda.synchronize()
_lib.PROFILE = []
w.step(100)
torch.cuda.synchronize()
prof, _lib.PROFILE = _lib.PROFILE, None
tot = 0.0
for name, tag, a, b, args in prof:
    ms = a.elapsed_time(b)
    tot += ms
    fl = bench.call_flops(name, args)
    if name in bench._CONV_ARGPOS:
        p = bench._CONV_ARGPOS[name]
        geom = 'B%d L%d Cin%d Lout%d Cout%d k%d s%d' % tuple(int(v) for v in args[p:p + 7])
    elif name in bench._DENSE_ARGPOS:
        p = bench._DENSE_ARGPOS[name]
        geom = 'M%d K%d N%d' % tuple(int(v) for v in args[p:p + 3])
    else:
        geom = ''
    if ms > 0.05:
        print('%-30s %-46s %8.3f ms %8.1f TF/s' % (name, geom, ms, fl / ms / 1e9 if fl else 0.0))
print('total %.2f ms' % tot)
