"""A few device-resident steps of one bench workload and nothing else (the command that goes under ncu).
usage: python scratch/prof_step.py {pe|gan} [mode] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gennet_b200 import nn
cfg = sys.argv[1] if len(sys.argv) > 1 else 'pe'
mode = sys.argv[2] if len(sys.argv) > 2 else 'bf16x3'
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
nn.set_seed(1)
nn.set_compute_dtype(bench.MODES[mode])
s, t, l = bench.make_inputs(7, dev)
w = {'pe': bench.PEWorkload, 'gan': bench.GANWorkload}[cfg](dev, 0, 1, s, t, l)
for it in range(steps):
    r = w.step(it)
torch.cuda.synchronize()
print('ok', cfg, mode, r.detach().cpu().numpy())
