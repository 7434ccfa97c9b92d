"""2-GPU data-parallel GAN step in bf16 mode vs the same global batch on one GPU (rank 0 runs both)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gennet_b200 import nn, bbh, parallel
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0'))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
dp = parallel.init_data_parallel('nccl') if world > 1 else None
nn.set_seed(1); nn.set_compute_dtype('bfloat16'); bbh.n_pix = 256
L, B = 256, 16
rs = np.random.RandomState(0)
noise_signal = rs.normal(size=(L, 1)).astype(np.float32)
G, D, DG, _ = bbh.build_gan(noise_signal)
if dp is not None:
    for m in (G, D):
        parallel.broadcast_weights(m)
z = rs.uniform(-1, 1, (B, 100)).astype(np.float32)
sX = rs.normal(size=(2 * B, L, 2, 1)).astype(np.float32)
sy = np.array([1.0] * B + [0.0] * B, dtype=np.float32)
# feed dropout masks so the comparison does not depend on the Philox offsets of the ranks
def masks(model, x, seed):
    r = np.random.RandomState(seed); out = {}
    shp = None
    for l in model.all_layers():
        if type(l).__name__ == 'Dropout':
            out[l.name] = None
    return out
if world > 1:
    sl = slice(rank * B // world, (rank + 1) * B // world)
    sl2 = np.r_[np.arange(B)[sl], B + np.arange(B)[sl]]
    for l in G.all_layers() + D.all_layers():
        if type(l).__name__ == 'Dropout':
            l.rate = 0.0
    rd = D.train_on_batch(sX[sl2], sy[sl2])
    rg = DG.train_on_batch(z[sl], [1] * (B // world))
    if rank == 0:
        np.save('/tmp/dp_w.npy', np.concatenate([w.ravel() for w in G.get_weights() + D.get_weights()]))
        print('dp losses', rd, rg)
    torch.distributed.barrier(); parallel.shutdown()
else:
    for l in G.all_layers() + D.all_layers():
        if type(l).__name__ == 'Dropout':
            l.rate = 0.0
    rd = D.train_on_batch(sX, sy)
    rg = DG.train_on_batch(z, [1] * B)
    w = np.concatenate([w.ravel() for w in G.get_weights() + D.get_weights()])
    print('single losses', rd, rg)
    if os.path.exists('/tmp/dp_w.npy'):
        wd = np.load('/tmp/dp_w.npy')
        print('weights: max abs diff %.3e (scale %.3e), rel L2 %.3e' % (np.abs(w - wd).max(), np.abs(w).max(), np.linalg.norm(w - wd) / np.linalg.norm(w)))
