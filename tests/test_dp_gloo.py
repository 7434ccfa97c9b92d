"""Data-parallel host logic on CPU: 2 gloo ranks, each with half of the global batch, must reproduce the
single-process global-batch step (gradients all-reduced, SyncBN statistics, global loss scaling)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


class _MP:
    def setattr(self, obj, name, val):
        setattr(obj, name, val)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(kind, n_pix, seed, mode='float32'):
    from gennet_b200 import nn, bbh
    nn.clear_session()
    nn.set_seed(seed)
    nn.set_compute_dtype(mode)
    bbh.n_pix = n_pix
    if kind == 'pe':
        m = bbh.signal_pe_model()
        m.compile(loss='mean_squared_error', optimizer=nn.Adam(lr=9e-5, beta_1=0.5), metrics=['accuracy'])
    else:
        m = bbh.generator_model()           # Dense + BatchNorm + Dropout + fused upsampling convs
        m.compile(loss='mean_squared_error', optimizer=nn.Adam(lr=9e-5, beta_1=0.5), metrics=['accuracy'])
    return m


def _data(kind, n_pix, B):
    rs = np.random.RandomState(0)
    if kind == 'pe':
        return rs.normal(size=(B, n_pix, 1)).astype(np.float32), [rs.uniform(0.2, 1, B).astype(np.float32),
                                                                   rs.uniform(0.5, 1, B).astype(np.float32)]
    return rs.uniform(-1, 1, (B, 100)).astype(np.float32), rs.normal(size=(B, n_pix, 1)).astype(np.float32)


def _noise_for(model, B, lo, hi):
    """Identical dropout masks in the single-process and sharded runs: drawn for the global batch, sliced."""
    from tests.parity_cases import noise_layers
    rs = np.random.RandomState(5)
    out = {}
    for l in noise_layers(model):
        full = (rs.uniform(size=(B,) + tuple(l.output_shape)) >= l.rate).astype(np.float32)
        out[l.name] = full[lo:hi]
    return out


def _worker(rank, world, port, kind, n_pix, B, q, mode='float32'):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from tests import fake_backend
    fake_backend.install(_MP())
    from gennet_b200 import parallel
    dp = parallel.init_data_parallel('gloo')
    m = _build(kind, n_pix, 3, mode)
    parallel.broadcast_weights(m)
    x, y = _data(kind, n_pix, B)
    lo, hi = dp.shard(B)
    ys = [t[lo:hi] for t in y] if isinstance(y, list) else y[lo:hi]
    r = m.train_on_batch(x[lo:hi], ys, _noise=_noise_for(m, B, lo, hi))
    if rank == 0:
        q.put((r, m.get_gradients(), m.get_weights()))
    parallel.shutdown()


@pytest.mark.parametrize('kind,mode', [('pe', 'float32'), ('gen', 'float32'), ('gen', 'f16x2')])
def test_two_rank_step_equals_single_process(kind, mode, monkeypatch):
    """mode 'f16x2': the BatchNormalization statistics come out of the convolution that feeds the layer and are
    all-reduced like the ones of the separate pass; every rank scales its operand planes by its own shard's maximum."""
    n_pix, B = 64, 8
    from tests import fake_backend
    from gennet_b200 import nn
    fake_backend.install(monkeypatch)
    monkeypatch.setitem(nn._STATE, 'dtype', nn._STATE['dtype'])      # restored after the test
    m = _build(kind, n_pix, 3, mode)
    x, y = _data(kind, n_pix, B)
    r1 = m.train_on_batch(x, y, _noise=_noise_for(m, B, 0, B))
    g1, w1 = m.get_gradients(), m.get_weights()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, n_pix, B, q, mode)) for r in range(2)]
    for p in procs:
        p.start()
    r2, g2, w2 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.allclose(r1, r2, rtol=1e-5, atol=1e-7)
    gmax = max(np.abs(a).max() for a in g1)
    for a, b in zip(g1, g2):
        assert np.abs(a - b).max() <= 2e-5 * max(np.abs(a).max(), 1e-2 * gmax), (a.shape, np.abs(a - b).max())
    names = [p.name for p in m.params]
    for n, a, b in zip(names, w1, w2):
        if 'moving_' in n:        # SyncBN: moving statistics come from the GLOBAL batch on every rank
            assert np.abs(a - b).max() <= 1e-5 * max(np.abs(a).max(), 1e-6) + 1e-7, n
        else:                     # Adam turns float-rounding differences on ~zero gradients into +-lr steps
            assert np.abs(a - b).max() <= 2 * 9e-5, n
