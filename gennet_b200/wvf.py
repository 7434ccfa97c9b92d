"""Builders and loop helpers of train_on_wvf_version/nn.py (BASELINE config 4), same names."""
import glob

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream
from .nn import device
from .nn import (Activation, Conv1D, Dense, Dropout, Flatten, Input, Model, Reshape, Adam, SGD, set_trainability)


def sample_data(n_samples=10000, x_vals=np.arange(0, 5, .1), max_offset=100, mul_range=[1, 2], rng=np.random):
    """nn.py:58-70 (host; a 50-point toy generator, not on the device path)."""
    vectors = []
    for i in range(n_samples):
        offset = rng.random_sample() * max_offset
        mul = mul_range[0] + rng.random_sample() * (mul_range[1] - mul_range[0])
        vectors.append(np.sin(offset + x_vals * mul) / 2 + .5)
    return np.array(vectors)


def get_generative(G_in, dense_dim=300, out_dim=8192, lr=0.425e-1):
    """nn.py:72-81."""
    x = Dense(dense_dim)(G_in)
    x = Activation('relu')(x)
    x = Dense(150)(x)
    x = Activation('relu')(x)
    G_out = Dense(out_dim, activation='tanh')(x)
    G = Model(G_in, G_out)
    G.compile(loss='binary_crossentropy', optimizer=SGD(lr=lr))
    return G, G_out


def get_discriminative(D_in, lr=1e-6, drate=.25, n_channels=25, conv_sz=5, leak=.2):
    """nn.py:83-93."""
    x = Reshape((-1, 1))(D_in)
    x = Conv1D(n_channels, conv_sz, activation='relu')(x)
    x = Dropout(drate)(x)
    x = Flatten()(x)
    x = Dense(n_channels)(x)
    D_out = Dense(2, activation='sigmoid')(x)
    D = Model(D_in, D_out)
    D.compile(loss='binary_crossentropy', optimizer=Adam(lr=lr))
    return D, D_out


def make_gan(GAN_in, G, D):
    """nn.py:100-106."""
    set_trainability(D, False)
    x = G(GAN_in)
    GAN_out = D(x)
    GAN = Model(GAN_in, GAN_out)
    GAN.compile(loss='binary_crossentropy', optimizer=G.optimizer)
    return GAN, GAN_out


def sample_data_and_gen(G, x_train, noise_dim=10, n_samples=10000, rng=np.random):
    """nn.py:108-122."""
    XT = x_train[:n_samples]
    XN_noise = rng.uniform(0, 1, size=[n_samples, noise_dim])
    XN = G.predict(XN_noise)
    X = np.concatenate((XT, XN))
    y = np.zeros((2 * n_samples, 2))
    y[:n_samples, 1] = 1
    y[n_samples:, 0] = 1
    return X, y


def pretrain(G, D, x_train, noise_dim=10, n_samples=10000, batch_size=32):
    """nn.py:124-128."""
    X, y = sample_data_and_gen(G, x_train, n_samples=n_samples, noise_dim=noise_dim)
    set_trainability(D, True)
    D.fit(X, y, epochs=1, batch_size=batch_size)


def sample_noise(G, noise_dim=10, n_samples=10000, rng=np.random):
    """nn.py:130-134."""
    X = rng.uniform(0, 1, size=[n_samples, noise_dim])
    y = np.zeros((n_samples, 2))
    y[:, 1] = 1
    return X, y


def train(GAN, G, D, x_train, epochs=1000, n_samples=10000, noise_dim=10, batch_size=32, verbose=False, v_freq=1):
    """nn.py:136-152."""
    d_loss, g_loss = [], []
    for epoch in range(epochs):
        X, y = sample_data_and_gen(G, x_train, n_samples=n_samples, noise_dim=noise_dim)
        set_trainability(D, True)
        d_loss.append(D.train_on_batch(X, y))
        X, y = sample_noise(G, n_samples=n_samples, noise_dim=noise_dim)
        set_trainability(D, False)
        g_loss.append(GAN.train_on_batch(X, y))
        if verbose and (epoch + 1) % v_freq == 0:
            print("Epoch #{}: Generative Loss: {}, Discriminative Loss: {}".format(epoch + 1, g_loss[-1], d_loss[-1]))
    return d_loss, g_loss


# ----------------------------------------------------------------------------- waveform ingest (load_txtwfs.py)
_RESAMPLE_OPS = {}


def resample_matrix(Nx, num=512):
    """The linear map of ``scipy.signal.resample(x, num)`` (scipy 1.1.0 semantics, see oracle.resample_fft) for real
    series of length Nx, as an (Nx, num) float32 matrix kept in HBM:  R[j, m] = D(m/num - j/Nx) / Nx  with the
    Dirichlet kernel D(t) = sum_{|k| < ceil(N/2)} exp(2 pi i k t), N = min(num, Nx).  Built once per input length in
    float64 on the host."""
    key = (int(Nx), int(num))
    if key not in _RESAMPLE_OPS:
        N = min(num, Nx)
        K = (N - 1) // 2                                  # highest retained harmonic on the negative side
        Kp = (N + 1) // 2 - 1                             # ... on the positive side (equal to K for even N)
        t = (np.arange(num)[None, :] / float(num)) - (np.arange(Nx)[:, None] / float(Nx))
        R = np.ones((Nx, num))
        # sum_{k=1..K} 2 cos(2 pi k t) (+ the unpaired positive harmonic when N is odd... K == Kp then) in closed form
        assert K == Kp or Kp == K + 1
        with np.errstate(divide='ignore', invalid='ignore'):
            s = np.sin(np.pi * t)
            D = np.where(np.abs(s) < 1e-12, 2.0 * K + 1.0, np.sin(np.pi * (2 * K + 1) * t) / s)
        R = D
        if Kp == K + 1:                                   # odd N never happens for num = 512 >= ... kept for generality
            R = R + np.cos(2 * np.pi * Kp * t)            # real part of the unpaired harmonic (scipy takes .real)
        _RESAMPLE_OPS[key] = torch.as_tensor((R / float(Nx)).astype(np.float32)).to(device()).contiguous()
    return _RESAMPLE_OPS[key]


def ingest_waveforms(data, offsets=None, num=512):
    """Batched load_txtwfs.py:47-50 on the device: ``resample(data, num)`` -> ``/= max`` -> ``roll(offset)`` for a
    (B, Nx) array of equally long waveforms; returns a (B, num) float32 CUDA tensor."""
    x = torch.as_tensor(np.ascontiguousarray(np.asarray(data, dtype=np.float32))).to(device()) \
        if not isinstance(data, torch.Tensor) else data.to(device(), torch.float32)
    x = x.reshape(-1, x.shape[-1]).contiguous()
    B, Nx = x.shape
    R = resample_matrix(Nx, num)
    y = torch.empty((B, num), dtype=torch.float32, device=x.device)
    call('gn_dense_fwd_f32', ptr(x), ptr(R), None, ptr(y), B, Nx, num, _lib.ACT_NONE, 0.0, stream())
    out = torch.empty_like(y)
    off = None
    if offsets is not None:
        off = torch.as_tensor(np.asarray(offsets, dtype=np.int32)).to(x.device).contiguous()
    call('gn_maxnorm_roll_f32', ptr(y), ptr(off, torch.int32) if off is not None else None, ptr(out), B, num, stream())
    return out


def load_data(data_path, n_samples, frequencies=None, num=512, rng=np.random):
    """load_txtwfs.py:31-77 without minke: reads ``<data_path>/*.txt`` (one long time series per file), draws the
    position offset ~U(-100,100) per waveform as the reference does, ingests them on the device (grouped by input
    length) and returns ``(data (n, num) float64, data_pars (n, 2) = [num/2 + offset, frequency])``.
    ``frequencies[i]`` stands in for ``mdcset.waveforms[i].frequency`` (the minke XML catalogue)."""
    files = list(glob.iglob('%s/*.txt' % data_path))[:n_samples]
    series, offs = [], []
    for f in files:
        offs.append(int(rng.uniform(-100, 100)))
        series.append(np.loadtxt(f))
    data = np.zeros((len(series), num))
    by_len = {}
    for i, s_ in enumerate(series):
        by_len.setdefault(len(s_), []).append(i)
    for ln, ids in by_len.items():
        out = ingest_waveforms(np.stack([series[i] for i in ids]), [offs[i] for i in ids], num)
        data[ids] = out.cpu().numpy().astype(np.float64)
    freq = [frequencies[i] if frequencies is not None else np.nan for i in range(len(series))]
    pars = np.array([[(num / 2) + o, fr] for o, fr in zip(offs, freq)]) if series else np.zeros((0, 2))
    return data, pars
