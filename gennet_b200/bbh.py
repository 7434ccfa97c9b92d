"""Model builders and training steps of BBH_version/bbhMahoGANy.py, same names and layer lists.

The module-level globals mirror the reference's config block (bbhMahoGANy.py:84-113); builders read
``n_pix`` at call time exactly as the reference's do, so ``bbh.n_pix = 2048`` selects BASELINE configs 2/3.
"""
import numpy as np
import torch

from . import nn
from .nn import (Activation, BatchNormalization, Conv1D, Conv2D, Dense, Dropout, Flatten, Input, LeakyReLU, Model,
                 ReLU, Reshape, Sequential, UpSampling1D, Adam, set_trainable, chisquare_Loss)
from ._lib import call, ptr, stream

# bbhMahoGANy.py:84-113
n_pix = 1024
n_sig = 1.0
batch_size = 8
pe_batch_size = 8
lr = 9e-5
chi_loss = False
comb_pe_model = False
n_noise_real = 1
cnn_noise_frac = 1.0 / 8.0


class MyLayer(nn.StackResidual):
    """bbhMahoGANy.py:164-188."""


def data_subtraction_model(noise_signal, npix):
    """bbhMahoGANy.py:190-210."""
    model = Sequential()
    model.add(MyLayer(noise_signal, input_shape=(npix, 1)))
    return model


def generator_model():
    """bbhMahoGANy.py:212-295."""
    model = Sequential()
    act, momentum, drate, padding, weights, filtsize = 'tanh', 0.99, 0.2, 'same', 'glorot_uniform', 5
    model.add(Dense(256 * 1 * int(n_pix / 2), kernel_initializer=weights, input_shape=(100,)))
    model.add(BatchNormalization(momentum=momentum))
    model.add(Activation(act))
    model.add(Dropout(drate))
    model.add(Reshape((int(n_pix / 2), 256)))
    for filters, strides, up in ((64, 2, True), (128, 1, True), (256, 1, False), (512, 1, False), (1024, 1, False)):
        if up:
            model.add(UpSampling1D(size=2))
        model.add(Conv1D(filters, filtsize, kernel_initializer=weights, strides=strides, padding=padding))
        model.add(BatchNormalization(momentum=momentum))
        model.add(Activation(act))
        model.add(Dropout(drate))
    model.add(Conv1D(1, filtsize, padding=padding))
    model.add(Activation('linear'))
    return model


def signal_pe_model():
    """bbhMahoGANy.py:297-406, the shipped ``comb_pe_model = False`` branch (:356-404).  The
    ``comb_pe_model`` branch references an undefined ``batchnorm`` (:318) and cannot run in the reference."""
    if comb_pe_model:
        raise NotImplementedError('comb_pe_model=True is dead code in the reference (NameError at bbhMahoGANy.py:318)')
    inputs = Input(shape=(n_pix, 1))
    act = 'relu'
    mc_branch = Conv1D(64, 5, strides=2, padding='same')(inputs)
    mc_branch = Activation(act)(mc_branch)
    for f in (128, 256, 512):
        mc_branch = Conv1D(f, 5, strides=2)(mc_branch)
        mc_branch = Activation(act)(mc_branch)
    mc_branch = Flatten()(mc_branch)
    mc_branch = Dense(1)(mc_branch)
    mc_branch = Activation('relu')(mc_branch)

    q_branch = Conv1D(64, 5, strides=1, padding='same')(inputs)
    q_branch = Activation(act)(q_branch)
    for f, s in ((128, 1), (256, 1), (512, 2), (1024, 2)):
        q_branch = Conv1D(f, 5, strides=s)(q_branch)
        q_branch = Activation(act)(q_branch)
    q_branch = Flatten()(q_branch)
    q_branch = Dense(1)(q_branch)
    q_branch = ReLU(max_value=1.0)(q_branch)
    return Model(inputs=inputs, outputs=[mc_branch, q_branch], name='pe net')


def signal_discriminator_model():
    """bbhMahoGANy.py:408-498 (num_lays = 2, no batchnorm, no maxpool)."""
    weights, drate, alpha, padding, filtsize, n_neuron_scale = 'glorot_uniform', 0.4, 0.2, 'same', (5, 5), 4
    model = Sequential()
    model.add(Conv2D(64 * n_neuron_scale, filtsize, kernel_initializer=weights, input_shape=(n_pix, 2, 1),
                     strides=(2, 1), padding=padding))
    model.add(LeakyReLU(alpha=alpha))
    model.add(Dropout(drate))
    model.add(Conv2D(128 * n_neuron_scale, filtsize, kernel_initializer=weights, strides=(2, 1), padding=padding))
    model.add(LeakyReLU(alpha=alpha))
    model.add(Dropout(drate))
    model.add(Flatten())
    model.add(Dense(1))
    model.add(Activation('sigmoid'))
    return model


def generator_after_subtracting_noise(generator, data_subtraction):
    """bbhMahoGANy.py:500-519."""
    model = Sequential()
    model.add(generator)
    model.add(data_subtraction)
    return model


def generator_containing_signal_discriminator(generator, signal_discriminator):
    """bbhMahoGANy.py:521-539."""
    model = Sequential()
    model.add(generator)
    model.add(signal_discriminator)
    return model


def build_gan(noise_signal):
    """Model set-up of main(), bbhMahoGANy.py:1088-1119: returns the compiled
    (generator, signal_discriminator, signal_discriminator_on_generator, data_subtraction_on_generator)."""
    generator = generator_model()
    signal_discriminator = signal_discriminator_model()
    data_subtraction = data_subtraction_model(noise_signal, n_pix)
    data_subtraction_on_generator = generator_after_subtracting_noise(generator, data_subtraction)
    data_subtraction_on_generator.compile(loss='binary_crossentropy', optimizer=Adam(lr=lr, beta_1=0.5),
                                          metrics=['accuracy'])
    signal_discriminator_on_generator = generator_containing_signal_discriminator(data_subtraction_on_generator,
                                                                                  signal_discriminator)
    set_trainable(signal_discriminator, False)
    loss = chisquare_Loss(n_sig) if chi_loss else 'binary_crossentropy'
    signal_discriminator_on_generator.compile(loss=loss, optimizer=Adam(lr=lr, beta_1=0.5), metrics=['accuracy'])
    set_trainable(signal_discriminator, True)
    signal_discriminator.compile(loss='binary_crossentropy', optimizer=Adam(lr=lr, beta_1=0.5), metrics=['accuracy'])
    return generator, signal_discriminator, signal_discriminator_on_generator, data_subtraction_on_generator


def pe_train_step(signal_pe, templates, pars, idx, noise, sigma):
    """One iteration of the CNN loop, bbhMahoGANy.py:1153-1166, with the batch assembled on the device:
    templates (n, n_pix) CUDA f32, pars (n, 2) CUDA f32, idx (B,) CUDA int32, noise (B//8, n_pix) CUDA f32
    standard normals, sigma = the np.random.uniform(0,5) draw of :1161."""
    B = idx.shape[0]
    L = templates.shape[1]
    batch = torch.empty((B, L), dtype=torch.float32, device=templates.device)
    call('gn_gather_rows_f32', ptr(templates), ptr(idx, torch.int32), ptr(batch), B, L, stream())
    nn_rows = int(B * cnn_noise_frac)
    if nn_rows > 0:
        call('gn_add_scaled_f32', ptr(batch), ptr(noise), float(sigma), nn_rows * L, stream())
    p = torch.empty((B, 2), dtype=torch.float32, device=templates.device)
    call('gn_gather_rows_f32', ptr(pars), ptr(idx, torch.int32), ptr(p), B, 2, stream())
    return signal_pe.train_on_batch(batch.reshape(B, L, 1), [p[:, 0].contiguous(), p[:, 1].contiguous()])


_LABELS = {}


def _gan_labels(B, dev):
    """sy = [1]*B + [0]*B (:1289) and the generator step's [1]*B (:1296), resident on the device per batch size."""
    key = (B, str(dev))
    if key not in _LABELS:
        sy = torch.cat([torch.ones(B, device=dev), torch.zeros(B, device=dev)])
        _LABELS[key] = (sy, torch.ones(B, device=dev))
    return _LABELS[key]


def gan_train_step(generator, signal_discriminator, signal_discriminator_on_generator, noise_signal_dev,
                   real, z1, real_noise, z2, _noise_d=None, _noise_g=None, _return_device=False):
    """One iteration of the GAN loop, bbhMahoGANy.py:1241-1299, on device-resident inputs:
    real (B, n_pix) templates, z1/z2 (B,100) U(-1,1) latents, real_noise (B, n_pix) N(0,1).
    Returns (sd_loss, sg_loss) as the reference's [loss, acc] lists (device tensors with ``_return_device``: no
    host synchronisation inside the iteration)."""
    B, L = real.shape
    dev = real.device
    fake = generator.forward(z1, nn.Ctx(False))                              # generator.predict (:1248)
    sX = torch.empty((2 * B, L, 2, 1), dtype=torch.float32, device=dev)
    # real images: stack(template, N(0,1)) (:1276-1284); fake: stack(G(z), noise_signal - G(z)) (:1268-1272)
    call('gn_stack_pair_f32', ptr(real.contiguous()), ptr(real_noise.contiguous()), ptr(sX[:B]), B * L, stream())
    call('gn_stack_residual_fwd_f32', ptr(fake.reshape(B, L).contiguous()), ptr(noise_signal_dev), ptr(sX[B:]), B, L,
         stream())
    sy, ones = _gan_labels(B, dev)
    sd_loss = signal_discriminator.train_on_batch(sX, sy, _noise=_noise_d, _return_device=_return_device)      # :1292
    sg_loss = signal_discriminator_on_generator.train_on_batch(z2, ones, _noise=_noise_g,
                                                               _return_device=_return_device)                  # :1296
    return sd_loss, sg_loss


def posterior_samples(generator, signal_pe, n=4000, seed=0, batch=1000, z=None):
    """The evaluation stage of the GAN loop, bbhMahoGANy.py:1311-1343: ``generator.predict`` of ``n`` latents
    ~U(-1,1)^100 followed by ``signal_pe.predict`` of the generated waveforms, chained on the device (no host round
    trip between the two networks; latents from the Philox stream unless ``z`` (n,100) is fed in).
    Returns ``(pe_samples, generated)``: ``pe_samples`` = [mc (n,1), q (n,1)] NumPy arrays as ``signal_pe.predict``
    gives them, ``generated`` (n, n_pix, 1) NumPy."""
    dev = nn.device()
    ctx = nn.Ctx(False)
    outs, gens = None, []
    for i in range(0, n, batch):
        b = min(batch, n - i)
        if z is None:
            zz = torch.empty((b, 100), dtype=torch.float32, device=dev)
            call('gn_uniform_f32', ptr(zz), zz.numel(), -1.0, 1.0, int(seed), int(i) * 100, stream())
        else:
            zz = nn._to_device(z[i:i + b])
        g = generator.forward(zz, ctx)
        o = signal_pe.forward(g.reshape(b, -1, 1).contiguous(), ctx)
        o = o if isinstance(o, list) else [o]
        if outs is None:
            outs = [[] for _ in o]
        for k, t in enumerate(o):
            outs[k].append(t)
        gens.append(g)
    pe = [torch.cat(ts, 0).detach().float().cpu().numpy() for ts in outs]
    return pe, torch.cat(gens, 0).detach().float().cpu().numpy()


def waveform_percentiles(generated, percentiles=(90, 75, 25, 5)):
    """Percentile curves of plot_waveform_est, bbhMahoGANy.py:913-921, on the device (gn_percentiles_f32: per time
    sample a shared-memory sort of the n generated values, np.percentile's linear interpolation) instead of the
    reference's Python loop over time samples: generated (n, n_pix[, 1]) NumPy array or CUDA tensor -> {p: (n_pix,)}."""
    g = generated if isinstance(generated, torch.Tensor) else nn._to_device(np.asarray(generated, dtype=np.float32))
    g = g.to(torch.float32).reshape(g.shape[0], g.shape[1]).contiguous()
    n, L = g.shape
    pc = nn._to_device(np.asarray(percentiles, dtype=np.float32))
    out = torch.empty((len(percentiles), L), dtype=torch.float32, device=g.device)
    call('gn_percentiles_f32', ptr(g), n, L, ptr(pc), len(percentiles), ptr(out), stream())
    vals = out.cpu().numpy()
    return {p: vals[i] for i, p in enumerate(percentiles)}


class gaussian_kde2d:
    """``scipy.stats.gaussian_kde(dataset)`` for the two-dimensional (chirp mass, mass ratio) sample sets of
    make_contour_plot, bbhMahoGANy.py:787-791, evaluated on the device.  ``dataset`` (2, n) as the reference passes it.
    Bandwidth: Scott's rule ``n**(-1/6)`` on the unbiased data covariance, exactly scipy's default; the small 2x2
    algebra is done in float64 on the host, the n x m sum of Gaussians by ``gn_kde2d_pdf_f32``."""

    def __init__(self, dataset):
        d = np.atleast_2d(np.asarray(dataset, dtype=np.float64))
        if d.shape[0] != 2 or d.shape[1] < 2:
            raise ValueError('gaussian_kde2d expects a (2, n) dataset with n > 1')
        self.dataset = d
        self.d, self.n = d.shape
        self.factor = float(self.n) ** (-1.0 / (self.d + 4))
        self._data_covariance = np.atleast_2d(np.cov(d, rowvar=1, bias=False))
        self.covariance = self._data_covariance * self.factor ** 2
        self.inv_cov = np.linalg.inv(self._data_covariance) / self.factor ** 2
        self._norm_factor = np.sqrt(np.linalg.det(2 * np.pi * self.covariance)) * self.n
        self._mean = d.mean(axis=1)
        self._dev = nn._to_device(np.ascontiguousarray((d - self._mean[:, None]).T.astype(np.float32)))

    def pdf_device(self, positions):
        """positions (2, m) -> CUDA float32 tensor (m)."""
        p = np.atleast_2d(np.asarray(positions, dtype=np.float64))
        if p.shape[0] != 2:
            raise ValueError('positions must be (2, m)')
        pos = nn._to_device(np.ascontiguousarray((p - self._mean[:, None]).T.astype(np.float32)))
        out = torch.empty((p.shape[1],), dtype=torch.float32, device=pos.device)
        call('gn_kde2d_pdf_f32', ptr(self._dev), int(self.n), ptr(pos), int(p.shape[1]), float(self.inv_cov[0, 0]),
             float(self.inv_cov[0, 1]), float(self.inv_cov[1, 1]), float(1.0 / self._norm_factor), ptr(out), stream())
        return out

    def pdf(self, positions):
        return self.pdf_device(positions).cpu().numpy().astype(np.float64)

    evaluate = pdf
    __call__ = pdf


def overlap_beta(pred_samp, lalinf_samp, kernel_cnn=None, kernel_lalinf=None, n_grid=100):
    """The overlap statistic of overlap_tests, bbhMahoGANy.py:853-870: both kernel density estimates on the
    ``n_grid x n_grid`` ``np.mgrid`` spanning the pooled samples, beta = sum(p q) / sqrt(sum(p^2) sum(q^2)).
    ``pred_samp``: what ``signal_pe.predict`` returns ([mc (n,1), q (n,1)], or (n,2) for the combined model);
    ``lalinf_samp``: [mc (m,), q (m,)].  The reference's K-S and Anderson-Darling scores stay with SciPy."""
    if isinstance(pred_samp, (list, tuple)):
        px, py = np.asarray(pred_samp[0], np.float64).reshape(-1), np.asarray(pred_samp[1], np.float64).reshape(-1)
    else:
        ps = np.asarray(pred_samp, np.float64)
        px, py = ps[:, 0].reshape(-1), ps[:, 1].reshape(-1)
    lx, ly = np.asarray(lalinf_samp[0], np.float64).reshape(-1), np.asarray(lalinf_samp[1], np.float64).reshape(-1)
    comb_mc, comb_q = np.concatenate((px, lx)), np.concatenate((py, ly))
    X, Y = np.mgrid[np.min(comb_mc):np.max(comb_mc):complex(0, n_grid), np.min(comb_q):np.max(comb_q):complex(0, n_grid)]
    positions = np.vstack([X.ravel(), Y.ravel()])
    kernel_cnn = kernel_cnn or gaussian_kde2d(np.array([px, py]))
    kernel_lalinf = kernel_lalinf or gaussian_kde2d(np.array([lx, ly]))
    a, b = kernel_cnn.pdf_device(positions), kernel_lalinf.pdf_device(positions)
    sums = torch.empty((3,), dtype=torch.float64, device=a.device)
    call('gn_overlap_sums_f32', ptr(a), ptr(b), int(a.numel()), ptr(sums, torch.float64), stream())
    s = sums.cpu().numpy()
    return float(s[0] / np.sqrt(s[1] * s[2]))


def overlap_tests(pred_samp, lalinf_samp, true_vals=None, kernel_cnn=None, kernel_lalinf=None):
    """bbhMahoGANy.py:811-871, same arguments and return value ``(ks_score, ad_score, beta_score)``.  The two-sample
    K-S and Anderson-Darling scores are the reference's own SciPy calls on a few thousand host samples (:836-850); the
    overlap score runs on the device (``overlap_beta``).  ``pred_samp`` is what ``signal_pe.predict`` returned:
    [mc (n,1), q (n,1)], or (n,2) with ``comb_pe_model``; kernels default to the KDEs of the two sample sets."""
    from scipy.stats import anderson_ksamp, ks_2samp
    if isinstance(pred_samp, (list, tuple)):
        px, py = np.asarray(pred_samp[0]).reshape(-1), np.asarray(pred_samp[1]).reshape(-1)
    else:
        ps = np.asarray(pred_samp)
        px, py = ps[:, 0].reshape(-1), ps[:, 1].reshape(-1)
    lx, ly = np.asarray(lalinf_samp[0][:]).reshape(-1), np.asarray(lalinf_samp[1][:]).reshape(-1)
    ks_score = np.array([ks_2samp(px, lx), ks_2samp(py, ly)])
    ad_score = [anderson_ksamp([px, lx]), anderson_ksamp([py, ly])]
    beta_score = overlap_beta([px, py], [lx, ly], kernel_cnn=kernel_cnn, kernel_lalinf=kernel_lalinf)
    return ks_score, ad_score, beta_score
