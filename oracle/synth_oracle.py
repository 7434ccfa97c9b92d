"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the sample-synthesis half of the hot path.

This is a float64 NumPy restatement of the reference's template/noise synthesis.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it; the product (``gennet_b200``) never does.

Each function cites the reference lines it follows (paths relative to the
reference checkout, ``BBH_version/gw_template_maker.py`` unless noted).

Pinning: ``tukey``, ``gen_noise``, ``whiten_data``, ``convert_beta`` and
``make_burst_waveforms`` are pinned against the *reference's own source* executed
under Python 3 (``tests/golden/make_golden.py`` extracts the function bodies from
the read-only checkout and records their outputs in ``tests/golden/synth_ref.npz``).
``gen_bbh`` depends on LALSuite (absent) and is pinned only through the
arithmetic after the LAL call, with FD polarisations fed in as arrays.
"""
import numpy as np

safe = 2  # gw_template_maker.py:54


class bbhparams:
    """gw_template_maker.py:69-85 (+ ``fmin`` as in bbhMahoGANy.py:129-144)."""

    def __init__(self, mc, M, eta, m1, m2, ra, dec, iota, phi, psi, idx, snr=None, SNR=None, fmin=None):
        self.mc, self.M, self.eta, self.m1, self.m2 = mc, M, eta, m1, m2
        self.ra, self.dec, self.iota, self.phi, self.psi = ra, dec, iota, phi, psi
        self.idx, self.fmin, self.snr, self.SNR = idx, fmin, snr, SNR


def tukey(M, alpha=0.5):
    """gw_template_maker.py:87-113."""
    M = int(M)
    n = np.arange(0, M)
    width = int(np.floor(alpha * (M - 1) / 2.0))
    n1 = n[0:width + 1]
    n2 = n[width + 1:M - width - 1]
    n3 = n[M - width - 1:]
    w1 = 0.5 * (1 + np.cos(np.pi * (-1 + 2.0 * n1 / alpha / (M - 1))))
    w2 = np.ones(n2.shape)
    w3 = 0.5 * (1 + np.cos(np.pi * (-2.0 / alpha + 1 + 2.0 * n3 / alpha / (M - 1))))
    return np.concatenate((w1, w2, w3))[:M]


def convert_beta(beta, fs, T_obs):
    """gw_template_maker.py:133-159."""
    newbeta = np.array([(beta[0] + 0.5 * safe - 0.5), (beta[1] + 0.5 * safe - 0.5)]) / safe
    return int(T_obs * fs * newbeta[0]), int(T_obs * fs * newbeta[1])


def noise_amplitude(T_obs, psd):
    """amp of gw_template_maker.py:184-186."""
    psd = np.asarray(psd, dtype=np.float64)
    amp = np.sqrt(0.25 * T_obs * psd)
    amp[psd == 0.0] = 0.0
    return amp


def gen_noise(fs, T_obs, psd, normals=None, rng=None):
    """gw_template_maker.py:161-193.

    ``normals`` (2, Nf) replaces the two ``np.random.normal(0,1,Nf)`` draws
    (:187-188, re first then im) so the RNG is excluded from parity.
    """
    N = int(T_obs * fs)
    Nf = N // 2 + 1
    df = 1.0 / T_obs
    amp = noise_amplitude(T_obs, psd)
    if normals is None:
        rng = np.random if rng is None else rng
        normals = np.stack([rng.normal(0, 1, Nf), rng.normal(0, 1, Nf)])
    re = amp * normals[0]
    im = amp * normals[1]
    re[0] = 0.0
    im[0] = 0.0
    return N * np.fft.irfft(re + 1j * im) * df


def whiten_weights(psd, sample_rate):
    """sqrt(2*invpsd/fs) with undefined bins zeroed and DC removed (:273-279)."""
    psd = np.asarray(psd, dtype=np.float64)
    invpsd = np.zeros(psd.size)
    pos = psd > 0.0
    invpsd[pos] = 1.0 / psd[pos]
    w = np.sqrt(2.0 * invpsd / sample_rate)
    w[0] = 0.0
    return w


def whiten_data(data, duration, sample_rate, psd, flag='td'):
    """gw_template_maker.py:243-286 (does not mutate ``data``; the reference's
    in-place multiply of an 'fd' input is an aliasing quirk callers never rely on)."""
    if flag == 'td':
        win = tukey(duration * sample_rate, alpha=1.0 / 8.0)
        xf = np.fft.rfft(win * np.asarray(data, dtype=np.float64))
    else:
        xf = np.array(data, dtype=np.complex128)
    xf = xf * whiten_weights(psd, sample_rate)
    xf[0] = 0.0
    if flag == 'td':
        return np.fft.irfft(xf)
    return xf


def crop_central(t, fs, T_obs):
    """central ``fs`` samples, gw_template_maker.py:695."""
    lo = int(((T_obs / 2) * fs) - fs / 2)
    hi = int(((T_obs / 2) * fs) + fs / 2)
    return t[..., lo:hi]


def signal_window(N):
    """aggressive window of gen_bbh, :536-538."""
    win = np.zeros(N)
    tempwin = tukey(int((16.0 / 15.0) * N / safe), alpha=1.0 / 8.0)
    lo = int((N - tempwin.size) / 2)
    win[lo:lo + tempwin.size] = tempwin
    return win


def make_bbh(hp, hc, Fp, Fc):
    """gw_template_maker.py:577-630 with the antenna pattern fed in.

    The spline time shift (:621-628) is computed and discarded by the
    reference, so the returned series are unshifted (:630)."""
    return hp * Fp + hc * Fc, hp, hc


def place_signal(ht, ref_idx, idx, N, lead=11):
    """gw_template_maker.py:554-565 including Python's negative-start wrap."""
    start = int(ref_idx - idx - lead)
    tmp = ht[start:]
    out = np.zeros(N)
    if len(tmp) < N:
        out[:len(tmp)] = tmp
    else:
        out[:] = tmp[:N]
    return out


def gen_bbh_from_fd(hp_fd, hc_fd, fs, T_obs, psd, idx, Fp, Fc, lead=11):
    """Everything in gen_bbh after the LAL call (:518-575), one detector.

    hp_fd/hc_fd: FD polarisations on the rfft grid (Nf,), as
    ``SimInspiralChooseFDWaveform(...).data.data`` would deliver them."""
    N = int(T_obs * fs)
    whp = whiten_data(hp_fd, T_obs, fs, psd, flag='fd')
    whc = whiten_data(hc_fd, T_obs, fs, psd, flag='fd')
    orig_hp = np.roll(np.fft.irfft(whp, N), int(-fs))
    orig_hc = np.roll(np.fft.irfft(whc, N), int(-fs))
    ref_idx = int(np.argmax(orig_hp ** 2 + orig_hc ** 2))
    win = signal_window(N)
    ht, hp, hc = make_bbh(orig_hp, orig_hc, Fp, Fc)
    ts = place_signal(ht, ref_idx, idx, N, lead) * win
    return ts, ref_idx


def synth_sample(template, normals, fs, T_obs, psd, scale=1.0):
    """Intent of sim_data's noise branch (:685-691):
    whiten(gen_noise(psd)+h,'td') -> central crop -> x gw_norm_constant (:813-814)."""
    x = gen_noise(fs, T_obs, psd, normals=normals) + template
    w = whiten_data(x, T_obs, fs, psd, flag='td')
    return crop_central(w, fs, T_obs) * scale


def gw_norm_constant(wht_wvf):
    """gw_template_maker.py:782."""
    return 1.0 / np.std(wht_wvf)


def gen_masses(rng, m_min=5.0, M_max=100.0, mdist='astro'):
    """gw_template_maker.py:289-370 (rng = np.random.RandomState-like)."""
    log_m_max = np.log(M_max - m_min)
    if mdist in ('astro', 'hunt_constrain'):
        while True:
            m12 = np.exp(np.log(m_min) + rng.uniform(0, 1, 2) * (log_m_max - np.log(m_min)))
            eta = m12[0] * m12[1] / (m12[0] + m12[1]) ** 2
            mc = np.sum(m12) * eta ** (3.0 / 5.0)
            ok = (np.sum(m12) < M_max) and np.all(m12 > m_min) and (m12[0] >= m12[1])
            if mdist == 'hunt_constrain':
                ok = ok and (m12[1] / m12[0] >= 0.5) and (mc >= 20.0) and (mc <= 35.0)
            if ok:
                return m12, mc, eta
    elif mdist == 'gh':
        m12 = np.zeros(2)
        while True:
            q = rng.uniform(1.0, 10.0, 1)
            m12[1] = rng.uniform(5.0, 75.0, 1)[0]
            m12[0] = m12[1] * q[0]
            if np.all(m12 < 75.0) and np.all(m12 > 5.0) and (m12[0] >= m12[1]):
                break
        eta = m12[0] * m12[1] / (m12[0] + m12[1]) ** 2
        return m12, np.sum(m12) * eta ** (3.0 / 5.0), eta
    elif mdist == 'metric':
        M_min = 2.0 * m_min
        eta_min = m_min * (M_max - m_min) / M_max ** 2
        while True:
            M = (M_min ** (-7.0 / 3.0) - rng.uniform(0, 1, 1)[0] * (M_min ** (-7.0 / 3.0) - M_max ** (-7.0 / 3.0))) ** (-3.0 / 7.0)
            eta = (eta_min ** (-2.0) - rng.uniform(0, 1, 1)[0] * (eta_min ** (-2.0) - 16.0)) ** (-1.0 / 2.0)
            m12 = np.zeros(2)
            m12[0] = 0.5 * M + M * np.sqrt(0.25 - eta)
            m12[1] = M - m12[0]
            if (np.sum(m12) < M_max) and np.all(m12 > m_min) and (m12[0] >= m12[1]):
                return m12, np.sum(m12) * eta ** (3.0 / 5.0), eta
    raise ValueError('unknown mass distribution')


def gen_par(rng, fs, T_obs, mdist='astro', beta=(0.75, 0.95), gw_tmp=False):
    """gw_template_maker.py:372-460 (same RNG call order)."""
    m12, mc, eta = gen_masses(rng, 5.0, 100.0, mdist)
    M = np.sum(m12)
    rng.rand(); rng.rand(); rng.rand(); rng.rand(); rng.rand()  # iota, psi, phi, ra, dec draws (:403-416), overwritten
    if gw_tmp:
        beta = [0.5, 0.5]
    lo, hi = convert_beta(beta, fs, T_obs)
    idx = lo if lo == hi else int(rng.randint(lo, hi, 1)[0])
    ra, dec, iota, phi, psi = 2.21535724066, -1.23649695537, 2.5, 1.5, 1.75
    if gw_tmp:
        m1, m2 = 36.0, 29.0
        eta = m1 * m2 / (m1 + m2) ** 2
        M = m1 + m2
        return bbhparams(M * eta ** (3.0 / 5.0), M, eta, m1, m2, ra, dec, iota, phi, psi, idx)
    return bbhparams(mc, M, eta, m12[0], m12[1], ra, dec, iota, phi, psi, idx)


def make_burst_waveforms(N_sig, amp=1, freq=100, dt=1.0 / 512, N=512, t_0=0.5, phi=2 * np.pi,
                         tau=1.0 / 25.0, rand5=None, draws=None):
    """tests/burstMahoGANy.py:76-98; ``draws`` (N_sig,2) replaces random.uniform."""
    data, pars = [], []
    for i in range(N_sig):
        if rand5 is True:
            t_0, tau = draws[i]
        t = dt * np.arange(0, N, 1)
        data.append(amp * np.sin(2 * np.pi * freq * (t - t_0) + phi) * np.exp(-(t - t_0) ** 2 / (tau ** 2)))
        pars.append([t_0, tau])
    return np.array(data), np.array(pars)


def sample_data(draws, x_vals=np.arange(0, 5, .1), max_offset=100, mul_range=(1, 2)):
    """train_on_wvf_version/nn.py:58-70; ``draws`` (n,2) replaces np.random.random()."""
    out = []
    for u0, u1 in draws:
        offset = u0 * max_offset
        mul = mul_range[0] + u1 * (mul_range[1] - mul_range[0])
        out.append(np.sin(offset + x_vals * mul) / 2 + .5)
    return np.array(out)


# --------------------------------------------------------------------------
# Declared synthetic stand-ins for the LALSuite inputs (NOT reference code):
# used by bench / tests to make inputs of the reference's shape.
# --------------------------------------------------------------------------

def analytic_psd(fs, T_obs, f_low=10.0):
    """aLIGO-like analytic one-sided PSD on the rfft grid, zero below f_low
    (LAL's SimNoisePSD* fills zeros below its flow=10 Hz argument, :221)."""
    N = int(T_obs * fs)
    f = np.arange(N // 2 + 1) / float(T_obs)
    x = np.maximum(f, 1e-3) / 215.0
    s = 1e-49 * (x ** -4.14 - 5.0 * x ** -2 + 111.0 * (1 - x ** 2 + 0.5 * x ** 4) / (1 + 0.5 * x ** 2))
    s[f < f_low] = 0.0
    return s


def newtonian_chirp_fd(m1, m2, fs, T_obs, iota=2.5, phi=1.5, f_low=40.0, dist_mpc=410.0):
    """TaylorF2-style (Newtonian phase + 1PN term) FD chirp with an ISCO taper;
    a declared stand-in for SimInspiralChooseFDWaveform(IMRPhenomPv2) (:507-516)."""
    G, c, Msun, pc = 6.67430e-11, 299792458.0, 1.98847e30, 3.085677581491367e16
    N = int(T_obs * fs)
    f = np.arange(N // 2 + 1) / float(T_obs)
    M = (m1 + m2) * Msun * G / c ** 3
    eta = m1 * m2 / (m1 + m2) ** 2
    mc = M * eta ** 0.6
    D = dist_mpc * 1e6 * pc / c
    f_isco = 1.0 / (6 ** 1.5 * np.pi * M)
    h = np.zeros(f.size, dtype=np.complex128)
    band = (f >= f_low) & (f <= min(2.5 * f_isco, fs / 2.0))
    fb = f[band]
    v = (np.pi * M * fb) ** (1.0 / 3.0)
    amp = np.sqrt(5.0 / 24.0) * np.pi ** (-2.0 / 3.0) * mc ** (5.0 / 6.0) / D * fb ** (-7.0 / 6.0)
    amp = amp / (1.0 + (fb / (1.3 * f_isco)) ** 6)
    psi = 2 * np.pi * fb * (T_obs * 0.0) - phi - np.pi / 4 + 3.0 / (128 * eta) * v ** -5 * (
        1 + 20.0 / 9.0 * (743.0 / 336.0 + 11.0 / 4.0 * eta) * v ** 2 - 16 * np.pi * v ** 3)
    h[band] = amp * np.exp(-1j * psi)
    ci = np.cos(iota)
    return 0.5 * (1 + ci ** 2) * h, -1j * ci * h


def resample_fft(x, num):
    """``scipy.signal.resample(x, num)`` as scipy 1.1.0 (the version the reference pins, requirements.txt:42)
    computes it for a real 1-D series: FFT, keep bins [0, (N+1)//2) and the last (N-1)//2 with N = min(num, Nx)
    -- for even N the bin N/2 is NOT carried over, unlike scipy >= 1.4 -- inverse FFT, scale num/Nx.
    Call site: train_on_wvf_version/load_txtwfs.py:48,66."""
    x = np.asarray(x, dtype=np.float64)
    Nx = x.shape[-1]
    X = np.fft.fft(x, axis=-1)
    N = int(min(num, Nx))
    Y = np.zeros(x.shape[:-1] + (num,), dtype=np.complex128)
    Y[..., 0:(N + 1) // 2] = X[..., 0:(N + 1) // 2]
    if (N - 1) // 2 > 0:
        Y[..., -((N - 1) // 2):] = X[..., -((N - 1) // 2):]
    return (np.fft.ifft(Y, axis=-1) * (float(num) / float(Nx))).real


def ingest_waveform(x, offset, num=512):
    """load_txtwfs.py:47-50: resample to `num` points, divide by the maximum, roll by `offset`."""
    d = resample_fft(x, num)
    d = d / np.max(d)
    return np.roll(d, offset)
