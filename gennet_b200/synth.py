"""Host-side mirror of BBH_version/gw_template_maker.py's synthesis API.

Same function names and argument meaning as the reference (``tukey``, ``convert_beta``, ``gen_noise``,
``whiten_data``, ``gen_masses``, ``gen_par``, ``gen_bbh``, ``make_bbh``, ``sim_data``), NumPy in / NumPy out,
plus batched device-resident entry points (``Synthesizer``) that the training loops call per batch.
All array arithmetic runs in the sm_100a kernels behind the C ABI (float32 on device; the reference is
float64 NumPy on the host).  Scalar parameter draws (gen_masses / gen_par) are host control code, as in
the reference.  LALSuite is not restated: FD waveforms, PSDs and antenna factors enter as arrays or
callables.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream
from .nn import device

safe = 2                # gw_template_maker.py:54
verb = False


class bbhparams:
    """gw_template_maker.py:69-85 (with ``fmin`` as redefined in bbhMahoGANy.py:129-144)."""

    def __init__(self, mc, M, eta, m1, m2, ra, dec, iota, phi, psi, idx, snr=None, SNR=None, fmin=None):
        self.mc, self.M, self.eta, self.m1, self.m2 = mc, M, eta, m1, m2
        self.ra, self.dec, self.iota, self.phi, self.psi = ra, dec, iota, phi, psi
        self.idx, self.fmin, self.snr, self.SNR = idx, fmin, snr, SNR


def tukey(M, alpha=0.5):
    """gw_template_maker.py:87-113 (host float64; evaluated once per plan and kept in HBM)."""
    M = int(M)
    n = np.arange(0, M)
    width = int(np.floor(alpha * (M - 1) / 2.0))
    n1, n2, n3 = n[0:width + 1], n[width + 1:M - width - 1], n[M - width - 1:]
    w1 = 0.5 * (1 + np.cos(np.pi * (-1 + 2.0 * n1 / alpha / (M - 1))))
    w3 = 0.5 * (1 + np.cos(np.pi * (-2.0 / alpha + 1 + 2.0 * n3 / alpha / (M - 1))))
    return np.concatenate((w1, np.ones(n2.shape), w3))[:M]


def convert_beta(beta, fs, T_obs):
    """gw_template_maker.py:133-159."""
    newbeta = np.array([(beta[0] + 0.5 * safe - 0.5), (beta[1] + 0.5 * safe - 0.5)]) / safe
    return int(T_obs * fs * newbeta[0]), int(T_obs * fs * newbeta[1])


def gen_masses(m_min=5.0, M_max=100.0, mdist='astro', rng=np.random):
    """gw_template_maker.py:289-370 (same rejection loops and RNG call order)."""
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
            M = (M_min ** (-7.0 / 3.0) - rng.uniform(0, 1, 1)[0] *
                 (M_min ** (-7.0 / 3.0) - M_max ** (-7.0 / 3.0))) ** (-3.0 / 7.0)
            eta = (eta_min ** (-2.0) - rng.uniform(0, 1, 1)[0] * (eta_min ** (-2.0) - 16.0)) ** (-1.0 / 2.0)
            m12 = np.zeros(2)
            m12[0] = 0.5 * M + M * np.sqrt(0.25 - eta)
            m12[1] = M - m12[0]
            if (np.sum(m12) < M_max) and np.all(m12 > m_min) and (m12[0] >= m12[1]):
                return m12, np.sum(m12) * eta ** (3.0 / 5.0), eta
    print('ERROR, unknown mass distribution. Exiting.')
    raise SystemExit(1)      # the reference prints and exit(1)s (:369-370)


def gen_par(fs, T_obs, mdist='astro', beta=[0.75, 0.95], gw_tmp=False, rng=np.random):
    """gw_template_maker.py:372-460."""
    m12, mc, eta = gen_masses(5.0, 100.0, mdist=mdist, rng=rng)
    M = np.sum(m12)
    for _ in range(5):          # iota, psi, phi, ra, dec draws (:403-416); values are then fixed (:433-437)
        rng.rand()
    if gw_tmp:
        beta = [0.5, 0.5]
    low_idx, high_idx = convert_beta(beta, fs, T_obs)
    idx = low_idx if low_idx == high_idx else int(rng.randint(low_idx, high_idx, 1)[0])
    ra, dec, iota, phi, psi = 2.21535724066, -1.23649695537, 2.5, 1.5, 1.75
    if gw_tmp:
        m1, m2 = 36.0, 29.0
        eta = m1 * m2 / (m1 + m2) ** 2
        M = m1 + m2
        return bbhparams(M * eta ** (3.0 / 5.0), M, eta, m1, m2, ra, dec, iota, phi, psi, idx)
    return bbhparams(mc, M, eta, m12[0], m12[1], ra, dec, iota, phi, psi, idx)


def _dev(a, dtype=torch.float32):
    if isinstance(a, torch.Tensor):
        return a.to(device=device(), dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a)).to(device=device(), dtype=dtype).contiguous()


class Synthesizer:
    """Device-resident state for one (fs, T_obs, psd): FFT plan, Tukey window, whitening weights,
    noise amplitudes.  Batched methods take/return CUDA tensors; nothing is staged through the host."""

    def __init__(self, fs, T_obs, psd, lead=None):
        self.fs, self.T_obs = int(fs), T_obs
        self.N = int(T_obs * fs)
        self.Nf = self.N // 2 + 1
        psd = np.asarray(psd, dtype=np.float64)
        assert psd.shape == (self.Nf,), 'psd must live on the rfft grid (N/2+1,)'
        plan = ctypes.c_void_p()
        call('gn_fft_plan_create', self.N, ctypes.byref(plan))
        self._plan = plan
        self.psd = psd
        # whiten_data :267 and :273-279 ; gen_noise :184-186 ; gen_bbh :536-538
        self.window = _dev(tukey(self.N, alpha=1.0 / 8.0))
        invpsd = np.zeros(self.Nf)
        pos = psd > 0.0
        invpsd[pos] = 1.0 / psd[pos]
        w = np.sqrt(2.0 * invpsd / fs)
        w[0] = 0.0
        self.weights = _dev(w)
        amp = np.sqrt(0.25 * T_obs * psd)
        amp[psd == 0.0] = 0.0
        self.amp = _dev(amp)
        win = np.zeros(self.N)
        tempwin = tukey(int((16.0 / 15.0) * self.N / safe), alpha=1.0 / 8.0)
        lo = int((self.N - tempwin.size) / 2)
        win[lo:lo + tempwin.size] = tempwin
        self.signal_window = _dev(win)
        self.crop_lo = int(((T_obs / 2) * fs) - fs / 2)       # :695
        self.crop_len = int(((T_obs / 2) * fs) + fs / 2) - self.crop_lo
        self.lead = 11 if lead is None else int(lead)          # :554 ("use 21 if sampling at 2kHz")

    def __del__(self):
        try:
            if getattr(self, '_plan', None):
                call('gn_fft_plan_destroy', self._plan)
                self._plan = None
        except Exception:
            pass

    # -- whiten_data(..., 'td') ---------------------------------------------------------------
    def whiten_td(self, x, crop=False, scale=1.0):
        """whiten_data(..., 'td') (:243-286) of a batch; x: CUDA tensor, or a host array (copied through pinned staging)."""
        if isinstance(x, np.ndarray) and x.dtype == np.float32:
            from . import nn
            x = nn._to_device(x)
        x = _dev(x).reshape(-1, self.N)
        lo, ln = (self.crop_lo, self.crop_len) if crop else (0, self.N)
        y = torch.empty((x.shape[0], ln), dtype=torch.float32, device=x.device)
        call('gn_whiten_td_f32', self._plan, ptr(x), ptr(self.window), ptr(self.weights), ptr(y), x.shape[0], lo, ln,
             float(scale), stream())
        return y

    # -- irfft(xf * weights), rolled ------------------------------------------------------------
    def irfft(self, xf, weights=None, scale=1.0, roll=0, drop_dc=False):
        """xf: (batch, Nf) complex64 tensor / complex ndarray."""
        if not isinstance(xf, torch.Tensor):
            xf = torch.as_tensor(np.ascontiguousarray(np.asarray(xf).astype(np.complex64)))
        xf = xf.to(device()).to(torch.complex64).reshape(-1, self.Nf).contiguous()
        xr = torch.view_as_real(xf).contiguous()
        y = torch.empty((xf.shape[0], self.N), dtype=torch.float32, device=xf.device)
        call('gn_irfft_f32', self._plan, ptr(xr), ptr(weights) if weights is not None else None, ptr(y), xf.shape[0],
             float(scale), int(roll), 1 if drop_dc else 0, stream())
        return y

    def gen_noise(self, normals):
        """gen_noise (:161-193) for a batch of fed-in normals (batch, 2, Nf)."""
        nrm = _dev(normals).reshape(-1, 2, self.Nf)
        xf = torch.complex(nrm[:, 0], nrm[:, 1])
        return self.irfft(xf, weights=self.amp, scale=float(self.N) / self.T_obs, drop_dc=True)

    # -- fused synthesis ------------------------------------------------------------------------
    def synth(self, batch, templates=None, tidx=None, normals=None, scale=1.0, seed=0, sample_offset=0, out=None):
        """out[b] = scale * crop(whiten_td(gen_noise() + templates[tidx[b]])); see gn_synth_f32."""
        if out is None:
            out = torch.empty((batch, self.crop_len), dtype=torch.float32, device=device())
        nt = 0 if templates is None else templates.shape[0]
        call('gn_synth_f32', self._plan, ptr(normals) if normals is not None else None, ptr(self.amp),
             ptr(templates) if templates is not None else None,
             ptr(tidx, torch.int32) if tidx is not None else None, ptr(self.window), ptr(self.weights), ptr(out),
             int(batch), int(nt), self.crop_lo, self.crop_len, float(self.N) / self.T_obs, float(scale), int(seed),
             int(sample_offset), stream())
        return out

    # -- gen_bbh after the waveform call ----------------------------------------------------------
    def bbh_from_fd(self, hp_fd, hc_fd, idx, Fp, Fc, crop=False, scale=1.0):
        """Batched gen_bbh :518-575 (one detector): returns (ts, ref_idx) as CUDA tensors."""
        hp = self.irfft(hp_fd, weights=self.weights, roll=-self.fs, drop_dc=True)
        hc = self.irfft(hc_fd, weights=self.weights, roll=-self.fs, drop_dc=True)
        B = hp.shape[0]
        lo, ln = (self.crop_lo, self.crop_len) if crop else (0, self.N)
        out = torch.empty((B, ln), dtype=torch.float32, device=hp.device)
        ref = torch.empty(B, dtype=torch.int32, device=hp.device)
        # keep the staged arguments alive until the launch is enqueued (a freed temporary's block could be
        # handed to the next staging copy before the kernel reads it)
        fp_d = _dev(np.broadcast_to(np.asarray(Fp, np.float32), (B,)))
        fc_d = _dev(np.broadcast_to(np.asarray(Fc, np.float32), (B,)))
        idx_d = _dev(np.broadcast_to(np.asarray(idx, np.int32), (B,)), torch.int32)
        call('gn_bbh_assemble_f32', ptr(hp), ptr(hc), ptr(fp_d), ptr(fc_d), ptr(idx_d, torch.int32),
             self.lead, ptr(self.signal_window), ptr(out), ptr(ref, torch.int32), B, self.N, lo, ln, float(scale),
             stream())
        return out, ref

    def norm_constant(self, wht_wvf):
        """gw_norm_constant = 1/std(wht_wvf), gw_template_maker.py:782."""
        x = _dev(wht_wvf).reshape(-1)
        out = torch.empty(2, dtype=torch.float32, device=x.device)
        call('gn_mean_std_f32', ptr(x), x.numel(), ptr(out), stream())
        return 1.0 / float(out[1].item())


_SYNTH_CACHE = {}


def _synth_for(fs, T_obs, psd):
    psd = np.asarray(psd, dtype=np.float64)
    key = (int(fs), float(T_obs), psd.shape, hash(psd.tobytes()))
    s = _SYNTH_CACHE.get(key)
    if s is None:
        if len(_SYNTH_CACHE) > 8:
            _SYNTH_CACHE.clear()
        s = _SYNTH_CACHE[key] = Synthesizer(fs, T_obs, psd)
    return s


def gen_noise(fs, T_obs, psd, normals=None, rng=np.random):
    """gw_template_maker.py:161-193: NumPy in / NumPy out (float64 container, float32 arithmetic on the GPU).
    The two np.random.normal(0,1,Nf) draws (:187-188) are made on the host unless `normals` (2,Nf) is given."""
    s = _synth_for(fs, T_obs, psd)
    if normals is None:
        normals = np.stack([rng.normal(0, 1, s.Nf), rng.normal(0, 1, s.Nf)])
    return s.gen_noise(np.asarray(normals, np.float32)[None])[0].cpu().numpy().astype(np.float64)


def whiten_data(data, duration, sample_rate, psd, flag='td'):
    """gw_template_maker.py:243-286.  'td': real series in, whitened series out; 'fd': complex half
    spectrum in, weighted half spectrum out (an elementwise product, done on the device too)."""
    s = _synth_for(sample_rate, duration, psd)
    if flag == 'td':
        return s.whiten_td(np.asarray(data, np.float32)[None])[0].cpu().numpy().astype(np.float64)
    xf = torch.as_tensor(np.asarray(data).astype(np.complex64)).to(device())
    out = xf * s.weights
    out[0] = 0
    return out.cpu().numpy().astype(np.complex128)


def make_bbh(hp, hc, fs, ra, dec, psi, det, antenna=None):
    """gw_template_maker.py:577-630: ht = Fp*hp + Fc*hc.  The spline time shift the reference computes is
    discarded by it (:621-630) and is not reproduced.  `antenna(ra,dec,psi,det)->(Fp,Fc)` stands in for
    pylal.antenna.response (:612)."""
    Fp, Fc = antenna(ra, dec, psi, det) if antenna is not None else (1.0, 0.0)
    hp = np.asarray(hp)
    hc = np.asarray(hc)
    return hp * Fp + hc * Fc, hp, hc


def gen_bbh(fs, T_obs, psds, dets=['H1'], beta=[0.75, 0.95], par=None, gw_tmp=False, waveform=None, antenna=None):
    """gw_template_maker.py:462-575 with the LAL call replaced by ``waveform(par, fs, T_obs) -> (hp_fd, hc_fd)``
    on the rfft grid.  Returns (ts, hp, hc, ts) with ts of shape (1, N), as the reference does."""
    assert waveform is not None, 'LALSuite is not available: pass waveform(par, fs, T_obs) -> (hp_fd, hc_fd)'
    s = _synth_for(fs, T_obs, psds)
    hp_fd, hc_fd = waveform(par, fs, T_obs)
    Fp, Fc = antenna(par.ra, par.dec, par.psi, dets[0]) if antenna is not None else (1.0, 0.0)
    ts, _ = s.bbh_from_fd(np.asarray(hp_fd)[None], np.asarray(hc_fd)[None], par.idx, Fp, Fc)
    tsn = ts.cpu().numpy().astype(np.float64)
    hp, _ = s.bbh_from_fd(np.asarray(hp_fd)[None], np.asarray(hc_fd)[None], par.idx, 1.0, 0.0)
    hc, _ = s.bbh_from_fd(np.asarray(hp_fd)[None], np.asarray(hc_fd)[None], par.idx, 0.0, 1.0)
    return tsn, hp.cpu().numpy().astype(np.float64), hc.cpu().numpy().astype(np.float64), tsn


def sim_data(fs, T_obs, psds, dets=['H1'], Nnoise=25, size=1000, mdist='astro', beta=[0.75, 0.95], waveform=None,
             antenna=None, gw_tmp=True, rng=np.random, chunk=256):
    """gw_template_maker.py:632-740 (one detector, do_time_grid off as shipped).  Template synthesis is
    batched on the GPU in chunks; returns ([ts (size,1,fs), yval], list[bbhparams]).

    ``Nnoise > 0`` (:684-692, "not typically used"): every template is followed by ``Nnoise`` noise realisations and the
    counter advances by one per realisation, so ``ceil(size / Nnoise)`` templates are drawn and ``ceil(size / Nnoise) *
    Nnoise`` rows come back, each the FULL-length whitened series (``N = T_obs * fs`` samples: this branch does not
    crop).  With the module default ``gw_tmp = True`` the reference then fails at :737, concatenating the ``(1, 1, fs)``
    GW150914-like template onto ``(rows, 1, N)`` rows; that combination raises here as well."""
    assert waveform is not None, 'LALSuite is not available: pass waveform(par, fs, T_obs) -> (hp_fd, hc_fd)'
    s = _synth_for(fs, T_obs, psds)
    n = size - 1 if gw_tmp else size
    if Nnoise > 0:
        if gw_tmp:
            raise ValueError('sim_data(Nnoise > 0, gw_tmp=True): the reference concatenates a (1,1,fs) template onto '
                             'full-length noisy rows and fails (gw_template_maker.py:737); use gw_tmp=False')
        n = -(-n // Nnoise)          # templates drawn before the row counter reaches `size`
    pars = [gen_par(fs, T_obs, mdist=mdist, beta=beta, gw_tmp=False, rng=rng) for _ in range(n)]
    ts = []
    for c0 in range(0, n, chunk):
        ps = pars[c0:c0 + chunk]
        fd = [waveform(p, fs, T_obs) for p in ps]
        hp = np.stack([f[0] for f in fd])
        hc = np.stack([f[1] for f in fd])
        ant = [antenna(p.ra, p.dec, p.psi, dets[0]) if antenna is not None else (1.0, 0.0) for p in ps]
        idx = np.array([p.idx for p in ps], dtype=np.int32)
        Fp = np.array([a[0] for a in ant], np.float32)
        Fc = np.array([a[1] for a in ant], np.float32)
        if Nnoise > 0:
            full, _ = s.bbh_from_fd(hp, hc, idx, Fp, Fc, crop=False)
            per_t = []
            for j in range(Nnoise):
                nrm = rng.normal(0, 1, (len(ps), 2, s.Nf)).astype(np.float32)
                noisy = s.gen_noise(_dev(nrm)) + full                           # ts_noise + ts_new (:687-688)
                per_t.append(s.whiten_td(noisy, crop=False).cpu().numpy())
            ts.append(np.stack(per_t, axis=1).reshape(len(ps) * Nnoise, -1))    # template-major, as the reference appends
        else:
            out, _ = s.bbh_from_fd(hp, hc, idx, Fp, Fc, crop=True)
            ts.append(out.cpu().numpy())
    if Nnoise > 0:
        pars = [p for p in pars for _ in range(Nnoise)]
    ts = np.concatenate(ts).astype(np.float64)[:, None, :] if ts else np.zeros((0, 1, fs))
    yval = np.ones(len(ts), dtype=int)
    order = rng.permutation(len(ts))
    ts, yval = ts[order], yval[order]
    pars = [pars[i] for i in order]
    if gw_tmp:
        p = gen_par(fs, T_obs, mdist=mdist, beta=beta, gw_tmp=True, rng=rng)
        hp_fd, hc_fd = waveform(p, fs, T_obs)
        Fp, Fc = antenna(p.ra, p.dec, p.psi, dets[0]) if antenna is not None else (1.0, 0.0)
        out, _ = s.bbh_from_fd(np.asarray(hp_fd)[None], np.asarray(hc_fd)[None], p.idx, Fp, Fc, crop=True)
        ts = np.concatenate((ts, out.cpu().numpy().astype(np.float64).reshape(1, 1, -1)))
        pars.append(p)
        yval = np.append(yval, 1)
    return [ts, yval], pars


def make_template_bank(event_fd, event_noise_fd, psd, fs=1024, T_obs=2, Nsamp=1000, Nblock=1000, Nnoise=0,
                       dets=['H1'], mdist='astro', oversamp=True, basename='templates/', event_name='gw150914',
                       sample_num=None, tag='', waveform=None, antenna=None, rng=np.random, write=True):
    """The generation loop of ``main()`` (gw_template_maker.py:742-849) with the lalinference text files replaced
    by arrays on the rfft grid of N = safe*T_obs*fs samples: ``event_fd`` = event in noise
    (``...freqDataWithInjection.dat``), ``event_noise_fd`` = noise alone (``...freqData.dat``), ``psd`` (``...PSD.dat``
    column 1).  Whitens the event FD -> TD (:774-777), takes ``gw_norm_constant = 1/std`` of the whitened event over
    the full segment (:782), generates ``ceil(Nsamp/Nblock)`` blocks with :func:`sim_data` (``hunt_constrain`` prior
    when ``oversamp``, beta [0.45, 0.55], :806-808), scales every template by the constant (:813-814), drops the
    appended GW150914-like template from all but the last block (:817-819) and pickles each block the way
    ``cPickle.dump(.., protocol=HIGHEST_PROTOCOL)`` did (:840-849):
        <basename><event>_ts_<i>_<sample_num>Samp<tag>.sav       [ts (n,1,fs) float64, yval]
        <basename><event>_params_<i>_<sample_num>Samp<tag>.sav   list[bbhparams]
    Returns a dict with the norm constant, the cropped whitened event / noise-free event, and the file names."""
    from . import io as gio
    safeT = safe * T_obs
    N = int(fs * safeT)
    event_fd = np.array(event_fd, dtype=np.complex128)
    noise_fd = np.array(event_noise_fd, dtype=np.complex128)
    event_fd[np.isnan(event_fd)] = 0
    noise_fd[np.isnan(noise_fd)] = 0
    h_fd = event_fd - noise_fd
    s = _synth_for(fs, safeT, psd)
    # whiten_data(.., 'fd') then np.fft.irfft(.., N): spectral weights (DC zeroed) fused into the inverse transform
    wht = s.irfft(np.stack([event_fd, h_fd]), weights=s.weights).cpu().numpy().astype(np.float64)
    wht_wvf, h_t = wht[0], wht[1]
    gw_norm_constant = 1.0 / np.std(wht_wvf)
    lo, hi = int((safeT / 2) * fs - fs / 2.0), int((safeT / 2) * fs + fs / 2.0)
    wht_wvf, h_t = wht_wvf[lo:hi], h_t[lo:hi]
    nblock = int(np.ceil(float(Nsamp) / float(Nblock)))
    sample_num = Nsamp if sample_num is None else sample_num
    files = []
    last = None
    for i in range(nblock):
        ts, par = sim_data(fs, safeT, psd, dets, Nnoise, size=Nblock, mdist='hunt_constrain' if oversamp else mdist,
                           beta=[0.45, 0.55], waveform=waveform, antenna=antenna, rng=rng)
        ts[0] = ts[0] * gw_norm_constant
        if i != nblock - 1:
            ts[0] = ts[0][:-1]
            par = par[:-1]
        if write:
            f_ts = '%s%s_ts_%d_%sSamp%s.sav' % (basename, event_name, i, sample_num, tag)
            f_par = '%s%s_params_%d_%sSamp%s.sav' % (basename, event_name, i, sample_num, tag)
            d = os.path.dirname(f_ts)
            if d and not os.path.isdir(d):
                os.makedirs(d)
            gio.dump_pickle(ts, f_ts)
            gio.dump_pickle(par, f_par)
            files.append((f_ts, f_par))
        last = (ts, par)
    return {'gw_norm_constant': gw_norm_constant, 'event': wht_wvf, 'event_noise_free': h_t, 'files': files,
            'last_block': last}


def make_burst_waveforms(N_sig, amp=1, freq=100, dt=1.0 / 512, N=512, t_0=0.5, phi=2 * np.pi, tau=1.0 / 25.0,
                         rand5=None, rng=None):
    """tests/burstMahoGANy.py:76-98 on the device; (t0, tau) are drawn on the host with `random.uniform`
    in the reference's order."""
    import random as _random
    r = rng if rng is not None else _random
    pars = np.empty((N_sig, 2), dtype=np.float64)
    for i in range(N_sig):
        if rand5 is True:
            t_0 = r.uniform(0.25, 0.75)
            tau = r.uniform(1.0 / 60.0, 1.0 / 15.0)
        pars[i] = (t_0, tau)
    out = torch.empty((N_sig, N), dtype=torch.float32, device=device())
    pars_d = _dev(pars, torch.float64)
    call('gn_burst_waveforms_f32', ptr(pars_d, torch.float64), ptr(out), N_sig, N, float(amp), float(freq), float(dt),
         float(phi), stream())
    return out.cpu().numpy().astype(np.float64), pars


def analytic_psd(fs, T_obs, f_low=10.0):
    """Declared synthetic aLIGO-like PSD (stand-in for lalsimulation.SimNoisePSD*, :217-233)."""
    N = int(T_obs * fs)
    f = np.arange(N // 2 + 1) / float(T_obs)
    x = np.maximum(f, 1e-3) / 215.0
    s = 1e-49 * (x ** -4.14 - 5.0 * x ** -2 + 111.0 * (1 - x ** 2 + 0.5 * x ** 4) / (1 + 0.5 * x ** 2))
    s[f < f_low] = 0.0
    return s


def newtonian_chirp_fd(par, fs, T_obs, f_low=40.0, dist_mpc=410.0):
    """Declared synthetic TaylorF2-style FD chirp (stand-in for SimInspiralChooseFDWaveform, :507-516)."""
    G, c, Msun, pc = 6.67430e-11, 299792458.0, 1.98847e30, 3.085677581491367e16
    N = int(T_obs * fs)
    f = np.arange(N // 2 + 1) / float(T_obs)
    M = (par.m1 + par.m2) * Msun * G / c ** 3
    eta = par.m1 * par.m2 / (par.m1 + par.m2) ** 2
    mc = M * eta ** 0.6
    D = dist_mpc * 1e6 * pc / c
    f_isco = 1.0 / (6 ** 1.5 * np.pi * M)
    h = np.zeros(f.size, dtype=np.complex128)
    band = (f >= f_low) & (f <= min(2.5 * f_isco, fs / 2.0))
    fb = f[band]
    v = (np.pi * M * fb) ** (1.0 / 3.0)
    amp = np.sqrt(5.0 / 24.0) * np.pi ** (-2.0 / 3.0) * mc ** (5.0 / 6.0) / D * fb ** (-7.0 / 6.0)
    amp = amp / (1.0 + (fb / (1.3 * f_isco)) ** 6)
    psi = -par.phi - np.pi / 4 + 3.0 / (128 * eta) * v ** -5 * (
        1 + 20.0 / 9.0 * (743.0 / 336.0 + 11.0 / 4.0 * eta) * v ** 2 - 16 * np.pi * v ** 3)
    h[band] = amp * np.exp(-1j * psi)
    ci = np.cos(par.iota)
    return 0.5 * (1 + ci ** 2) * h, -1j * ci * h
