"""The algebra the register organisation of the whitening kernel relies on (gennet_b200/csrc/synth.cu), checked in
NumPy float64 without a GPU: packing a real series into a half-length complex FFT, the two-coefficient whitening step
Z'[k] = alpha_k Z[k] + beta_k i conj(Z[M-k]) that whiten_coef_kernel tabulates, and the partner ownership
M - (tid + j*M/16) = (M/16 - tid) + (15 - j)*M/16 that makes the exchange a single conflict-free read."""
import numpy as np
import pytest

from oracle import synth_oracle as so


@pytest.mark.parametrize('N', [512, 8192])
def test_two_coefficient_whitening_step_equals_rfft_multiply_irfft(N):
    rs = np.random.RandomState(N)
    M = N // 2
    x = rs.normal(size=N)
    w = rs.uniform(0.5, 2.0, M + 1)
    w[0] = 0.0
    ref = np.fft.irfft(np.fft.rfft(x) * w, N)
    Z = np.fft.fft(x[0::2] + 1j * x[1::2])
    k = np.arange(M)
    wk, wm = w[k], w[M - k]
    th = 2.0 * np.pi * k / N
    alpha = 0.5 * (wk + wm) - 0.5 * (wk - wm) * np.sin(th)
    beta = 0.5 * (wk - wm) * np.cos(th)
    Zm = Z[(M - k) % M]
    Zp = alpha * Z + beta * (Zm.imag + 1j * Zm.real)          # i * conj(Zm) = Zm.imag + i Zm.real
    y = np.fft.ifft(Zp)
    out = np.empty(N)
    out[0::2], out[1::2] = y.real, y.imag
    assert np.abs(out - ref).max() < 1e-12 * np.abs(ref).max()


def test_whitening_step_matches_the_oracle_whiten_data():
    fs, T = 512, 4
    N = fs * T
    psd = so.analytic_psd(fs, T)
    rs = np.random.RandomState(3)
    x = rs.normal(size=N) * 1e-21
    ref = so.whiten_data(x, T, fs, psd, 'td')
    w = so.whiten_weights(psd, fs)
    xw = x * so.tukey(N, alpha=1.0 / 8.0)
    M = N // 2
    Z = np.fft.fft(xw[0::2] + 1j * xw[1::2])
    k = np.arange(M)
    th = 2.0 * np.pi * k / N
    alpha = 0.5 * (w[k] + w[M - k]) - 0.5 * (w[k] - w[M - k]) * np.sin(th)
    beta = 0.5 * (w[k] - w[M - k]) * np.cos(th)
    Zm = Z[(M - k) % M]
    y = np.fft.ifft(alpha * Z + beta * (Zm.imag + 1j * Zm.real))
    out = np.empty(N)
    out[0::2], out[1::2] = y.real, y.imag
    assert np.abs(out - ref).max() < 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize('M', [256, 4096, 16384])
def test_partner_bins_live_in_the_mirrored_thread(M):
    NT = M // 16
    tid = np.arange(NT)[:, None]
    j = np.arange(16)[None, :]
    k = tid + j * NT
    partner = (M - k) % M
    # exchange buffer layout [j][tid]: the partner of (tid, j) is slot ((16 - j) * NT - tid) mod M, contiguous in tid
    assert np.array_equal(partner, ((16 - j) * NT - tid) % M)
    own_thread, own_slot = partner % NT, partner // NT
    assert np.array_equal(own_thread, np.broadcast_to((NT - tid) % NT, (NT, 16)))
    assert np.array_equal(own_slot[1:], np.broadcast_to(15 - j, (NT, 16))[1:])        # tid >= 1
    assert np.array_equal(own_slot[0], (16 - np.arange(16)) % 16)                      # thread 0 pairs with itself


def test_product_twiddles_stay_within_two_ulp():
    """Six table rows (r = 1,2,3,4,8,12) and nine float32 products replace fifteen loads per butterfly."""
    P = 256
    k = np.arange(P)
    tw = lambda r: np.exp(-2j * np.pi * r * k / (16.0 * P))
    c64 = lambda z: z.real.astype(np.float32) + 1j * z.imag.astype(np.float32)
    rows = {r: c64(tw(r)) for r in (1, 2, 3, 4, 8, 12)}
    worst = 0.0
    for r in range(1, 16):
        a, b = r >> 2, r & 3
        if a == 0 or b == 0:
            w = rows[r]
        else:
            wh, wl = rows[4 * a], rows[b]
            re = (wh.real * wl.real).astype(np.float32) - (wh.imag * wl.imag).astype(np.float32)
            im = (wh.real * wl.imag).astype(np.float32) + (wh.imag * wl.real).astype(np.float32)
            w = re.astype(np.float32) + 1j * im.astype(np.float32)
        worst = max(worst, np.abs(w - tw(r)).max())
    assert worst < 2.0 * 2.0 ** -23
