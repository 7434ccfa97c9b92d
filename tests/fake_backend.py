"""CPU stand-in for libgennet_b200.so used ONLY by the `-m "not gpu"` tests of the host logic.

The product has no CPU path; to exercise the Python host code (graph executor, Keras protocol, arenas,
data-parallel wiring) on a box without a GPU, the tests monkeypatch `call`/`ptr`/`device` so that every C-ABI
entry point is answered by a small torch-CPU function with the semantics documented in
include/gennet_b200.h.  Nothing here ships or is importable from the package.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _act(x, kind, p):
    if kind == 1:
        return torch.relu(x)
    if kind == 2:
        return torch.tanh(x)
    if kind == 3:
        return torch.sigmoid(x)
    if kind == 4:
        return torch.where(x >= 0, x, p * x)
    if kind == 5:
        return torch.clamp(x, 0.0, p)
    if kind == 6:
        return torch.where(x > 0, x, torch.expm1(x))
    return x


def _act_bwd(y, kind, p):
    if kind == 1:
        return (y > 0).float()
    if kind == 2:
        return 1 - y * y
    if kind == 3:
        return y * (1 - y)
    if kind == 4:
        return torch.where(y >= 0, torch.ones_like(y), torch.full_like(y, p))
    if kind == 5:
        return ((y > 0) & (y < p)).float()
    if kind == 6:
        return torch.where(y > 0, torch.ones_like(y), y + 1.0)
    return torch.ones_like(y)


def _conv_in(x, B, L, Cin, up):
    x = x.reshape(B, L // up, Cin)
    if up > 1:
        x = x.repeat_interleave(up, dim=1)
    return x


def _conv(xl, w, Lout, k, s, p):
    B, L, Cin = xl.shape
    need = (Lout - 1) * s + k
    right = max(need - p - L, 0)
    xt = F.pad(xl.permute(0, 2, 1), (p, right))
    return F.conv1d(xt, w.permute(2, 1, 0), stride=s)[:, :, :Lout].permute(0, 2, 1)


class Fake:
    def __init__(self):
        self.plans = {}

    # -- plumbing --------------------------------------------------------------
    def ptr(self, t, dtype=torch.float32):
        if t is None:
            return None
        assert t.is_contiguous(), 'expected a contiguous tensor'
        if dtype is not None:
            assert t.dtype == dtype, (t.dtype, dtype)
        return t

    def stream(self):
        return None

    def call(self, name, *a):
        getattr(self, name)(*a)

    # -- layers ----------------------------------------------------------------
    def gn_conv1d_fwd_f32(self, x, w, b, y, B, L, Cin, Lout, Cout, k, s, p, up, act, ap, st):
        xl = _conv_in(x, B, L, Cin, up)
        o = _conv(xl, w.reshape(k, Cin, Cout), Lout, k, s, p)
        if b is not None:
            o = o + b
        y.reshape(B, Lout, Cout).copy_(_act(o, act, ap))

    def _conv_grads(self, x, w, dy, B, L, Cin, Lout, Cout, k, s, p, up):
        xs = x.reshape(B, L // up, Cin).clone().requires_grad_(True)
        ws = w.reshape(k, Cin, Cout).clone().requires_grad_(True)
        xl = xs.repeat_interleave(up, dim=1) if up > 1 else xs
        o = _conv(xl, ws, Lout, k, s, p)
        gx, gw = torch.autograd.grad(o, [xs, ws], dy.reshape(B, Lout, Cout))
        return gx, gw

    def gn_conv1d_dgrad_f32(self, dy, w, dx, B, L, Cin, Lout, Cout, k, s, p, up, st):
        gx, _ = self._conv_grads(torch.zeros(B, L // up, Cin), w, dy, B, L, Cin, Lout, Cout, k, s, p, up)
        dx.reshape(B, L // up, Cin).copy_(gx)

    def gn_conv1d_wgrad_f32(self, x, dy, dw, db, B, L, Cin, Lout, Cout, k, s, p, up, st):
        _, gw = self._conv_grads(x, torch.zeros(k, Cin, Cout), dy, B, L, Cin, Lout, Cout, k, s, p, up)
        dw.reshape(k, Cin, Cout).copy_(gw)
        if db is not None:
            db.copy_(dy.reshape(-1, Cout).sum(0))

    # bandwidth-bound edge layers with float32 activations: same maths as the generic kernels
    def gn_conv1d_smallcin_fwd_f32(self, x, w, b, y, B, L, Cin, Lout, Cout, k, s, p, act, ap, st):
        self.gn_conv1d_fwd_f32(x, w, b, y, B, L, Cin, Lout, Cout, k, s, p, 1, act, ap, st)

    def gn_conv1d_smallcin_wgrad_f32(self, x, dy, dw, db, B, L, Cin, Lout, Cout, k, s, p, st):
        self.gn_conv1d_wgrad_f32(x, dy, dw, db, B, L, Cin, Lout, Cout, k, s, p, 1, st)

    def gn_conv1d_smallcin_dgrad_f32(self, dy, w, dx, B, L, Cin, Lout, Cout, k, s, p, st):
        self.gn_conv1d_dgrad_f32(dy, w, dx, B, L, Cin, Lout, Cout, k, s, p, 1, st)

    def gn_conv1d_cout1_fwd_f32(self, x, w, b, y, B, L, Cin, Lout, k, p, st):
        self.gn_conv1d_fwd_f32(x, w, b, y, B, L, Cin, Lout, 1, k, 1, p, 1, 0, 0.0, st)

    def gn_conv1d_cout1_dgrad_f32(self, dy, w, dx, B, L, Cin, Lout, k, p, st):
        self.gn_conv1d_dgrad_f32(dy, w, dx, B, L, Cin, Lout, 1, k, 1, p, 1, st)

    def gn_conv1d_cout1_wgrad_f32(self, x, dy, dw, db, B, L, Cin, Lout, k, p, st):
        self.gn_conv1d_wgrad_f32(x, dy, dw, db, B, L, Cin, Lout, 1, k, 1, p, 1, st)

    def gn_dense_small_fwd_f32(self, x, w, b, y, M, K, N, act, ap, st):
        self.gn_dense_fwd_f32(x, w, b, y, M, K, N, act, ap, st)

    def gn_dense_small_wgrad_f32(self, x, dy, dw, db, M, K, N, st):
        self.gn_dense_wgrad_f32(x, dy, dw, db, M, K, N, st)

    def gn_dense_small_dgrad_f32(self, dy, w, xin, dx, colsum, C, M, K, N, in_act, ap, st):
        g = dy.reshape(M, N) @ w.reshape(K, N).t()
        if xin is not None and in_act != 0:
            g = g * _act_bwd(xin.reshape(M, K), in_act, ap)
        dx.reshape(M, K).copy_(g)
        if colsum is not None:
            colsum[:C] = g.reshape(M, K // C, C).sum((0, 1))

    # BatchNormalization -> activation -> dropout chains over float32 activations (gn_chain_*_f32)
    def _chain_noise(self, kind, rate, r, seed, off, shape):
        if kind < 0:
            return torch.ones(shape)
        if r is None:      # stand-in for the device Philox stream: any draw that is identical in forward and backward
            g = torch.Generator().manual_seed((int(seed) * 1000003 + int(off)) % (2 ** 63))
            r = torch.rand(shape, generator=g) if kind == 0 else torch.randn(shape, generator=g)
            if kind == 0:
                r = (r >= rate).float()
        r = r.reshape(shape)
        if kind == 0:
            return r / (1.0 - rate)
        return 1.0 + r * math.sqrt(rate / (1.0 - rate))

    def gn_bn_sums_f32(self, x, rows, C, sums, st):
        xv = x.reshape(rows, C).double()
        sums[:C] = xv.sum(0)
        sums[C:] = (xv * xv).sum(0)

    def _chain_pre(self, x, mean, scale, gamma, beta, use_var, eps, rows, C):
        xv = x.reshape(rows, C)
        if mean is None:
            return xv, None, None
        inv = 1.0 / torch.sqrt(scale[:C] + eps) if use_var else scale[:C]
        g = gamma[:C] if gamma is not None else torch.ones(C)
        b = beta[:C] if beta is not None else torch.zeros(C)
        xhat = (xv - mean[:C]) * inv
        return xhat * g + b, xhat, g * inv

    def gn_chain_fwd_f32(self, x, y, mean, scale, gamma, beta, use_var, eps, act, ap, noise, rate, r, seed, off, rows, C, st):
        h, _, _ = self._chain_pre(x, mean, scale, gamma, beta, use_var, eps, rows, C)
        y.reshape(rows, C).copy_(_act(h, act, ap) * self._chain_noise(noise, rate, r, seed, off, (rows, C)))

    def _chain_g(self, x, dy, mean, invstd, gamma, beta, act, ap, noise, rate, r, seed, off, rows, C):
        h, xhat, sc = self._chain_pre(x, mean, invstd, gamma, beta, 0, 0.0, rows, C)
        g = dy.reshape(rows, C) * self._chain_noise(noise, rate, r, seed, off, (rows, C)) * _act_bwd(_act(h, act, ap), act, ap)
        return g, xhat, sc

    def gn_chain_bwd_sums_f32(self, x, dy, mean, invstd, gamma, beta, act, ap, noise, rate, r, seed, off, rows, C, sums, st):
        g, xhat, _ = self._chain_g(x, dy, mean, invstd, gamma, beta, act, ap, noise, rate, r, seed, off, rows, C)
        sums[:C] = g.double().sum(0)
        sums[C:] = (g.double() * xhat.double()).sum(0)

    def gn_chain_bwd_f32(self, x, dy, dx, mean, invstd, gamma, beta, sums, n, act, ap, noise, rate, r, seed, off, dgamma, dbeta,
                         rows, C, st):
        g, xhat, sc = self._chain_g(x, dy, mean, invstd, gamma, beta, act, ap, noise, rate, r, seed, off, rows, C)
        if mean is None:
            dx.reshape(rows, C).copy_(g)
            return
        m0, m1 = (sums[:C] / n).float(), (sums[C:2 * C] / n).float()
        dx.reshape(rows, C).copy_(sc * (g - m0 - xhat * m1))
        if dbeta is not None:
            dbeta.copy_(sums[:C].float())
        if dgamma is not None:
            dgamma.copy_(sums[C:2 * C].float())

    # ---- split-operand tensor-core entry points with scaled fp16 pairs ("f16x2"): the planes are built exactly as the
    # library builds them (include/gennet_b200.h), the products are taken on the reconstructed operands
    @staticmethod
    def _f16s_scale(amax):
        a = float(amax.reshape(-1)[0])
        e = math.floor(math.log2(a)) if a > 0 else -100
        return 2.0 ** (14 - max(e, -100))

    def _split_into(self, x, planes, amax, have):
        xv = x.reshape(-1).float()
        if not have:
            amax.reshape(-1)[0] = xv.abs().max() if xv.numel() else 0.0
        s = self._f16s_scale(amax)
        t0 = (xv * s).half()
        t1 = (((xv * s) - t0.float()) * 2048.0).half()
        p = planes.reshape(2, -1)
        p[0].copy_(t0)
        p[1].copy_(t1)

    def _unsplit(self, planes, amax):
        p = planes.reshape(2, -1)
        return ((p[0].double() + p[1].double() / 2048.0) / self._f16s_scale(amax)).float()

    def gn_amax_f32(self, x, n, amax, st):
        amax.reshape(-1)[0] = x.reshape(-1)[:n].abs().max()

    def gn_split_f32_f16x2(self, x, planes, amax, have, n, st):
        self._split_into(x, planes, amax, have)

    def gn_split_colsum_f32_f16x2(self, x, planes, amax, have, rows, C, colsum, st):
        self._split_into(x, planes, amax, have)
        colsum[:C] = x.reshape(rows, C).sum(0)

    def gn_conv_w_split_f16x2(self, w, wk, wt, amax, k, Cin, Cout, st):
        self._split_into(w, wk, amax, False)
        wt.copy_(wk.reshape(2, k, Cin, Cout).permute(0, 1, 3, 2).reshape(wt.shape))

    def gn_upsample1d_fwd_bf16(self, x, y, B, L, C, size, st):      # a 16-bit copy: works on planes of either format
        y.reshape(B, L * size, C).copy_(x.reshape(B, L, C).repeat_interleave(size, dim=1))

    def _w_from_t(self, wts, wa, k, Cin, Cout):
        return self._unsplit(wts, wa).reshape(k, Cout, Cin).permute(0, 2, 1).contiguous()

    def gn_conv1d_fwd_f16x2(self, xs, xa, wts, wa, b, y, y_amax, B, L, Cin, Lout, Cout, k, s, p, act, ap, st):
        x = self._unsplit(xs, xa)
        self.gn_conv1d_fwd_f32(x, self._w_from_t(wts, wa, k, Cin, Cout), b, y, B, L, Cin, Lout, Cout, k, s, p, 1, act, ap, st)
        if y_amax is not None:
            y_amax.reshape(-1)[0] = y.abs().max()

    def gn_conv1d_fwd_stats_f16x2(self, xs, xa, wts, wa, b, y, sums, B, L, Cin, Lout, Cout, k, s, p, act, ap, st):
        self.gn_conv1d_fwd_f16x2(xs, xa, wts, wa, b, y, None, B, L, Cin, Lout, Cout, k, s, p, act, ap, st)
        self.gn_bn_sums_f32(y, B * Lout, Cout, sums, st)

    def gn_conv1d_dgrad_f16x2(self, dys, da, wks, wa, x_in, dx, colsum, dx_amax, B, L, Cin, Lout, Cout, k, s, p, in_act, ap, st):
        dy = self._unsplit(dys, da)
        self.gn_conv1d_dgrad_f32(dy, self._unsplit(wks, wa), dx, B, L, Cin, Lout, Cout, k, s, p, 1, st)
        if x_in is not None and in_act != 0:
            dx.mul_(_act_bwd(x_in.reshape(dx.shape), in_act, ap))
        if colsum is not None:
            colsum[:Cin] = dx.reshape(-1, Cin).sum(0)
        if dx_amax is not None:
            dx_amax.reshape(-1)[0] = dx.abs().max()

    def gn_conv1d_wgrad_f16x2(self, xs, xa, dys, da, dy, dw, db, B, L, Cin, Lout, Cout, k, s, p, st):
        self.gn_conv1d_wgrad_f32(self._unsplit(xs, xa), self._unsplit(dys, da), dw, None, B, L, Cin, Lout, Cout, k, s, p, 1, st)
        if db is not None:
            db.copy_(dy.reshape(-1, Cout).sum(0))

    def gn_split_pad_f32_f16x2(self, x, planes, amax, rows, K, Kp, st):
        xp = torch.zeros(rows, Kp)
        xp[:, :K] = x.reshape(rows, K)
        amax.reshape(-1)[0] = x.abs().max()
        self._split_into(xp, planes, amax, True)

    def gn_dense_w_split_f16x2(self, w, wk, wt, amax, K, Kp, N, st):
        wp = torch.zeros(Kp, N)
        wp[:K] = w.reshape(K, N)
        amax.reshape(-1)[0] = w.abs().max()
        self._split_into(wp, wk, amax, True)
        wt.copy_(wk.reshape(2, Kp, N).permute(0, 2, 1).reshape(wt.shape))

    def gn_dense_fwd_f16x2(self, xs, xa, wts, wa, b, y, M, Kp, N, act, ap, st):
        w = self._unsplit(wts, wa).reshape(N, Kp).t().contiguous()
        self.gn_dense_fwd_f32(self._unsplit(xs, xa), w, b, y, M, Kp, N, act, ap, st)

    def gn_dense_dgrad_f16x2(self, dys, da, wks, wa, x_in, dx, colsum, M, K, N, in_act, ap, st):
        g = self._unsplit(dys, da).reshape(M, N) @ self._unsplit(wks, wa).reshape(K, N).t()
        if x_in is not None and in_act != 0:
            g = g * _act_bwd(x_in.reshape(M, K), in_act, ap)
        dx.reshape(M, K).copy_(g)
        if colsum is not None:
            colsum.copy_(g.reshape(M, K // colsum.numel(), colsum.numel()).sum((0, 1)))

    def gn_dense_wgrad_f16x2(self, xs, xa, dys, da, dy, dw, db, M, K, N, Kp, st):
        x = self._unsplit(xs, xa).reshape(M, Kp)[:, :K]
        dw.reshape(K, N).copy_(x.t() @ self._unsplit(dys, da).reshape(M, N))
        if db is not None:
            db.copy_(dy.reshape(M, N).sum(0))

    def gn_chain_fwd_amax_f32(self, x, y, mean, scale, gamma, beta, use_var, eps, act, ap, noise, rate, r, seed, off, rows, C,
                              y_amax, st):
        self.gn_chain_fwd_f32(x, y, mean, scale, gamma, beta, use_var, eps, act, ap, noise, rate, r, seed, off, rows, C, st)
        y_amax.reshape(-1)[0] = y.abs().max()

    def gn_chain_fwd_planes_f32(self, x, y, mean, scale, gamma, beta, use_var, eps, act, ap, noise, rate, r, seed, off, rows, C,
                                planes, y_amax, bound, st):
        self.gn_chain_fwd_f32(x, y, mean, scale, gamma, beta, use_var, eps, act, ap, noise, rate, r, seed, off, rows, C, st)
        assert float(y.abs().max()) <= bound * (1 + 1e-6), 'a-priori bound violated'
        y_amax.reshape(-1)[0] = bound
        self._split_into(y, planes, y_amax, True)

    def gn_chain_bwd_amax_f32(self, x, dy, dx, mean, invstd, gamma, beta, sums, n, act, ap, noise, rate, r, seed, off, dgamma,
                              dbeta, rows, C, dx_amax, st):
        self.gn_chain_bwd_f32(x, dy, dx, mean, invstd, gamma, beta, sums, n, act, ap, noise, rate, r, seed, off, dgamma, dbeta,
                              rows, C, st)
        dx_amax.reshape(-1)[0] = dx.abs().max()

    def gn_gap_fwd_f32(self, x, y, B, L, C, st):
        y.reshape(B, C).copy_(x.reshape(B, L, C).mean(1))

    def gn_gap_bwd_f32(self, dy, dx, B, L, C, st):
        dx.reshape(B, L, C).copy_((dy.reshape(B, 1, C) / L).expand(B, L, C))

    def gn_transpose_f32(self, x, y, B, R, C, st):
        y.reshape(B, C, R).copy_(x.reshape(B, R, C).permute(0, 2, 1))

    def gn_reg_terms_f32(self, x, g, n, l1, l2, loss, st):
        xv = x.reshape(-1)[:n]
        if loss is not None:
            loss += (l1 * xv.abs().double().sum() + l2 * (xv.double() ** 2).sum())
        if g is not None:
            g.reshape(-1)[:n] += l1 * torch.sign(xv) + 2 * l2 * xv

    def gn_conv2d_w2_pack_f32(self, w2, b, w1, b1, kh, kw, Cin, Cout, pw, st):
        w2 = w2.reshape(kh, kw, Cin, Cout)
        o = torch.zeros(kh, 2, Cin, 2, Cout)
        for wi in range(2):
            for wo in range(2):
                q = wi - wo + pw
                if 0 <= q < kw:
                    o[:, wi, :, wo, :] = w2[:, q]
        w1.reshape(kh, 2, Cin, 2, Cout).copy_(o)
        if b1 is not None:
            b1.copy_(torch.cat([b, b]))

    def gn_conv2d_w2_unpack_f32(self, dw1, db1, dw2, db, kh, kw, Cin, Cout, pw, st):
        d = dw1.reshape(kh, 2, Cin, 2, Cout)
        o = torch.zeros(kh, kw, Cin, Cout)
        for wi in range(2):
            for wo in range(2):
                q = wi - wo + pw
                if 0 <= q < kw:
                    o[:, q] += d[:, wi, :, wo, :]
        dw2.reshape(kh, kw, Cin, Cout).copy_(o)
        if db is not None:
            db.copy_(db1[:Cout] + db1[Cout:])

    def gn_dense_fwd_f32(self, x, w, b, y, M, K, N, act, ap, st):
        o = x.reshape(M, K) @ w.reshape(K, N)
        if b is not None:
            o = o + b
        y.reshape(M, N).copy_(_act(o, act, ap))

    def gn_dense_dgrad_f32(self, dy, w, dx, M, K, N, st):
        dx.reshape(M, K).copy_(dy.reshape(M, N) @ w.reshape(K, N).t())

    def gn_dense_wgrad_f32(self, x, dy, dw, db, M, K, N, st):
        dw.reshape(K, N).copy_(x.reshape(M, K).t() @ dy.reshape(M, N))
        if db is not None:
            db.copy_(dy.reshape(M, N).sum(0))

    def gn_bn_stats_f32(self, x, rows, C, sums, shift, st):
        xv = x.reshape(rows, C).double()
        sums[:C] = xv.sum(0)
        sh = shift[:C].double() if shift is not None else 0.0
        sums[C:] = ((xv - sh) ** 2).sum(0)

    def gn_bn_finalize_f32(self, sum_x, sum_sq, n, C, eps, mom, stats, mm, mv, phase, biased, debias, st):
        if phase == 0:
            stats[:C] = (sum_x[:C] / n).float()
            return
        var = sum_sq[:C] / n
        stats[C:2 * C] = (1.0 / torch.sqrt(var + eps)).float()
        if mm is not None and biased is not None:
            biased[:C] = (biased[:C].double() * mom + stats[:C].double() * (1 - mom)).float()
            biased[C:2 * C] = (biased[C:2 * C].double() * mom + var * (n / (n - (1.0 + eps))) * (1 - mom)).float()
            mm.copy_((biased[:C].double() * debias).float())
            mv.copy_((biased[C:2 * C].double() * debias).float())
        elif mm is not None:
            mm.copy_((mm.double() * mom + stats[:C].double() * (1 - mom)).float())
            mv.copy_((mv.double() * mom + var * (n / (n - (1.0 + eps))) * (1 - mom)).float())

    def gn_bn_apply_f32(self, x, mean, inv, g, b, y, rows, C, eps, use_var, st):
        i = 1.0 / torch.sqrt(inv[:C] + eps) if use_var else inv[:C]
        y.reshape(rows, C).copy_((x.reshape(rows, C) - mean[:C]) * i * g + b)

    def gn_bn_bwd_sums_f32(self, x, dy, stats, rows, C, sums, st):
        xh = (x.reshape(rows, C) - stats[:C]) * stats[C:]
        d = dy.reshape(rows, C).double()
        sums[:C] = d.sum(0)
        sums[C:] = (d * xh.double()).sum(0)

    def gn_bn_bwd_apply_f32(self, x, dy, stats, g, sums, n, dx, dg, db, rows, C, st):
        xh = (x.reshape(rows, C) - stats[:C]) * stats[C:]
        sdy = (sums[:C] / n).float()
        sdx = (sums[C:] / n).float()
        dx.reshape(rows, C).copy_(g * stats[C:] * (dy.reshape(rows, C) - sdy - xh * sdx))
        if dg is not None:
            dg.copy_(sums[C:].float())
            db.copy_(sums[:C].float())

    def gn_act_fwd_f32(self, x, y, n, act, p, st):
        y.copy_(_act(x, act, p))

    def gn_act_bwd_f32(self, dy, y, dx, n, act, p, st):
        dx.copy_(dy * _act_bwd(y, act, p))

    def _nf(self, r, kind, rate):
        return r / (1 - rate) if kind == 0 else 1 + r * math.sqrt(rate / (1 - rate))

    def gn_noise_fwd_f32(self, x, r, y, n, kind, rate, st):
        y.copy_(x + r * rate if kind == 2 else x * self._nf(r, kind, rate))

    def gn_noise_bwd_f32(self, dy, r, dx, n, kind, rate, st):
        dx.copy_(dy if kind == 2 else dy * self._nf(r, kind, rate))

    def gn_noise_draw_f32(self, r, n, kind, rate, seed, off, st):
        g = torch.Generator().manual_seed((seed * 1000003 + off) % (2 ** 31))
        if kind == 0:
            r.copy_((torch.rand(r.shape, generator=g) >= rate).float())
        else:
            r.copy_(torch.randn(r.shape, generator=g))

    def gn_upsample1d_fwd_f32(self, x, y, B, L, C, size, st):
        y.reshape(B, L * size, C).copy_(x.reshape(B, L, C).repeat_interleave(size, dim=1))

    def gn_upsample1d_bwd_f32(self, dy, dx, B, L, C, size, st):
        dx.reshape(B, L, C).copy_(dy.reshape(B, L, size, C).sum(2))

    def gn_maxpool1d_fwd_f32(self, x, y, B, L, C, pool, st):
        y.copy_(F.max_pool1d(x.reshape(B, L, C).permute(0, 2, 1), pool).permute(0, 2, 1))

    def gn_maxpool1d_bwd_f32(self, x, y, dy, dx, B, L, C, pool, st):
        xs = x.reshape(B, L, C).clone().requires_grad_(True)
        o = F.max_pool1d(xs.permute(0, 2, 1), pool).permute(0, 2, 1)
        dx.copy_(torch.autograd.grad(o, xs, dy.reshape(o.shape))[0])

    def gn_axpy_f32(self, a, b, alpha, n, st):
        a.add_(b, alpha=alpha)

    def gn_gather_rows_f32(self, src, idx, out, n, row_len, st):
        out.reshape(n, row_len).copy_(src.reshape(-1, row_len)[idx.long()])

    def gn_kde2d_pdf_f32(self, data, n, pos, m, a11, a12, a22, inv_norm, pdf, st):
        d = data.reshape(n, 2).double()
        q = pos.reshape(m, 2).double()
        dx = d[:, 0][None, :] - q[:, 0][:, None]
        dy = d[:, 1][None, :] - q[:, 1][:, None]
        e = 0.5 * (a11 * dx * dx + 2.0 * a12 * dx * dy + a22 * dy * dy)
        pdf.copy_((torch.exp(-e).sum(1) * inv_norm).float())

    def gn_overlap_sums_f32(self, a, b, n, out3, st):
        x, y = a.reshape(-1).double(), b.reshape(-1).double()
        out3.copy_(torch.stack([(x * y).sum(), (x * x).sum(), (y * y).sum()]))

    def gn_flip_transpose_f32(self, src, out, k, A, Bn, st):
        out.reshape(k, A, Bn).copy_(src.reshape(k, Bn, A).flip(0).permute(0, 2, 1))

    def gn_maxnorm_roll_f32(self, x, off, y, B, N, st):
        xv = x.reshape(B, N)
        for b in range(B):
            sh = int(off[b]) if off is not None else 0
            y.reshape(B, N)[b].copy_(torch.roll(xv[b] / xv[b].max(), sh))

    def gn_add_scaled_f32(self, x, r, sigma, n, st):
        x.reshape(-1)[:n].add_(r.reshape(-1)[:n], alpha=sigma)

    def gn_stack_residual_fwd_f32(self, x, cst, y, B, L, st):
        xv = x.reshape(B, L)
        y.reshape(B, L, 2).copy_(torch.stack([xv, cst.reshape(L) - xv], dim=2))

    def gn_stack_residual_bwd_f32(self, dy, dx, B, L, st):
        d = dy.reshape(B, L, 2)
        dx.reshape(B, L).copy_(d[..., 0] - d[..., 1])

    def gn_residual_moments_fwd_f32(self, x, cst, sums, B, L, st):
        d = (cst.reshape(L) - x.reshape(B, L)).double()
        sums[0] = d.sum()
        sums[1] = (d * d).sum()

    def gn_residual_moments_bwd_f32(self, x, cst, dout, dx, B, L, n, st):
        d = cst.reshape(L) - x.reshape(B, L)
        dx.reshape(B, L).copy_(-(dout[0] + 2 * dout[1] * d) / n)

    def gn_loss_fwd_bwd_f32(self, pred, target, out, dpred, B, D, kind, param, inv_batch, vec, metric_kind, st):
        p = pred.reshape(1, D).expand(B, D) if vec else pred.reshape(B, D)
        t = target.reshape(B, D)
        pl = p.clone().requires_grad_(True)
        if kind == 0:
            pc = torch.clamp(pl, 1e-7, 1 - 1e-7)
            x = torch.log(pc / (1 - pc))
            l = (torch.clamp(x, min=0) - x * t + torch.log1p(torch.exp(-x.abs()))).mean(-1)
        elif kind == 1:
            l = ((pl - t) ** 2).mean(-1)
        else:
            l = (((t - pl) ** 2) / param ** 2).sum(-1)
        g = torch.autograd.grad(l.sum() * inv_batch, pl)[0]
        out[0] += l.sum().detach()
        if metric_kind == 0:
            out[1] += (t == torch.round(p)).float().mean(-1).sum()
        else:
            out[1] += (t.argmax(-1) == p.argmax(-1)).float().sum()
        if dpred is not None:
            if vec:
                dpred.add_(g.sum(0))
            else:
                dpred.reshape(B, D).copy_(g)

    def gn_adam_step_f32(self, p, g, m, v, n, lr_t, b1, b2, eps, gs, st):
        gi = g * gs
        m.mul_(b1).add_(gi, alpha=1 - b1)
        v.mul_(b2).add_(gi * gi, alpha=1 - b2)
        p.sub_(lr_t * m / (v.sqrt() + eps))

    def gn_sgd_step_f32(self, p, g, n, lr, gs, st):
        p.sub_(lr * gs * g)


def install(monkeypatch):
    """Route gennet_b200's C-ABI calls to the CPU stand-in for the duration of a test."""
    import gennet_b200.nn as nn
    import gennet_b200.bbh as bbh
    from gennet_b200 import _lib
    fake = Fake()
    cpu = torch.device('cpu')
    monkeypatch.setattr(_lib, 'require_device', lambda: None)
    import gennet_b200.wvf as wvf
    monkeypatch.setattr(wvf, 'device', lambda: cpu)
    wvf._RESAMPLE_OPS.clear()
    for mod in (nn, bbh, wvf):
        monkeypatch.setattr(mod, 'call', fake.call)
        monkeypatch.setattr(mod, 'ptr', fake.ptr)
        monkeypatch.setattr(mod, 'stream', fake.stream)
    monkeypatch.setattr(nn, 'device', lambda: cpu)

    def to_dev(x):
        if isinstance(x, torch.Tensor):
            return x.to(torch.float32).contiguous()
        return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))
    monkeypatch.setattr(nn, '_to_device', to_dev)
    nn.clear_session()
    return fake
