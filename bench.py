#!/usr/bin/env python
"""Benchmark of the GenNet hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config gan|pe] [--mode f16x2|bf16x3|bf16x2|bf16|fp32] [--check]

Workloads (BASELINE.json `configs`, metric "train samples/sec (synth+whiten+G/D step)"):
  gan (default, configs[2]): bbhMahoGANy.py GAN waveform estimator, n_pix 2048, 128 samples per GPU (global batch 1024
      on 8 GPUs), SyncBN.  One step = one iteration of the loop at bbhMahoGANy.py:1241-1299 with its inputs made on
      the device: whitening of the batch's templates (gn_whiten_td_f32), a PSD-coloured whitened noise channel
      (gn_synth_f32), generator.predict, discriminator train step on 2B images, generator train step through the
      frozen discriminator.
  pe  (configs[1]): CNN point estimator, n_pix 2048, batch 512 per GPU.  One step = Philox noise + injected chirp ->
      whiten -> crop (gn_synth_f32) fused with signal_pe.train_on_batch.  Always measured too and reported under
      `extra.pe` of the same JSON line.
Modes: f16x2 (default) = float32 tensors, Conv1D / Conv2D / Dense on the tcgen05 tensor cores with scaled fp16-pair
operands (three MMAs per product, float32-class accuracy: every rtol-1e-4 whole-step parity test runs in this mode);
bf16x3 = the same with three bf16 planes (six MMAs per product, element-wise 24-bit operands); bf16x2; bf16 (throughput
mode, stated tolerance); fp32 (SIMT).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS, T_OBS, N_TEMPLATES = 2048, 4, 1024
PE_BATCH, GAN_BATCH = 512, 128
UNIT = 'samples/s'
METRICS = {'gan': 'train samples/sec (synth+whiten+G/D step)', 'pe': 'train samples/sec (synth+whiten+CNN-PE step)'}
MODES = {'f16x2': 'f16x2', 'bf16x3': 'bf16x3', 'bf16x2': 'bf16x2', 'bf16': 'bfloat16', 'fp32': 'float32'}
DTYPES = {'f16x2': 'f32 (scaled fp16-pair operands on tcgen05, fp32 accumulate)',
          'bf16x3': 'f32 (3-plane split-bf16 operands on tcgen05, fp32 accumulate)',
          'bf16x2': 'f32 (2-plane split-bf16 operands on tcgen05, fp32 accumulate)', 'bf16': 'bf16', 'fp32': 'f32'}
PLANE_PRODUCTS = {'f16x2': 3, 'bf16x3': 6, 'bf16x2': 3, 'bf16': 1}


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return p['hbm_gbs'], p['bf16_tflops'], p.get('bf16_tflops_sustained', p['bf16_tflops']), 'measured'
    except Exception:
        return 6650.0, 1590.0, 1400.0, 'fallback'


class ClockSampler:
    """SM clock and throttle reasons of one GPU sampled DURING the timed region (NVML from a thread;
    `nvidia-smi -lms` block-buffers its stdout into a pipe, so short regions would see no samples)."""
    BITS = (('hw_slowdown', 0x8), ('sw_power_cap', 0x4), ('sw_thermal_slowdown', 0x20),
            ('hw_thermal_slowdown', 0x40), ('hw_power_brake', 0x80))

    def __init__(self, index, period_s=0.02):
        self.index, self.period, self.samples, self.h, self.nv = index, period_s, [], None, None
        self._stop = threading.Event()
        self.t = None

    def start(self):
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:                                   # honour CUDA_VISIBLE_DEVICES remapping
                uuid = 'GPU-' + str(torch.cuda.get_device_properties(self.index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._sample()
            self.samples.clear()
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:                     # pragma: no cover (no NVML on the CPU container)
            self.h, self.err = None, repr(e)
        return self

    def _sample(self):
        nv = self.nv
        mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        try:
            power = nv.nvmlDeviceGetPowerUsage(self.h) / 1e3
        except Exception:
            power = None
        self.samples.append((mhz, reasons, power))

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop.wait(self.period)

    def stop(self):
        if self.h is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvml unavailable: %s' % getattr(self, 'err', '')],
                    'samples': 0}
        self._stop.set()
        self.t.join(timeout=2)
        try:
            self._sample()                         # at least one sample even for a very short region
        except Exception:
            pass
        sm = [s[0] for s in self.samples]
        mask = 0
        for s in self.samples:
            mask |= s[1]
        pw = [s[2] for s in self.samples if s[2] is not None]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(n for n, b in self.BITS if mask & b), 'samples': len(sm),
                'power_w_max': max(pw) if pw else None}


def ncu_traffic(pattern):
    """DRAM bytes per launch of a kernel family from the latest committed `ncu --set full` capture
    (profiles/rNN_<pattern>.json, written by profiles/summarize_step_traffic.py); (None, None) if absent."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r[0-9][0-9]_' + pattern + '.json')))
    if not files:
        return None, None
    try:
        with open(files[-1]) as f:
            d = json.load(f)
        return d['dram_bytes_per_launch'], os.path.basename(files[-1])
    except Exception:
        return None, None


def make_inputs(seed, device):
    """Synthetic chirp bank: unwhitened time-domain templates (n,N) f32 in HBM, labels (mc/35, q), analytic PSD."""
    import torch
    from gennet_b200 import synth
    rs = np.random.RandomState(seed)
    psd = synth.analytic_psd(FS, T_OBS)
    s = synth.Synthesizer(FS, T_OBS, psd)
    pars, hp = [], []
    for i in range(N_TEMPLATES):
        p = synth.gen_par(FS, T_OBS, mdist='hunt_constrain', beta=[0.45, 0.55], rng=rs)
        pars.append([p.mc / 35.0, p.m2 / p.m1])
        h, _ = synth.newtonian_chirp_fd(p, FS, T_OBS)
        hp.append(h)
    # unwhitened strain templates in the time domain (what gets injected into coloured noise)
    td = s.irfft(np.stack(hp), scale=float(FS), roll=-FS)
    return s, td.contiguous(), torch.as_tensor(np.array(pars, dtype=np.float32)).to(device)


# ---- algorithmic flops of a GEMM-shaped entry point, from its own arguments ---------------------------------------
# position of B (followed by L, Cin, Lout, Cout, k) in the argument list of every Conv1D entry point
_CONV_ARGPOS = {'gn_conv1d_fwd_f32': 4, 'gn_conv1d_dgrad_f32': 3, 'gn_conv1d_wgrad_f32': 4,
                'gn_conv1d_fwd_bf16': 4, 'gn_conv1d_dgrad_bf16': 5, 'gn_conv1d_wgrad_bf16': 4,
                'gn_conv1d_fwd_bf16x3': 5, 'gn_conv1d_dgrad_bf16x3': 6, 'gn_conv1d_wgrad_bf16x3': 5,
                'gn_conv1d_fwd_f16x2': 7, 'gn_conv1d_fwd_stats_f16x2': 7, 'gn_conv1d_dgrad_f16x2': 8, 'gn_conv1d_wgrad_f16x2': 7,
                'gn_conv1d_smallcin_fwd_bf16': 4, 'gn_conv1d_smallcin_wgrad_bf16': 4, 'gn_conv1d_smallcin_dgrad_bf16': 3,
                'gn_conv1d_smallcin_fwd_f32': 4, 'gn_conv1d_smallcin_wgrad_f32': 4, 'gn_conv1d_smallcin_dgrad_f32': 3}
_DENSE_ARGPOS = {'gn_dense_fwd_f32': 4, 'gn_dense_dgrad_f32': 3, 'gn_dense_wgrad_f32': 4,
                 'gn_dense_fwd_bf16x3': 5, 'gn_dense_dgrad_bf16x3': 5, 'gn_dense_wgrad_bf16x3': 5,
                 'gn_dense_fwd_f16x2': 6, 'gn_dense_dgrad_f16x2': 7, 'gn_dense_wgrad_f16x2': 7}
TENSOR_CORE_CALLS = ('gn_conv1d_fwd_bf16', 'gn_conv1d_dgrad_bf16', 'gn_conv1d_wgrad_bf16', 'gn_conv1d_fwd_bf16x3',
                     'gn_conv1d_dgrad_bf16x3', 'gn_conv1d_wgrad_bf16x3', 'gn_dense_fwd_bf16x3', 'gn_dense_dgrad_bf16x3',
                     'gn_dense_wgrad_bf16x3', 'gn_conv1d_fwd_f16x2', 'gn_conv1d_fwd_stats_f16x2', 'gn_conv1d_dgrad_f16x2', 'gn_conv1d_wgrad_f16x2',
                     'gn_dense_fwd_f16x2', 'gn_dense_dgrad_f16x2', 'gn_dense_wgrad_f16x2')


def call_flops(name, args):
    """2*L_out*k*Cin*Cout per sample and pass for a convolution, 2*M*K*N for a Dense GEMM."""
    if name in _CONV_ARGPOS:
        B, L, Cin, Lout, Cout, k = [int(v) for v in args[_CONV_ARGPOS[name]:_CONV_ARGPOS[name] + 6]]
        return 2.0 * B * Lout * k * Cin * Cout
    if name in _DENSE_ARGPOS:
        M, K, N = [int(v) for v in args[_DENSE_ARGPOS[name]:_DENSE_ARGPOS[name] + 3]]
        return 2.0 * M * K * N
    return 0.0


def profile_calls(step_fn, n=3):
    """Per-entry-point device time with CUDA events on the launching stream over n extra steps:
    {name: [ms per step, launches per step, algorithmic flops per step]}."""
    import torch
    from gennet_b200 import _lib
    torch.cuda.synchronize()
    _lib.PROFILE = []
    for it in range(n):
        step_fn(10 ** 6 + it)
    torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    tot = {}
    for name, tag, a, b, args in prof:
        d = tot.setdefault(name, [0.0, 0.0, 0.0])
        d[0] += a.elapsed_time(b) / n
        d[1] += 1.0 / n
        d[2] += call_flops(name, args) / n
    return tot


def tensor_roofline(tot, mode, step_ms_events):
    """Roofline of the dominant kernel family (tensor-core Conv1D / Dense implicit GEMMs of one step)."""
    hbm, tf_burst, tf_sust, which = peaks()
    names = [k for k in tot if k in TENSOR_CORE_CALLS]
    if not names:       # fp32 mode: the SIMT GEMM family stands in (no tensor-core launches exist)
        names = [k for k in tot if k in _CONV_ARGPOS or k in _DENSE_ARGPOS]
    ms = sum(tot[k][0] for k in names)
    launches = sum(tot[k][1] for k in names)
    flops = sum(tot[k][2] for k in names)
    ach = flops / (ms / 1e3) / 1e12 if ms > 0 else 0.0
    prod = PLANE_PRODUCTS.get(mode, 1)
    traffic, src = ncu_traffic({'f16x2': 'conv_f16x2_step_traffic', 'bf16x3': 'conv_tc3_step_traffic', 'bf16x2': 'conv_tc3_step_traffic'}.get(
        mode, 'conv_tc_step_traffic'))
    all_ms = sum(v[0] for v in tot.values())
    return {'bound': 'tensor', 'kernel': 'tcgen05 Conv1D/Dense implicit GEMM family (%s)' % ', '.join(sorted(names)),
            'achieved': ach, 'peak': tf_burst, 'unit': 'TFLOP/s', 'frac': ach / tf_burst,
            'frac_burst': ach / tf_burst, 'frac_sustained': ach / tf_sust, 'peak_sustained': tf_sust,
            'peak_source': which + ' (cuBLAS bf16; burst = kernel timed alone, sustained = inside a long step)',
            'plane_products_per_flop': prod, 'issued_tensor_tflops': ach * prod,
            'tensor_pipe_frac_burst': ach * prod / tf_burst, 'tensor_pipe_frac_sustained': ach * prod / tf_sust,
            'note': ('achieved counts ALGORITHMIC float32 flops; every product is issued as %d 16-bit tcgen05.mma '
                     '(split operands), so the tensor pipe executes achieved x %d' % (prod, prod)) if prod > 1 else
                    'bf16 operands: one tcgen05.mma per product',
            'traffic': traffic, 'traffic_source': src, 'avg_launch_ms': ms / max(launches, 1e-9),
            'launches_per_step': launches, 'share_of_step': ms / all_ms if all_ms else None,
            'algorithmic_flops_per_step': flops}


WHITEN_BATCHES = (8192, 32768)      # series per launch; the last one is the headline (2.1 GB in + out per launch)


def whiten_roofline(synth_obj, dev, batches=WHITEN_BATCHES, iters=20, sample_clocks=None):
    """BASELINE's second metric, "whitening HBM GB/s": gn_whiten_td_f32 alone (window -> rfft -> weights -> irfft) on
    `batch` resident series of N = 8192 samples, timed with CUDA events; algorithmic bytes 8*N per series (read + write
    the series once; window / weights / twiddles are batch-shared).  >= 537 MB per launch >> 126 MB L2."""
    import torch
    hbm, _, _, which = peaks()
    N = synth_obj.N
    by_batch, clocks, ms = {}, None, None
    for batch in batches:
        x = torch.randn(batch, N, device=dev) * 1e-21
        for _ in range(3):
            synth_obj.whiten_td(x)
        torch.cuda.synchronize()
        sampler = ClockSampler(sample_clocks, period_s=0.001).start() if sample_clocks is not None else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            synth_obj.whiten_td(x)
        e1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler is not None else None
        ms = e0.elapsed_time(e1) / iters
        gbs = batch * 8 * N / (ms / 1e3) / 1e9
        by_batch[str(batch)] = {'achieved': gbs, 'frac': gbs / hbm, 'avg_launch_ms': ms, 'clocks': clocks}
        del x
    batch = batches[-1]
    gbs = by_batch[str(batch)]['achieved']
    traffic = None
    try:
        import glob
        files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r[0-9][0-9]_whiten_traffic.json')))
        with open(files[-1]) as f:
            t = json.load(f)                  # ncu --set full capture of one launch; scaled per series if the batch differs
        traffic = t['dram_bytes_per_launch'] * (float(batch) / t.get('batch', 8192))
    except Exception:
        pass
    return {'bound': 'hbm', 'kernel': 'synth_kernel<12,0,1> (gn_whiten_td_f32: Tukey window, rfft, whitening weights, irfft; '
                                      'the timed call includes its ~3 us coefficient prologue whiten_coef_kernel)',
            'achieved': gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': gbs / hbm, 'traffic': traffic, 'peak_source': which,
            'avg_launch_ms': ms, 'algorithmic_bytes_per_launch': batch * 8 * N, 'batch': batch, 'iters': iters,
            'clocks': clocks, 'by_batch': by_batch}


# ---- workloads -----------------------------------------------------------------------------------------------
class Workload:
    def __init__(self, dev, rank, world, synth_obj, templates, labels):
        self.dev, self.rank, self.world = dev, rank, world
        self.synth, self.templates, self.labels = synth_obj, templates, labels
        self.L, self.N = FS, FS * T_OBS


class PEWorkload(Workload):
    name, B = 'pe', PE_BATCH

    def __init__(self, *a):
        super().__init__(*a)
        import torch
        from gennet_b200 import nn, bbh, parallel
        bbh.n_pix = FS
        self.model = bbh.signal_pe_model()
        self.model.compile(loss='mean_squared_error', optimizer=nn.Adam(lr=9e-5, beta_1=0.5), metrics=['accuracy'])
        if self.world > 1:
            parallel.broadcast_weights(self.model)
        B, L, dev = self.B, self.L, self.dev
        self.gen = torch.Generator(device=dev).manual_seed(100 + self.rank)
        self.batch = torch.empty((B, L), dtype=torch.float32, device=dev)
        self.tgt = torch.empty((B, 2), dtype=torch.float32, device=dev)
        rs = np.random.RandomState(5 + self.rank)
        self.host_x = [(rs.normal(size=(B, self.N)) * 1e-21).astype(np.float32) for _ in range(2)]
        self.host_y = [rs.uniform(0.5, 1.0, (B,)).astype(np.float32), rs.uniform(0.5, 1.0, (B,)).astype(np.float32)]
        self.h2d = B * self.N * 4 + 2 * B * 4
        self.d2h = 5 * 4
        self._stage = None

    def describe(self):
        return ('bbhMahoGANy.py CNN point estimator (signal_pe_model), fs 2048 Hz, N=8192 -> n_pix 2048, batch %d per GPU, '
                'synthetic TaylorF2-style chirps + analytic aLIGO-like PSD, random-init weights' % self.B)

    def step(self, it):
        """Device-resident step: rank r synthesises global samples [it*B*world + r*B, ...) (the Philox stream depends on
        the global sample index only), gathers their labels and runs train_on_batch."""
        import torch
        from gennet_b200 import _lib
        B, L = self.B, self.L
        idx = torch.randint(0, N_TEMPLATES, (B,), device=self.dev, generator=self.gen, dtype=torch.int32)
        self.synth.synth(B, templates=self.templates, tidx=idx, scale=1.0, seed=2024,
                         sample_offset=(it * self.world + self.rank) * B, out=self.batch)
        _lib.call('gn_gather_rows_f32', _lib.ptr(self.labels), _lib.ptr(idx, torch.int32), _lib.ptr(self.tgt), B, 2,
                  _lib.stream())
        return self.model.train_on_batch(self.batch.reshape(B, L, 1),
                                         [self.tgt[:, 0].contiguous(), self.tgt[:, 1].contiguous()], _return_device=True)

    def api_step(self, it):
        """The call a script makes, host arrays in, host floats out: whiten the host strain batch (H2D inside
        Synthesizer.whiten_td), then signal_pe.train_on_batch with host labels (bbhMahoGANy.py:1165)."""
        w = self.synth.whiten_td(self.host_x[it % 2], crop=True, scale=1.0)
        return self.model.train_on_batch(w.reshape(self.B, self.L, 1), self.host_y)

    def overlapped_step(self, it):
        """A loader thread's job done with a copy stream: the pinned H2D copy of batch i+1 overlaps step i."""
        import torch
        if self._stage is None:
            self._stage = Stager(self.dev, [(self.B, self.N), (self.B,), (self.B,)],
                                 lambda k: [self.host_x[k % 2], self.host_y[0], self.host_y[1]])
        x, y0, y1 = self._stage.get(it)
        w = self.synth.whiten_td(x, crop=True, scale=1.0)
        r = self.model.train_on_batch(w.reshape(self.B, self.L, 1), [y0, y1])
        self._stage.done(it)
        return r


class GANWorkload(Workload):
    name, B = 'gan', GAN_BATCH

    def __init__(self, *a):
        super().__init__(*a)
        import torch
        from gennet_b200 import nn, bbh, parallel
        bbh.n_pix = FS
        rs = np.random.RandomState(11)                 # every rank: the same event (noise_signal of main(), :1088)
        B, L, dev = self.B, self.L, self.dev
        ev = self.synth.whiten_td(self.templates[:1].contiguous(), crop=True, scale=1.0)
        self.norm = 1.0 / float(ev.std().item())       # gw_norm_constant (gw_template_maker.py:782)
        noise_signal = (ev.cpu().numpy().reshape(L, 1) * self.norm + rs.normal(size=(L, 1))).astype(np.float32)
        self.G, self.D, self.DG, _ = bbh.build_gan(noise_signal)
        if self.world > 1:
            for m in (self.G, self.D):
                parallel.broadcast_weights(m)
        self.ns = torch.as_tensor(noise_signal.reshape(-1)).to(dev)
        self.gen = torch.Generator(device=dev).manual_seed(200 + self.rank)
        self.raw = torch.empty((B, self.N), dtype=torch.float32, device=dev)
        self.z = [torch.empty((B, 100), dtype=torch.float32, device=dev) for _ in range(2)]
        hr = np.random.RandomState(7 + self.rank)
        # host-side bank of whitened, normalised templates (what the reference's loop samples from, :1244)
        wt = self.synth.whiten_td(self.templates[:256].contiguous(), crop=True, scale=self.norm)
        self.host_bank = wt.cpu().numpy().astype(np.float32)
        self.host_rng = hr
        self.h2d = (B * 100 * 4) * 2 + 2 * B * L * 2 * 4 + 2 * B * 4 + B * 4
        self.d2h = B * L * 4 + 4 * 4

    def describe(self):
        return ('bbhMahoGANy.py GAN waveform estimator (generator_model + signal_discriminator_model, loop :1241-1299), '
                'n_pix 2048, batch %d per GPU (global %d), SyncBN across ranks, synthetic TaylorF2-style chirps + analytic '
                'aLIGO-like PSD, random-init weights' % (self.B, self.B * self.world))

    def step(self, it):
        """Device-resident iteration: gather + whiten the batch's templates, synthesise the whitened PSD-coloured noise
        channel, draw the two latent batches (Philox), then predict + D step + G step; nothing returns to the host."""
        import torch
        from gennet_b200 import _lib, bbh
        B, dev = self.B, self.dev
        idx = torch.randint(0, N_TEMPLATES, (B,), device=dev, generator=self.gen, dtype=torch.int32)
        _lib.call('gn_gather_rows_f32', _lib.ptr(self.templates), _lib.ptr(idx, torch.int32), _lib.ptr(self.raw), B, self.N,
                  _lib.stream())
        real = self.synth.whiten_td(self.raw, crop=True, scale=self.norm)
        off = (it * self.world + self.rank) * B
        noise_ch = self.synth.synth(B, templates=None, scale=1.0, seed=77, sample_offset=off)
        for k in range(2):
            _lib.call('gn_uniform_f32', _lib.ptr(self.z[k]), B * 100, -1.0, 1.0, 4242 + k, off * 100, _lib.stream())
        sd, sg = bbh.gan_train_step(self.G, self.D, self.DG, self.ns, real, self.z[0], noise_ch, self.z[1],
                                    _return_device=True)
        return torch.cat([sd, sg])

    def api_step(self, it):
        """The reference's loop body (bbhMahoGANy.py:1244-1296) verbatim on host arrays through the Keras protocol:
        generator.predict -> host restack -> signal_discriminator.train_on_batch -> stacked model train_on_batch."""
        B, L, rs = self.B, self.L, self.host_rng
        signal = self.host_bank[rs.randint(0, self.host_bank.shape[0], B)].reshape(B, L, 1)
        noise = rs.uniform(-1.0, 1.0, (B, 100)).astype(np.float32)
        generated = self.G.predict(noise)
        fake = np.concatenate((generated, self.ns_host() - generated), axis=2)
        real = np.concatenate((signal, rs.normal(0, 1, (B, L, 1)).astype(np.float32)), axis=2)
        sX = np.concatenate((real, fake)).reshape(2 * B, L, 2, 1)
        sy = [1.0] * B + [0.0] * B
        sd = self.D.train_on_batch(sX, sy)
        noise = rs.uniform(-1.0, 1.0, (B, 100)).astype(np.float32)
        sg = self.DG.train_on_batch(noise, [1] * B)
        return sd + sg

    def ns_host(self):
        if not hasattr(self, '_ns_host'):
            self._ns_host = self.ns.cpu().numpy().reshape(1, self.L, 1)
        return self._ns_host

    def overlapped_step(self, it):
        """Device-chained iteration (bbh.gan_train_step) fed from pinned host buffers on a copy stream: raw strain
        (B,N) for the real channel and the two latent batches are copied H2D every step, losses come back D2H."""
        if getattr(self, '_stage', None) is None:
            rs = np.random.RandomState(3 + self.rank)
            hx = [(rs.normal(size=(self.B, self.N)) * 1e-21).astype(np.float32) for _ in range(2)]
            hz = [rs.uniform(-1, 1, (self.B, 100)).astype(np.float32) for _ in range(2)]
            self._stage = Stager(self.dev, [(self.B, self.N), (self.B, 100), (self.B, 100)],
                                 lambda k: [hx[k % 2], hz[0], hz[1]])
            self.h2d_overlapped = self.B * self.N * 4 + 2 * self.B * 100 * 4
        from gennet_b200 import bbh
        x, z1, z2 = self._stage.get(it)
        real = self.synth.whiten_td(x, crop=True, scale=1.0)
        noise_ch = self.synth.synth(self.B, templates=None, scale=1.0, seed=78, sample_offset=it * self.B)
        sd, sg = bbh.gan_train_step(self.G, self.D, self.DG, self.ns, real, z1, noise_ch, z2)
        self._stage.done(it)
        return sd + sg


class Stager:
    """Double-buffered pinned staging on a copy stream: get(i) waits for batch i's copy, starts batch i+1's."""

    def __init__(self, dev, shapes, host_fn):
        import torch
        self.torch, self.dev, self.host_fn = torch, dev, host_fn
        self.stream = torch.cuda.Stream(device=dev)
        self.pinned = [[torch.empty(s, dtype=torch.float32).pin_memory() for s in shapes] for _ in range(2)]
        self.devb = [[torch.empty(s, dtype=torch.float32, device=dev) for s in shapes] for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        for ev in self.consumed:
            ev.record()
        self.next = None

    def _issue(self, it):
        torch, slot = self.torch, it % 2
        for p, h in zip(self.pinned[slot], self.host_fn(it)):
            p.copy_(torch.from_numpy(np.ascontiguousarray(h)))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.consumed[slot])      # the step that last used this slot has finished
            for d, p in zip(self.devb[slot], self.pinned[slot]):
                d.copy_(p, non_blocking=True)
            self.ready[slot].record(self.stream)
        self.next = it + 1

    def get(self, it):
        if self.next is None or self.next <= it:
            self._issue(it)
        self.torch.cuda.current_stream().wait_event(self.ready[it % 2])
        self._issue(it + 1)                                   # next batch's copy overlaps this step
        return self.devb[it % 2]

    def done(self, it):
        self.consumed[it % 2].record()


def timed(fn, first, steps, barrier, world, dev):
    """EXACTLY `steps` calls between CUDA events on the current stream, barrier + synchronize on both sides, max over ranks."""
    import torch
    import torch.distributed as dist
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    res = None
    for it in range(steps):
        res = fn(first + it)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), res


def measure(w, steps, warmup, barrier, world, dev, rank, local, profile=True, mode='f16x2'):
    """value (device-resident), e2e (public API, host arrays), e2e_overlapped, launch count, clocks, roofline."""
    import torch
    from gennet_b200 import _lib
    for it in range(warmup):
        w.step(it)
    sampler = ClockSampler(local).start() if rank == 0 else None
    calls0 = _lib.COUNTS['calls']
    ms, res = timed(w.step, warmup, steps, barrier, world, dev)
    clocks = sampler.stop() if sampler is not None else None
    launches = _lib.COUNTS['calls'] - calls0
    last = res.detach().cpu().numpy()
    assert np.isfinite(last).all(), 'training diverged'
    out = {'value': w.B * world * steps / (ms / 1e3), 'ms_per_step': ms / steps, 'gpu_launches': launches, 'clocks': clocks}
    # host time to ENQUEUE one step (no synchronisation inside): when it is below the device time of the step the
    # launches are hidden behind the running kernels and a CUDA graph has nothing to recover
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(3):
        w.step(2 * 10 ** 6 + it)
    host_ms = (time.perf_counter() - t0) / 3 * 1e3
    torch.cuda.synchronize()
    out['host_enqueue_ms_per_step'] = host_ms
    out['launch_bound'] = bool(host_ms > ms / steps)
    n_e2e = max(3, steps // 2)
    for it in range(2):
        w.api_step(it)
    ms_api, _ = timed(w.api_step, 2, n_e2e, barrier, world, dev)
    for it in range(2):
        w.overlapped_step(it)
    ms_ov, _ = timed(w.overlapped_step, 2, n_e2e, barrier, world, dev)
    out['e2e'] = {'value': w.B * world * n_e2e / (ms_api / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': w.h2d,
                  'd2h_bytes_per_step': w.d2h, 'steps': n_e2e, 'ms_per_step': ms_api / n_e2e,
                  'what': 'host NumPy arrays through the public API, as a script of the reference calls it: ' +
                          (w.api_step.__doc__ or '').strip().split('\n')[0]}
    out['e2e_overlapped'] = {'value': w.B * world * n_e2e / (ms_ov / 1e3), 'unit': UNIT, 'steps': n_e2e,
                             'ms_per_step': ms_ov / n_e2e,
                             'h2d_bytes_per_step': getattr(w, 'h2d_overlapped', w.h2d), 'd2h_bytes_per_step': 4 * 4,
                             'what': (w.overlapped_step.__doc__ or '').strip().split('\n')[0]}
    if profile:
        tot = profile_calls(w.step)
        out['roofline'] = tensor_roofline(tot, mode, ms / steps)
        out['kernel_time_ms_per_step'] = {k: round(v[0], 4) for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:10]}
        hbm, _, _, which = peaks()
        syn = tot.get('gn_synth_f32')
        if syn is not None and syn[1] > 0:
            syn_ms = syn[0] / syn[1]
            nbytes = w.B * (4 * w.N + 4 * w.L)
            gbs = nbytes / (syn_ms / 1e3) / 1e9
            out['roofline_synth'] = {'bound': 'hbm', 'kernel': 'synth_kernel (Philox noise [+ template] -> whiten -> crop)',
                                     'achieved': gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': gbs / hbm, 'traffic': None,
                                     'peak_source': which, 'avg_launch_ms': syn_ms, 'algorithmic_bytes_per_launch': nbytes}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from gennet_b200 import nn, parallel
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        parallel.init_data_parallel('nccl')
    nn.set_seed(1)
    nn.set_compute_dtype(MODES[args.mode])
    synth_obj, templates, labels = make_inputs(7, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # BASELINE's second metric, timed alone before the training loops start (and again after them)
    whiten_alone = whiten_roofline(synth_obj, dev, sample_clocks=local if rank == 0 else None)
    classes = {'gan': GANWorkload, 'pe': PEWorkload}
    head = classes[args.config](dev, rank, world, synth_obj, templates, labels)
    m = measure(head, args.steps, args.warmup, barrier, world, dev, rank, local, mode=args.mode)
    prod = PLANE_PRODUCTS.get(args.mode, 0)
    out = {'metric': METRICS[args.config], 'value': m['value'], 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
           'warmup': args.warmup, 'ms_per_step': m['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': DTYPES[args.mode], 'data': 'synthetic',
           'config': {'workload': head.describe(), 'name': args.config, 'batch_per_gpu': head.B,
                      'global_batch': head.B * world, 'n_pix': FS, 'fft_len': FS * T_OBS, 'parallelism': 'dp%d' % world,
                      'precision': {'f16x2': 'float32 activations / weights / gradients / Adam; Conv1D, Conv2D and Dense operands as scaled '
                                             'fp16 pairs (two fp16 planes of the tensor times a power of two from its max |x|: 22-23 bits '
                                             'relative to the tensor scale), three tcgen05.mma per K step into two fp32 TMEM accumulators '
                                             '(float32-class accuracy: tests/test_gpu_conv_f16x2.py, tests/test_gpu_models.py at rtol 1e-4)',
                                    'bf16x3': 'float32 activations / weights / gradients / Adam; Conv1D and Conv2D operands split into three '
                                              'bf16 planes, six tcgen05.mma per K step into two fp32 TMEM accumulators (float32-class '
                                              'accuracy: tests/test_gpu_conv_tc3.py, tests/test_gpu_models.py at rtol 1e-4)',
                                    'bf16x2': 'as bf16x3 with two planes / three products (~2^-16 relative)',
                                    'bf16': 'bf16 activations / conv operands, fp32 accumulation, fp32 master weights and Adam',
                                    'fp32': 'fp32 SIMT'}[args.mode],
                      'l2': 'no flush: per-step working set (several GB of activations) >> 126 MB L2'},
           'same_precision': args.mode != 'bf16',
           'parity': {'synthesis': 'pinned (reference source executed, tests/golden/synth_ref.npz)',
                      'network': 'unpinned (Keras/TensorFlow absent; float64 restatement, rtol 1e-4 in this mode)'},
           'e2e': m['e2e'], 'e2e_overlapped': m['e2e_overlapped'], 'gpu_launches': m['gpu_launches'], 'clocks': m['clocks']}
    for k in ('roofline', 'roofline_synth', 'kernel_time_ms_per_step', 'host_enqueue_ms_per_step', 'launch_bound'):
        if k in m:
            out[k] = m[k]
    extra = {}
    other = 'pe' if args.config == 'gan' else 'gan'
    if not args.no_extra:
        del head
        torch.cuda.empty_cache()
        ow = classes[other](dev, rank, world, synth_obj, templates, labels)
        om = measure(ow, max(5, args.steps // 2), 3, barrier, world, dev, rank, local, mode=args.mode)
        om.update({'metric': METRICS[other], 'unit': UNIT, 'workload': ow.describe(), 'batch_per_gpu': ow.B})
        extra[other] = om
        del ow
        torch.cuda.empty_cache()
        if world == 1 and args.mode == 'f16x2':
            # the other tensor-core modes on the same two workloads, clearly labelled
            for mname, key, note in (('bf16x3', 'bf16x3_mode', dict(same_precision=True, dtype=DTYPES['bf16x3'],
                                      tolerance='whole-step parity suite at rtol 1e-4 (tests/test_gpu_models.py); operands carry '
                                                '24 mantissa bits element-wise, six MMAs per product')),
                                     ('bf16x2', 'bf16x2_mode', dict(same_precision=True, dtype=DTYPES['bf16x2'],
                                      tolerance='whole-step parity suite at rtol 1e-4 (tests/test_gpu_models.py::test_step_parity_bf16x2); '
                                                'operands carry 16 mantissa bits')),
                                     ('bf16', 'bf16_throughput_mode', dict(same_precision=False, dtype='bf16',
                                      tolerance='outputs 2e-2, losses 3e-2, gradients 1.5e-1 relative L2 (tests/test_gpu_models.py)'))):
                nn.set_compute_dtype(MODES[mname])
                tm = {}
                for name in ('gan', 'pe'):
                    bw = classes[name](dev, rank, world, synth_obj, templates, labels)
                    bm = measure(bw, 5, 3, barrier, world, dev, rank, local, profile=True, mode=mname)
                    tm[name] = {k: bm[k] for k in ('value', 'ms_per_step', 'gpu_launches') if k in bm}
                    tm[name]['roofline'] = bm.get('roofline')
                    del bw
                    torch.cuda.empty_cache()
                extra[key] = dict(tm, **note)
            nn.set_compute_dtype(MODES[args.mode])
    after = whiten_roofline(synth_obj, dev, sample_clocks=local if rank == 0 else None)
    whiten_alone['after_training_loop'] = {k: after[k] for k in ('achieved', 'frac', 'avg_launch_ms', 'clocks', 'batch')}
    out['roofline_whiten'] = whiten_alone
    out['extra'] = extra
    if rank == 0:
        if world == 1:
            out['cpu_baseline'] = cpu_baseline(args.config)
            out['cpu_baseline']['same_config'] = 'same networks, n_pix and calls; bounded sample batch (see sample)'
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        parallel.shutdown()


# ---- the reference's CPU path restated (oracle port): the checker timed as the baseline ------------------------------
CPU_SAMPLE_BATCH = {'gan': 32, 'pe': 64}      # bounded sample: one full-batch step is 1-2 minutes of CPU work


def cpu_step_fn(config):
    """(step function, batch): NumPy float64 gen_noise + whiten_data per sample (gw_template_maker.py:161-193,243-286)
    and the torch-CPU float32 Keras-semantics training step of the same workload on all host cores, on a bounded
    sample batch (same per-sample work: same n_pix, same networks, same three calls per iteration)."""
    import torch
    from oracle import keras_oracle as ko, synth_oracle as so
    torch.set_num_threads(os.cpu_count() or 1)
    psd = so.analytic_psd(FS, T_OBS)
    rs = np.random.RandomState(0)
    hp, _ = so.newtonian_chirp_fd(36.0, 29.0, FS, T_OBS)
    templ = np.roll(np.fft.irfft(hp, FS * T_OBS) * FS, -FS)
    nf = FS * T_OBS // 2 + 1

    def synth(n, with_template):
        return np.stack([so.synth_sample(templ if with_template else np.zeros_like(templ),
                                         np.stack([rs.normal(0, 1, nf), rs.normal(0, 1, nf)]), FS, T_OBS, psd)
                         for _ in range(n)])
    ko.clear_session()
    if config == 'pe':
        B = CPU_SAMPLE_BATCH['pe']
        model = ko.build(ko.bbh_signal_pe_model(FS), seed=1, dtype=torch.float32)
        model.compile('mean_squared_error', ko.Adam(9e-5, beta_1=0.5))
        y = [rs.uniform(0.5, 1, B).astype(np.float32), rs.uniform(0.5, 1, B).astype(np.float32)]

        def step():
            xs = synth(B, True)
            return model.train_on_batch(xs[:, :, None].astype(np.float32), y)
        return step, B
    B = CPU_SAMPLE_BATCH['gan']
    noise_signal = rs.normal(size=(FS, 1))
    og = ko.build(ko.bbh_generator_model(FS), seed=1, dtype=torch.float32)
    od = ko.build(ko.bbh_signal_discriminator_model(FS), seed=2, dtype=torch.float32)
    osub = ko.Sequential([ko.StackResidual(noise_signal.astype(np.float32))])
    ocomp = ko.Sequential([ko.Sequential([og, osub]), od])
    ocomp.build((100,), None, torch.float32)
    ko.set_trainable(od, False)
    ocomp.compile('binary_crossentropy', ko.Adam(9e-5, beta_1=0.5))
    ko.set_trainable(od, True)
    od.compile('binary_crossentropy', ko.Adam(9e-5, beta_1=0.5))
    bank = synth(8, True)[:, :, None].astype(np.float32)

    def step():
        signal = bank[rs.randint(0, bank.shape[0], B)]
        noise_ch = synth(B, False)[:, :, None].astype(np.float32)          # whitened PSD-coloured noise channel
        z = rs.uniform(-1, 1, (B, 100)).astype(np.float32)
        gen = og.predict(z)
        fake = np.concatenate((gen, noise_signal[None].astype(np.float32) - gen), axis=2)
        sX = np.concatenate((np.concatenate((signal, noise_ch), axis=2), fake)).reshape(2 * B, FS, 2, 1)
        sd = od.train_on_batch(sX, np.array([1.0] * B + [0.0] * B, dtype=np.float32))
        sg = ocomp.train_on_batch(rs.uniform(-1, 1, (B, 100)).astype(np.float32), np.ones(B, dtype=np.float32))
        return sd + sg
    return step, B


def cpu_baseline(config, steps=1, warmup=0):
    import torch
    step, B = cpu_step_fn(config)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {'value': B * steps / dt, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': '%d step(s) of the same workload on a bounded sample batch of %d (the GPU arm runs %d per GPU; per-sample work '
                      'is identical): NumPy f64 synthesis per sample + torch-CPU f32 network with Keras semantics (oracle port; '
                      'TF-1.12-CPU is not installable here)' % (steps, B, {'gan': GAN_BATCH, 'pe': PE_BATCH}[config]),
            'sample_batch': B, 'seconds': dt}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    step, B = cpu_step_fn(args.config)
    warm = min(args.warmup, 1)
    steps = max(1, min(args.steps, 2))
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    v = B * steps / dt
    w = 'reference CPU path restated (oracle port), same workload as the `ours` arm: ' + (
        'bbhMahoGANy.py GAN iteration (predict + D step + G step), n_pix 2048, batch %d' % B if args.config == 'gan'
        else 'signal_pe_model train step, n_pix 2048, batch %d' % B)
    cb = {'value': v, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port', 'sample_batch': B,
          'same_config': 'same networks, n_pix and calls; bounded sample batch',
          'sample': '%d step(s) of a bounded sample batch of %d after %d warm-up (the GPU arm runs %d per GPU; a full-batch '
                    'step is 1-2 minutes of CPU work)' % (steps, B, warm, {'gan': GAN_BATCH, 'pe': PE_BATCH}[args.config])}
    print(json.dumps({'impl': 'reference', 'metric': METRICS[args.config], 'value': v, 'unit': UNIT, 'n_gpus': args.gpus,
                      'steps': steps, 'warmup': warm, 'ms_per_step': dt / steps * 1e3, 'higher_is_better': True,
                      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 network / f64 synthesis', 'data': 'synthetic',
                      'config': {'workload': w, 'name': args.config, 'batch_per_gpu': B}, 'cpu_baseline': cb,
                      'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def run_check(args):
    """`--gpus 2 --check` (torchrun): the data-parallel GAN iteration on N ranks x B/N samples (NCCL gradient and SyncBN
    all-reduces) reproduces the single-GPU global-batch iteration: every rank runs both and rank 0 reports."""
    import torch
    import torch.distributed as dist
    from gennet_b200 import nn, bbh, parallel
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    nn.set_compute_dtype(MODES[args.mode])
    L, B = 512, 16
    rs = np.random.RandomState(0)
    noise_signal = rs.normal(size=(L, 1)).astype(np.float32)
    z = rs.uniform(-1, 1, (B, 100)).astype(np.float32)
    sX = rs.normal(size=(2 * B, L, 2, 1)).astype(np.float32)
    sy = np.array([1.0] * B + [0.0] * B, dtype=np.float32)

    def run(dp):
        nn.clear_session()
        nn.set_seed(1)
        nn._STATE['dp'] = dp
        bbh.n_pix = L
        G, D, DG, _ = bbh.build_gan(noise_signal)
        for l in G.all_layers() + D.all_layers():
            if type(l).__name__ == 'Dropout':
                l.rate = 0.0                      # the comparison must not depend on the ranks' Philox offsets
        if dp is None:
            rd = D.train_on_batch(sX, sy)
            rg = DG.train_on_batch(z, [1] * B)
        else:
            sl = np.arange(B)[rank * B // world:(rank + 1) * B // world]
            rd = D.train_on_batch(sX[np.r_[sl, B + sl]], sy[np.r_[sl, B + sl]])
            rg = DG.train_on_batch(z[sl], [1] * len(sl))
        w = np.concatenate([a.ravel() for a in G.get_weights() + D.get_weights()])
        return rd + rg, w
    dp = parallel.init_data_parallel('nccl') if world > 1 else None
    r_dp, w_dp = run(dp)
    r_1, w_1 = run(None)
    rel = float(np.linalg.norm(w_dp - w_1) / np.linalg.norm(w_1))
    dl = float(np.abs(np.array(r_dp) - np.array(r_1)).max())
    ok = bool(rel < 1e-5 and dl < 1e-4)
    if rank == 0:
        print(json.dumps({'check': 'data-parallel GAN iteration (NCCL gradient + SyncBN all-reduce) vs single-GPU global batch',
                          'n_gpus': world, 'mode': args.mode, 'losses_dp': r_dp, 'losses_single': r_1, 'max_loss_diff': dl,
                          'updated_weights_rel_l2': rel, 'ok': ok}))
    if world > 1:
        dist.barrier()
        parallel.shutdown()
    if not ok:
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='gan', choices=['gan', 'pe'],
                    help='headline workload (the other one is measured too and reported under extra)')
    ap.add_argument('--mode', default='f16x2', choices=sorted(MODES),
                    help='f16x2: float32-accuracy split-operand tensor-core mode with scaled fp16 pairs (default, the parity '
                         'mode); bf16x3: the same with three bf16 planes; '
                         'bf16: throughput mode; fp32: SIMT')
    ap.add_argument('--no-extra', action='store_true', help='skip the secondary workload and the bf16 comparison lines')
    ap.add_argument('--check', action='store_true', help='data-parallel equivalence check (use with torchrun, N >= 2)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    elif args.check:
        run_check(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
