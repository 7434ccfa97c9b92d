#!/usr/bin/env python
"""Benchmark of the GenNet hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode fp32|bf16]

Workload = BASELINE.json configs[1]: bbhMahoGANy.py CNN point estimator on 1 s @ 2048 Hz whitened synthetic
BBH chirps, batch 512 per GPU.  One step = per-batch sample synthesis on the device (Philox PSD-coloured
noise -> irfft -> + template -> Tukey window -> rfft -> whitening -> irfft -> crop, N=8192 -> L=2048) fused
with signal_pe.train_on_batch (forward, backward, Keras Adam).  Metric: train samples/sec.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS, T_OBS, BATCH, N_TEMPLATES = 2048, 4, 512, 1024
METRIC, UNIT = 'train samples/sec (synth+whiten+CNN-PE step)', 'samples/s'


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


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel family from the latest committed `ncu --set full` capture
    (profiles/rNN_conv_tc_step_traffic.json, written by profiles/summarize_step_traffic.py); None if absent."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r[0-9][0-9]_conv_tc_step_traffic.json')))
    if not files:
        return None, None
    try:
        with open(files[-1]) as f:
            d = json.load(f)
        return d['dram_bytes_per_launch'], os.path.basename(files[-1])
    except Exception:
        return None, None


def make_inputs(seed, device):
    """Synthetic chirp bank of BASELINE config 2: unwhitened time-domain templates (n,N) f32 in HBM,
    labels (mc/35, q), analytic PSD."""
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


def step_flops():
    """Algorithmic conv/dense flops of one PE training step per sample at L=2048 (fwd + dgrad + wgrad;
    no dgrad for the two first layers): SURVEY 8d, 2*L_out*k*Cin*Cout per layer."""
    L = FS
    tot = 0.0
    # (Lin, Cin, Cout, stride, same, first)
    mc = [(L, 1, 64, 2, True, True), (None, 64, 128, 2, False, False), (None, 128, 256, 2, False, False),
          (None, 256, 512, 2, False, False)]
    q = [(L, 1, 64, 1, True, True), (None, 64, 128, 1, False, False), (None, 128, 256, 1, False, False),
         (None, 256, 512, 2, False, False), (None, 512, 1024, 2, False, False)]
    for tower in (mc, q):
        cur = L
        for Lin, cin, cout, s, same, first in tower:
            lo = -(-cur // s) if same else (cur - 5) // s + 1
            f = 2.0 * lo * 5 * cin * cout
            tot += f * (2 if first else 3)
            cur = lo
        tot += 2.0 * cur * tower[-1][2] * 3          # Dense(1) head
    return tot


def run_ours(args):
    import torch
    import torch.distributed as dist
    from gennet_b200 import nn, bbh, parallel, _lib
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dp = parallel.init_data_parallel('nccl') if world > 1 else None
    nn.set_seed(1)
    nn.set_compute_dtype('bfloat16' if args.mode == 'bf16' else 'float32')
    bbh.n_pix = FS
    synth_obj, templates, labels = make_inputs(7, dev)
    pe = bbh.signal_pe_model()
    pe.compile(loss='mean_squared_error', optimizer=nn.Adam(lr=9e-5, beta_1=0.5), metrics=['accuracy'])
    if dp is not None:
        parallel.broadcast_weights(pe)
    B, L, N = BATCH, FS, FS * T_OBS
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    batch = torch.empty((B, L), dtype=torch.float32, device=dev)
    tgt = torch.empty((B, 2), dtype=torch.float32, device=dev)

    def one_step(it):
        # data-parallel: rank r synthesises global samples [it*B*world + r*B, ...): the Philox stream
        # depends on the global sample index only
        idx = torch.randint(0, N_TEMPLATES, (B,), device=dev, generator=gen, dtype=torch.int32)
        synth_obj.synth(B, templates=templates, tidx=idx, scale=1.0, seed=2024,
                        sample_offset=(it * world + rank) * B, out=batch)
        _lib.call('gn_gather_rows_f32', _lib.ptr(labels), _lib.ptr(idx, torch.int32), _lib.ptr(tgt), B, 2, _lib.stream())
        return pe.train_on_batch(batch.reshape(B, L, 1), [tgt[:, 0].contiguous(), tgt[:, 1].contiguous()],
                                 _return_device=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # BASELINE's second metric, timed alone before the training loop starts (and again after it, see below)
    whiten_alone = whiten_roofline(synth_obj, dev, sample_clocks=local if rank == 0 else None)
    for it in range(args.warmup):
        one_step(it)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    calls0 = _lib.COUNTS['calls']
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for it in range(args.steps):
        res = one_step(args.warmup + it)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = _lib.COUNTS['calls'] - calls0
    last = res.detach().cpu().numpy()
    assert np.isfinite(last).all(), 'training diverged'
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world * args.steps / (ms / 1e3)

    # ---- end-to-end arm: host buffers through the public API ------------------------------------------
    # host strain segments (B,N) f32 pinned -> H2D -> whiten_td+crop on device -> train_on_batch -> D2H loss
    rs = np.random.RandomState(5 + rank)
    host_x = torch.empty((2, B, N), dtype=torch.float32).pin_memory()
    host_x.copy_(torch.as_tensor((rs.normal(size=(2, B, N)) * 1e-21).astype(np.float32)))
    host_y = [rs.uniform(0.5, 1.0, (B,)).astype(np.float32), rs.uniform(0.5, 1.0, (B,)).astype(np.float32)]
    ys_pinned = [torch.as_tensor(v).pin_memory() for v in host_y]

    # a loader thread's job, done here with a copy stream: the H2D copy of batch i+1 (pinned -> device, double
    # buffered) runs while batch i trains; every step still pays its own copy inside the timed region
    copy_stream = torch.cuda.Stream(device=dev)
    dev_x = [torch.empty((B, N), dtype=torch.float32, device=dev) for _ in range(2)]
    dev_y = [[torch.empty((B,), dtype=torch.float32, device=dev) for _ in range(2)] for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def stage(it):
        slot = it % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])          # the step that last used this slot has finished
            dev_x[slot].copy_(host_x[slot], non_blocking=True)
            for k in range(2):
                dev_y[slot][k].copy_(ys_pinned[k], non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_step(it):
        slot = it % 2
        torch.cuda.current_stream().wait_event(ready[slot])
        stage(it + 1)                                        # next batch's copy overlaps this step
        w = synth_obj.whiten_td(dev_x[slot], crop=True, scale=1.0)
        r = pe.train_on_batch(w.reshape(B, L, 1), dev_y[slot])   # returns host floats (D2H read of loss/metric)
        consumed[slot].record()
        return r

    for ev in consumed:
        ev.record()
    stage(0)
    for it in range(2):
        e2e_step(it)
    barrier()
    e0.record()
    n_e2e = max(3, args.steps // 2)
    for it in range(2, 2 + n_e2e):
        r = e2e_step(it)
    e1.record()
    barrier()
    copy_stream.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * world * n_e2e / (float(t.item()) / 1e3)
    h2d = B * N * 4 + 2 * B * 4
    d2h = 4 * 4

    out = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
           'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': 'bf16' if args.mode == 'bf16' else 'f32', 'data': 'synthetic',
           'config': {'workload': 'bbhMahoGANy.py CNN point estimator (signal_pe_model), fs 2048 Hz, N=8192 -> '
                                  'n_pix 2048, batch %d per GPU, synthetic TaylorF2-style chirps + analytic aLIGO-like '
                                  'PSD, random-init weights' % B,
                      'batch_per_gpu': B, 'global_batch': B * world, 'n_pix': L, 'fft_len': N,
                      'parallelism': 'dp%d' % world,
                      'precision': ('bf16 activations / conv operands on tcgen05 tensor cores, fp32 accumulation, fp32 '
                                    'master weights, gradients and Adam' if args.mode == 'bf16' else
                                    'fp32 SIMT (exact-parity path)'),
                      'l2': 'no flush: per-step working set (~4 GB activations) >> 126 MB L2'},
           'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                   'steps': n_e2e, 'what': 'pinned host strain (B,N) f32 -> H2D (copy stream, double buffered, overlapping the previous '
                                           'step) -> whiten_td+crop -> train_on_batch -> D2H [loss, acc]'},
           'gpu_launches': launches, 'clocks': clocks}

    prof = profile_pass(one_step, args, B, L, N)      # every rank: the steps contain collectives
    after = whiten_roofline(synth_obj, dev, sample_clocks=local if rank == 0 else None)
    whiten_alone['after_training_loop'] = {k: after[k] for k in ('achieved', 'frac', 'avg_launch_ms', 'clocks', 'batch')}
    prof['roofline_whiten'] = whiten_alone
    if rank == 0:
        out.update(prof)
        if world == 1:
            out['cpu_baseline'] = cpu_baseline(steps=3, warmup=1)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        parallel.shutdown()


def profile_pass(one_step, args, B, L, N):
    """Per-entry-point device time with CUDA events (on the launching stream) over a few extra steps; gives
    the roofline of the dominant kernel family (Conv1D implicit GEMM) and of the synthesis kernel."""
    import torch
    from gennet_b200 import _lib
    hbm, tf_burst, tf_sust, which = peaks()
    torch.cuda.synchronize()
    _lib.PROFILE = []
    n = 3
    for it in range(n):
        one_step(10 ** 6 + it)
    torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    tot = {}
    for name, tag, a, b in prof:
        d = tot.setdefault(name, [0.0, 0])
        d[0] += a.elapsed_time(b)
        d[1] += 1
    step_ms = sum(v[0] for v in tot.values()) / n
    conv_names = ('gn_conv1d_fwd_f32', 'gn_conv1d_dgrad_f32', 'gn_conv1d_wgrad_f32', 'gn_conv1d_fwd_bf16',
                  'gn_conv1d_dgrad_bf16', 'gn_conv1d_wgrad_bf16', 'gn_conv1d_smallcin_fwd_bf16',
                  'gn_conv1d_smallcin_wgrad_bf16')
    conv_ms = sum(tot[k][0] for k in conv_names if k in tot) / n
    conv_launches = sum(tot[k][1] for k in conv_names if k in tot) / n
    flops = step_flops() * B
    ach = flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    syn = tot.get('gn_synth_f32', [0.0, 1])
    syn_ms = syn[0] / max(syn[1], 1)
    syn_bytes = B * (4 * N + 4 * L)
    syn_gbs = syn_bytes / (syn_ms / 1e3) / 1e9 if syn_ms > 0 else 0.0
    traffic, traffic_src = ncu_traffic()
    return {
        'roofline': {'bound': 'tensor', 'kernel': 'Conv1D implicit GEMM (fwd+dgrad+wgrad launches of one step)',
                     'achieved': ach, 'peak': tf_sust, 'unit': 'TFLOP/s', 'frac': ach / tf_sust,
                     'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': which + ' (sustained bf16 cuBLAS; kernel timed inside a long step)',
                     'avg_launch_ms': conv_ms / max(conv_launches, 1), 'share_of_step': conv_ms / step_ms if step_ms else None,
                     'algorithmic_flops_per_step': flops},
        'roofline_synth': {'bound': 'hbm', 'kernel': 'synth_kernel (Philox noise + inject + whiten + crop)',
                           'achieved': syn_gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': syn_gbs / hbm, 'traffic': None,
                           'peak_source': which, 'avg_launch_ms': syn_ms, 'algorithmic_bytes_per_launch': syn_bytes},
        'kernel_time_ms_per_step': {k: v[0] / n for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:8]},
    }


WHITEN_BATCHES = (8192, 32768)      # series per launch; the last one is the headline (2.1 GB in + out per launch)


def whiten_roofline(synth_obj, dev, batches=WHITEN_BATCHES, iters=20, sample_clocks=None):
    """BASELINE's second metric, "whitening HBM GB/s": gn_whiten_td_f32 alone (window -> rfft -> weights -> irfft) on
    `batch` resident series of N = 8192 samples, timed with CUDA events; algorithmic bytes 8*N per series (read + write
    the series once; window / weights / twiddles are batch-shared).  >= 537 MB per launch >> 126 MB L2.  A launch has
    ~20 us of fixed cost (coefficient prologue, first wave with cold caches and aligned phases, tail of the
    grid: scratch/whiten_batch.py, scratch/whiten_waves.py), so the figure is reported for two batch sizes."""
    import torch
    hbm, _, _, which = peaks()
    N = synth_obj.N
    by_batch, clocks, ms = {}, None, None
    for batch in batches:
        x = torch.randn(batch, N, device=dev) * 1e-21
        for _ in range(3):
            synth_obj.whiten_td(x)
        torch.cuda.synchronize()
        sampler = ClockSampler(sample_clocks, period_s=0.001) if sample_clocks is not None else None
        if sampler is not None:
            sampler.start()
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
    nbytes = batch * 8 * N
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
            'avg_launch_ms': ms, 'algorithmic_bytes_per_launch': nbytes, 'batch': batch, 'iters': iters,
            'clocks': clocks, 'by_batch': by_batch,
            'note': 'on-chip bound: ~400k FP32 lane-ops and ~4000 L1/shared wavefronts per 64 KiB series (DESIGN.md section 6)'}


def cpu_reference_step_fn(sample_batch):
    """The reference's CPU path restated (oracle): NumPy float64 gen_noise + whiten_data per sample
    (gw_template_maker.py:161-193,243-286) and a torch-CPU float32 Keras-semantics PE train step."""
    import torch
    from oracle import keras_oracle as ko, synth_oracle as so
    torch.set_num_threads(os.cpu_count() or 1)
    psd = so.analytic_psd(FS, T_OBS)
    rs = np.random.RandomState(0)
    hp, _ = so.newtonian_chirp_fd(36.0, 29.0, FS, T_OBS)
    templ = np.roll(np.fft.irfft(hp, FS * T_OBS) * FS, -FS)
    model = ko.build(ko.bbh_signal_pe_model(FS), seed=1, dtype=torch.float32)
    model.compile('mean_squared_error', ko.Adam(9e-5, beta_1=0.5))
    y = [rs.uniform(0.5, 1, sample_batch).astype(np.float32), rs.uniform(0.5, 1, sample_batch).astype(np.float32)]

    def step():
        xs = np.stack([so.synth_sample(templ, None if False else np.stack([rs.normal(0, 1, FS * T_OBS // 2 + 1),
                                                                          rs.normal(0, 1, FS * T_OBS // 2 + 1)]),
                                       FS, T_OBS, psd) for _ in range(sample_batch)])
        return model.train_on_batch(xs[:, :, None].astype(np.float32), y)
    return step


def cpu_baseline(steps, warmup, sample_batch=16):
    step = cpu_reference_step_fn(sample_batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    import torch
    return {'value': sample_batch * steps / dt, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': '%d steps of batch %d (same N=8192 -> n_pix 2048 synthesis + PE train step; NumPy f64 synthesis '
                      '+ torch-CPU f32 network, stand-in for TF-1.12-CPU which is not installable)' % (steps, sample_batch)}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample_batch = 16
    step = cpu_reference_step_fn(sample_batch)
    for _ in range(min(args.warmup, 3)):
        step()
    steps = min(args.steps, 10)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    import torch
    v = sample_batch * steps / dt
    cb = {'value': v, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
          'sample': '%d steps of batch %d of the same workload' % (steps, sample_batch)}
    print(json.dumps({'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus,
                      'steps': steps, 'warmup': min(args.warmup, 3), 'ms_per_step': dt / steps * 1e3,
                      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32/f64',
                      'data': 'synthetic',
                      'config': {'workload': 'reference CPU path restated (oracle port): NumPy f64 gen_noise+whiten_data '
                                             'per sample + torch-CPU f32 signal_pe_model train step, n_pix 2048, '
                                             'bounded sample batch %d' % sample_batch},
                      'cpu_baseline': cb,
                      'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mode', default='bf16', choices=['bf16', 'fp32'],
                    help='bf16: tensor-core throughput path (default); fp32: exact-parity SIMT path')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
