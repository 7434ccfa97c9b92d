"""Generate golden vectors FROM THE REFERENCE'S OWN SOURCE (run in the build container only).

The reference is Python-2 script code that cannot be imported (print statements,
cPickle, LALSuite/Keras imports at module top).  Its pure-NumPy functions on the
hot path are however valid Python 3 once isolated, so this script slices their
source text out of the read-only checkout, executes it in a namespace that only
holds NumPy, and records inputs and outputs.  Nothing from the reference is
copied into the repo: only the numeric vectors (``synth_ref.npz``) are committed.

    python tests/golden/make_golden.py            # rewrites tests/golden/synth_ref.npz

Reference functions exercised (file:line):
  BBH_version/gw_template_maker.py: tukey 87-113, convert_beta 133-159,
  gen_noise 161-193, whiten_data 243-286, gen_masses 289-370, gen_par 372-460
  tests/burstMahoGANy.py: make_burst_waveforms 76-98
  train_on_wvf_version/nn.py: sample_data 58-70
"""
import os
import re
import random
import sys

import numpy as np

REF = os.environ.get('GENNET_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))


def slice_defs(path, names):
    """Return source text of top-level ``def``/``class`` blocks called ``names``."""
    lines = open(path).read().expandtabs(8).split('\n')
    out = []
    i = 0
    while i < len(lines):
        m = re.match(r'(def|class)\s+(\w+)', lines[i])
        if m and m.group(2) in names:
            j = i + 1
            while j < len(lines) and (lines[j].strip() == '' or lines[j][0] in ' #' or lines[j].startswith('"""')):
                j += 1
            block = lines[i:j]
            # py2 print statements only occur as diagnostics; neutralise them
            block = [re.sub(r"print\s+'.*$", 'pass', l) for l in block]
            out.append('\n'.join(block))
            i = j
        else:
            i += 1
    return '\n\n'.join(out)


def load_reference():
    ns = {'np': np, 'random': random, 'safe': 2, 'verb': False, 'time': __import__('time'),
          'exit': sys.exit}
    src = slice_defs(os.path.join(REF, 'BBH_version/gw_template_maker.py'),
                     {'bbhparams', 'tukey', 'convert_beta', 'gen_noise', 'whiten_data', 'gen_masses', 'gen_par'})
    exec(compile(src, 'gw_template_maker_slice', 'exec'), ns)
    ns_b = {'np': np, 'random': random}
    exec(compile(slice_defs(os.path.join(REF, 'tests/burstMahoGANy.py'), {'make_burst_waveforms'}),
                 'burst_slice', 'exec'), ns_b)
    ns_n = {'np': np}
    exec(compile(slice_defs(os.path.join(REF, 'train_on_wvf_version/nn.py'), {'sample_data'}),
                 'nn_slice', 'exec'), ns_n)
    return ns, ns_b, ns_n


def toy_psd(fs, T_obs, f_low=10.0):
    N = int(T_obs * fs)
    f = np.arange(N // 2 + 1) / float(T_obs)
    x = np.maximum(f, 1e-3) / 215.0
    s = 1e-49 * (x ** -4.14 - 5.0 * x ** -2 + 111.0 * (1 - x ** 2 + 0.5 * x ** 4) / (1 + 0.5 * x ** 2))
    s[f < f_low] = 0.0
    return s


def main():
    ns, ns_b, ns_n = load_reference()
    g = {}
    # a1: tukey at the sizes the path uses (+ gen_bbh's aggressive window length)
    for M, a in [(16, 0.5), (4096, 0.125), (8192, 0.125), (2184, 0.125), (4369, 0.125), (33, 0.3)]:
        g['tukey_%d_%g' % (M, a)] = ns['tukey'](M, alpha=a)
    # a6: convert_beta
    g['convert_beta'] = np.array([ns['convert_beta'](b, fs, T) for b, fs, T in
                                  [([0.75, 0.95], 1024, 4), ([0.45, 0.55], 1024, 4), ([0.5, 0.5], 2048, 4),
                                   ([0.45, 0.55], 4096, 8)]])
    # a2/a3: gen_noise and whiten_data for fs=1024 (code default) and 2048 (config 2)
    for fs in (1024, 2048):
        T = 4
        psd = toy_psd(fs, T)
        np.random.seed(1234 + fs)
        x = ns['gen_noise'](fs, T, psd.copy())
        g['noise_td_%d' % fs] = x
        g['whiten_td_%d' % fs] = ns['whiten_data'](x.copy(), T, fs, psd.copy(), 'td')
        rs = np.random.RandomState(77 + fs)
        xf = (rs.normal(size=fs * T // 2 + 1) + 1j * rs.normal(size=fs * T // 2 + 1)) * 1e-23
        g['fd_in_%d' % fs] = xf.copy()
        g['whiten_fd_%d' % fs] = ns['whiten_data'](xf.copy(), T, fs, psd.copy(), 'fd')
    # a6: gen_masses / gen_par (RNG call order)
    for k, mdist in enumerate(['astro', 'hunt_constrain', 'gh', 'metric']):
        np.random.seed(50 + k)
        rows = []
        for _ in range(8):
            m12, mc, eta = ns['gen_masses'](5.0, 100.0, mdist)
            rows.append([float(np.ravel(m12)[0]), float(np.ravel(m12)[1]), float(np.ravel(mc)[0]), float(np.ravel(eta)[0])])
        g['gen_masses_' + mdist] = np.array(rows)
    np.random.seed(99)
    rows = []
    for _ in range(6):
        p = ns['gen_par'](1024, 4, mdist='hunt_constrain', beta=[0.45, 0.55], gw_tmp=False)
        rows.append([p.mc, p.M, p.eta, p.m1, p.m2, p.idx])
    p = ns['gen_par'](1024, 4, mdist='hunt_constrain', beta=[0.45, 0.55], gw_tmp=True)
    rows.append([p.mc, p.M, p.eta, p.m1, p.m2, p.idx])
    g['gen_par'] = np.array(rows, dtype=np.float64)
    # a10: burst waveforms and sinusoids
    random.seed(5)
    d, pr = ns_b['make_burst_waveforms'](6, rand5=True)
    g['burst_data'], g['burst_pars'] = d, pr
    d1, p1 = ns_b['make_burst_waveforms'](1)
    g['burst_fixed'] = d1
    np.random.seed(8)
    g['nn_sample_data'] = ns_n['sample_data'](n_samples=5)
    np.savez_compressed(os.path.join(HERE, 'synth_ref.npz'), **g)
    print('wrote', os.path.join(HERE, 'synth_ref.npz'), sorted(g))


if __name__ == '__main__':
    main()
