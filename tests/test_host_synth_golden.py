"""The PRODUCT's host functions of the synthesis stage (gennet_b200.synth.tukey / convert_beta / gen_masses / gen_par,
the ones a script of the reference calls) against the vectors produced by the reference's own source
(tests/golden/make_golden.py -> synth_ref.npz): same windows bit for bit, same RNG call order, same parameters.
No kernel is launched here: these functions are float64 host code."""
import numpy as np
import pytest


@pytest.fixture(scope='module')
def synth():
    from gennet_b200 import synth as s      # loads the C-ABI library (no device needed for these host functions)
    return s


def test_product_tukey_matches_reference(golden, synth):
    keys = [k for k in golden.files if k.startswith('tukey_')]
    assert keys
    for key in keys:
        _, M, a = key.split('_')
        assert np.array_equal(synth.tukey(int(M), float(a)), golden[key]), key


def test_product_convert_beta_matches_reference(golden, synth):
    cases = [([0.75, 0.95], 1024, 4), ([0.45, 0.55], 1024, 4), ([0.5, 0.5], 2048, 4), ([0.45, 0.55], 4096, 8)]
    assert np.array_equal(np.array([synth.convert_beta(*c) for c in cases]), golden['convert_beta'])


def test_product_gen_masses_and_gen_par_follow_reference_rng_order(golden, synth):
    for k, mdist in enumerate(['astro', 'hunt_constrain', 'gh', 'metric']):
        rs = np.random.RandomState(50 + k)
        rows = []
        for _ in range(8):
            m12, mc, eta = synth.gen_masses(5.0, 100.0, mdist, rng=rs)
            rows.append([m12[0], m12[1], mc, eta])
        np.testing.assert_allclose(np.array(rows, dtype=np.float64), golden['gen_masses_' + mdist], rtol=1e-15)
    rs = np.random.RandomState(99)
    rows = []
    for _ in range(6):
        p = synth.gen_par(1024, 4, mdist='hunt_constrain', beta=[0.45, 0.55], rng=rs)
        rows.append([p.mc, p.M, p.eta, p.m1, p.m2, p.idx])
    p = synth.gen_par(1024, 4, mdist='hunt_constrain', beta=[0.45, 0.55], gw_tmp=True, rng=rs)
    rows.append([p.mc, p.M, p.eta, p.m1, p.m2, p.idx])
    np.testing.assert_allclose(np.array(rows, dtype=np.float64), golden['gen_par'], rtol=1e-15)


def test_product_gen_masses_unknown_distribution_exits(synth):
    with pytest.raises(SystemExit):           # the reference prints and exit(1)s (gw_template_maker.py:369-370)
        synth.gen_masses(5.0, 100.0, 'nope', rng=np.random.RandomState(0))
