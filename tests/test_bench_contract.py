"""The bench line committed from the last GPU run of the round carries every key of the measurement contract (this is
a schema check of the newest profiles/r*_bench_final.json, not a performance assertion), and bench.py parses its flags."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line():
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r[0-9][0-9]_bench_final.json'))) or \
        sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r[0-9][0-9]_bench_bf16_final.json')))
    assert files, 'no committed bench line'
    return json.loads(open(files[-1]).read().strip().splitlines()[-1])


def test_bench_line_has_contract_keys():
    d = _line()
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline', 'cpu_baseline'):
        assert k in d, k
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert 'workload' in d['config'] and 'l2' in d['config']
    assert d['steps'] >= 1 and d['warmup'] >= 3 and d['gpu_launches'] > 0
    for k in ('value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'):
        assert k in d['e2e'], k
    assert d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['d2h_bytes_per_step'] > 0
    for name in ('roofline', 'roofline_whiten'):
        r = d[name]
        for k in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'):
            assert k in r, (name, k)
        assert r['bound'] in ('hbm', 'tensor') and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    for k in ('value', 'unit', 'cores', 'kind', 'sample'):
        assert k in d['cpu_baseline'], k
    assert d['cpu_baseline']['kind'] in ('port', 'reference')
    for k in ('sm_mhz', 'sm_max_mhz', 'reasons'):
        assert k in d['clocks'], k
    assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}


def test_bench_cli_flags():
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(ROOT, 'bench.py'))
    src = open(os.path.join(ROOT, 'bench.py')).read()
    for flag in ('--gpus', '--steps', '--warmup', '--impl', '--config', '--mode', '--check'):
        assert flag in src
    assert spec is not None
