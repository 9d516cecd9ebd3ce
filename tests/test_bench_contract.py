"""The bench line committed under profiles/ (the output of `python bench.py` on a B200) carries every key of the driver's contract,
and bench.py's reference arm / argument defaults stay as the contract states them.  CPU-only: parses files, runs nothing on a GPU."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, 'profiles', name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_own_arm_line_has_the_contract_keys():
    d = _line('r1_bench_1gpu.json')
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'clocks', 'e2e', 'gpu_launches', 'roofline', 'cpu_baseline'):
        assert k in d, k
    assert d['n_gpus'] == 1 and d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['unit'] == 'Gevents/s' and d['dtype'] == 'f64' and d['data'] == 'synthetic'
    assert 'workload' in d['config'] and 'model' not in d['config'] and 'l2' in d['config']
    assert d['warmup'] >= 3 and d['gpu_launches'] > 0
    for k in ('value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'):
        assert k in d['e2e'], k
    assert d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['d2h_bytes_per_step'] > 0
    assert d['e2e']['value'] != d['value']                                   # measured separately, not a copy of the device number
    r = d['roofline']
    for k in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'):
        assert k in r, k
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s' and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-12
    c = d['cpu_baseline']
    for k in ('value', 'unit', 'cores', 'kind', 'sample'):
        assert k in c, k
    assert c['kind'] in ('port', 'reference') and c['cores'] >= 1
    for k in ('sm_mhz', 'sm_max_mhz', 'reasons'):
        assert k in d['clocks'], k
    assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    # value = units all ranks processed / time: events per step / ms_per_step
    ev = d['config']['events_per_step_per_gpu'] * d['n_gpus']
    assert abs(d['value'] - ev / (d['ms_per_step'] * 1e-3) / 1e9) < 1e-6 * d['value']


def test_reference_arm_line():
    d = _line('r1_reference_arm_1gpu.json')
    assert d['impl'] == 'reference' and d['unit'] == 'Gevents/s' and d['higher_is_better'] is True
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0 and d['e2e']['value'] == d['value']
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['value'] == d['value']


def test_bench_defaults_and_oracle_use():
    src = open(os.path.join(ROOT, 'bench.py')).read()
    assert re.search(r"'--gpus', type=int, default=1", src) and re.search(r"'--impl', default='own'", src)
    m = re.search(r"'--warmup', type=int, default=(\d+)", src)
    assert m and int(m.group(1)) >= 3
    # the oracle is only reached from the CPU-baseline / reference legs
    for m in re.finditer(r'^(\s*)from oracle|^(\s*)import oracle', src, flags=re.M):
        assert len(m.group(1) or m.group(2) or '') > 0, 'oracle imported at module level in bench.py'
