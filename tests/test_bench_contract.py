"""bench.py's contract, as far as it can be checked without a GPU: the reference arm
(`--impl reference`: the CPU port of the path on the host cores) prints one JSON line with the
keys the driver reads, ranks other than 0 stay silent, and the clock sampler parses nvidia-smi"""
import json
import os
import subprocess
import sys

from conftest import ROOT


def run(*arguments, **environment):
    return subprocess.run(
        [sys.executable, str(ROOT / 'bench.py'), *arguments], capture_output=True, text=True,
        timeout=600, env={**os.environ, **environment})


def test_reference_arm_prints_the_contract_line():
    result = run('--impl', 'reference', '--steps', '1', '--warmup', '0')
    assert result.returncode == 0, result.stderr
    lines = [line for line in result.stdout.splitlines() if line.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    baseline = json.loads((ROOT / 'BASELINE.json').read_text())
    assert line['impl'] == 'reference' and line['metric'] in baseline['metric']
    assert line['unit'] == 'samples/s' and line['higher_is_better'] is True
    assert (line['n_gpus'], line['steps'], line['warmup']) == (1, 1, 0)
    assert line['scaling'] == 'weak' and line['vs_baseline'] is None and line['data'] == 'synthetic'
    assert 'workload' in line['config'] and 'model' not in line['config']
    assert line['value'] > 0 and line['ms_per_step'] > 0
    cpu = line['cpu_baseline']
    # the unmodified reference (oracle/_ref, vendored by build()) when present, else the oracle port
    kind = 'reference' if (ROOT / 'oracle' / '_ref' / 'promonet').exists() else 'port'
    assert cpu['kind'] == kind and cpu['cores'] == os.cpu_count() and cpu['value'] == line['value']
    assert cpu['unit'] == line['unit'] and 'utterances' in cpu['sample']
    assert line['e2e'] == {
        'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_runs_on_rank_zero_only():
    result = run('--impl', 'reference', '--steps', '1', '--warmup', '0', RANK='1', WORLD_SIZE='2')
    assert result.returncode == 0 and result.stdout.strip() == ''


def test_clock_sampler_parses_nvidia_smi_lines():
    sys.path.insert(0, str(ROOT))
    import bench
    clocks = bench.Clocks(0)
    clocks.lines = [
        '1965, 1965, Not Active, Not Active, Not Active, Active\n',
        '1600, 1965, Not Active, Not Active, Not Active, Not Active\n',
        '1800, 1965, Not Active, Not Active, Not Active, Not Active\n',
        'garbage\n', '[N/A], 1965, Not Active, Not Active, Not Active, Not Active\n']
    assert clocks.summary() == {
        'sm_mhz': 1800., 'sm_max_mhz': 1965., 'reasons': ['sw_power_cap'], 'samples': 3}
    assert bench.Clocks(0).summary()['samples'] == 0
    assert bench.peaks()['hbm_gbs'] > 1000 and bench.peaks()['bf16_tflops_sustained'] > 100
