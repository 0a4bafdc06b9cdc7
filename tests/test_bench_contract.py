"""
The reference arm of bench.py (`--impl reference`) runs on the CPU alone, so its side of the
JSON contract is checked here: one line, the keys the driver reads, the CPU arm on the GPU arm's
configuration (same_config), and non-zero ranks leaving without work.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                           '--steps', '1', '--warmup', '0'],
                          capture_output=True, text=True, timeout=900, env=env)


def test_reference_arm_line():
    o = _run({'OMP_NUM_THREADS': '1'})          # as torchrun would set it: the leg overrides it
    assert o.returncode == 0, o.stderr[-500:]
    lines = [l for l in o.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    r = json.loads(lines[0])
    assert r['impl'] == 'reference' and r['higher_is_better'] is True
    assert r['metric'] == 'implicit TS throughput (ROSW steps x grid points)'
    assert r['unit'] == 'Mpts*steps/s' and r['value'] > 0 and r['steps'] == 1
    assert r['dtype'] == 'f64' and r['vs_baseline'] is None
    assert r['e2e'] == dict(value=r['value'], unit=r['unit'], h2d_bytes_per_step=0,
                            d2h_bytes_per_step=0)
    cb = r['cpu_baseline']
    assert cb['kind'] == 'port' and cb['value'] == r['value'] and cb['cores'] >= 1
    assert 'FULL 1024x1024' in cb['sample']
    assert r['same_config'] is True
    ops = r['cpu_operator_1024x1024']
    assert ops['residual']['mpts_per_s'] > 0 and ops['c_all_threads']['jvp']['mpts_per_s'] > 0
    try:
        nthreads = len(os.sched_getaffinity(0))
    except AttributeError:
        nthreads = os.cpu_count()
    assert cb['cores'] == nthreads


def test_reference_arm_other_ranks_exit_quietly():
    o = _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert o.returncode == 0 and o.stdout.strip() == ''
