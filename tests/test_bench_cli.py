"""bench.py contract checks that need no GPU: the reference arm (the unmodified reference
file staged under oracle/_ref -- or the oracle port when it is missing -- timed on the host
cores) standalone and under torchrun with world_size 2 (rank 0 alone prints,
the other rank exits 0 without work), and the product arm failing loudly without CUDA."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
            'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config',
            'cpu_baseline', 'e2e', 'impl')


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _json_lines(text):
    return [json.loads(ln) for ln in text.splitlines() if ln.startswith('{')]


def _check_reference_line(d, n_gpus):
    for k in REQUIRED:
        assert k in d, k
    assert d['impl'] == 'reference' and d['n_gpus'] == n_gpus and d['value'] > 0
    assert d['unit'] == 'pairs/s' and d['higher_is_better'] is True and d['dtype'] == 'f32'
    cb = d['cpu_baseline']
    from oracle import ref_loader
    want = 'reference' if ref_loader.reference_available() else 'port'
    assert cb['kind'] == want and cb['cores'] >= 1 and cb['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'pairs/s', 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_standalone():
    out = subprocess.run([sys.executable, 'bench.py', '--impl', 'reference', '--steps', '1',
                          '--warmup', '0'], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1
    _check_reference_line(lines[0], 1)


def test_reference_arm_under_torchrun_world_size_2():
    env = dict(os.environ)
    env.pop('OMP_NUM_THREADS', None)
    out = subprocess.run(
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
         '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), 'bench.py', '--impl',
         'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'],
        cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1, out.stdout          # rank 0 alone prints
    _check_reference_line(lines[0], 2)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_product_arm_fails_loudly_without_cuda():
    out = subprocess.run([sys.executable, 'bench.py', '--steps', '1', '--warmup', '0', '--no-e2e',
                          '--no-cpu'], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert 'no CUDA device' in (out.stderr + out.stdout)
    assert not _json_lines(out.stdout)          # no number is ever printed from a fallback
