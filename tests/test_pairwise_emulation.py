"""CPU run of the pairwise CUDA kernel's SOURCE under a thread-per-CUDA-thread emulation
(tests/host_math/pairwise_emul.cpp): ``gd_pairwise_kernel`` with one column per lane (``packed``
= 0 below: the round-1 mapping) and with two (``packed`` = 1: the default for m > 32).

What this pins without a GPU: tiling and partial tiles, odd row counts, dead lanes, 1/2
columns per lane and chunked columns, persistent CTAs walking several tiles, the advanced
store pointer, the column-key workspace protocol, NaN-first and lowest-index tie rules -- i.e.
everything around the per-pair arithmetic.  Both mappings run the same per-pair instruction
sequence, so their matrices and minima must agree BIT FOR BIT (here and on the device)."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from mmdet3d_gaussian_b200 import synth
from oracle import gd_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CUDA_INC = '/usr/local/cuda/include'
LOSS = {'gwd3d': 0, 'kld3d': 1, 'bd3d': 5}

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, 'cuda_runtime.h')),
                                reason='needs the CUDA headers (no GPU needed)')


@pytest.fixture(scope='module')
def emul():
    out = os.path.join(tempfile.mkdtemp(prefix='gd_pairwise_emul_'), 'pairwise_emul.so')
    subprocess.run(['g++', '-O1', '-std=c++20', '-ffp-contract=off', '-shared', '-fPIC', '-pthread',
                    '-w', '-I', CUDA_INC, '-I', os.path.join(ROOT, 'include'), '-x', 'c++',
                    os.path.join(HERE, 'host_math', 'pairwise_emul.cpp'), '-o', out], check=True)
    lib = ctypes.CDLL(out)
    lib.gd_emul_pairwise.restype = ctypes.c_int
    lib.gd_emul_pairwise.argtypes = [ctypes.c_int] * 4 + [ctypes.c_uint, ctypes.c_void_p,
                                                          ctypes.c_longlong, ctypes.c_void_p,
                                                          ctypes.c_longlong] + [ctypes.c_void_p] * 5 + [ctypes.c_int]
    return lib


def run(lib, loss, packed, reduce, b1, b2, want_matrix=True, force_cpl=0, cap=3, similarity=0):
    n, m = b1.shape[0], b2.shape[0]
    a1 = np.ascontiguousarray(b1.numpy().astype(np.float32))
    a2 = np.ascontiguousarray(b2.numpy().astype(np.float32))
    # every output sits between two guard bands: a write outside its array is caught
    G = 64
    bufs = {}

    def guarded(name, count, dtype, fill):
        full = np.full(count + 2 * G, fill, dtype)
        full[:G] = full[-G:] = 91
        bufs[name] = full
        return full[G:G + count]
    out = guarded('out', n * m, np.float32, -7.0).reshape(n, m) if want_matrix else None
    rmin, ridx = guarded('rmin', n, np.float32, -7.0), guarded('ridx', n, np.int32, -7)
    cmin, cidx = guarded('cmin', m, np.float32, -7.0), guarded('cidx', m, np.int32, -7)
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p) if x is not None else None   # noqa: E731
    dirty = lib.gd_emul_pairwise(LOSS[loss], packed, reduce, force_cpl, cap, p(a1), n, p(a2), m,
                                 p(out), p(rmin), p(ridx), p(cmin), p(cidx), similarity)
    assert dirty == 0, 'column-key workspace / ticket not restored'
    for name, full in bufs.items():
        assert (full[:G] == 91).all() and (full[-G:] == 91).all(), f'write outside {name}'
    return out, rmin, ridx, cmin, cidx


def first_argmin(mat, axis):
    """(min, lowest index attaining it); NaN counts as the minimum (torch.min semantics)."""
    key = np.where(np.isnan(mat), -np.inf, mat)
    idx = np.argmax(key == key.min(axis=axis, keepdims=True), axis=axis)
    val = np.take_along_axis(mat, np.expand_dims(idx, axis), axis).squeeze(axis)
    return val, idx


def boxes(n, m, degenerate=True):
    b1 = synth.make_anchor_grid(n, 'waymo')
    b2 = synth.make_targets(m, 'waymo', seed=n + m)
    b2[:, 0] = b2[:, 0] * 2.0 - 70.0
    if m > 4:
        b2[m - 1] = b2[1]                      # duplicate GT: ties between columns
    if n > 70:
        b1[n - 1] = b1[3]                      # duplicate anchors: ties between rows,
        b1[68] = b1[3]                         # in different tiles and pair halves
        if degenerate:
            b1[10, 4] = 1e-9                   # rows the FAST cores hand to the robust path
            b1[11, 3] = 2e7
    return b1, b2


def same_bits(a, b):
    return np.array_equal(a.view(np.int32), b.view(np.int32))


SHAPES = [(1, 1), (63, 5), (65, 33), (130, 64), (131, 129), (200, 256), (70, 300)]


@pytest.mark.parametrize('n,m', SHAPES)
def test_scalar_and_packed_matrix(emul, n, m):
    """Matrix mode: both kernels write every (row, column) exactly once, agree bit for bit in
    the host arithmetic and match the fp64 oracle; similarity = 1 - value."""
    b1, b2 = boxes(n, m)
    ref = gd_oracle.pairwise_distance(b1.double(), b2.double(), 'gwd3d', fun='log1p',
                                      tau=1.0).numpy()
    scalar = run(emul, 'gwd3d', 0, 0, b1, b2)[0]
    packed = run(emul, 'gwd3d', 1, 0, b1, b2)[0]
    assert not (scalar == -7.0).any() and not (packed == -7.0).any()      # no hole
    assert np.max(np.abs(scalar - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-5
    assert same_bits(packed, scalar)
    sim = run(emul, 'gwd3d', 1, 0, b1, b2, similarity=1)[0]
    assert same_bits(sim, (np.float32(1.0) - packed).astype(np.float32))


@pytest.mark.parametrize('loss', ['gwd3d', 'kld3d', 'bd3d'])
@pytest.mark.parametrize('n,m', [(65, 33), (131, 129), (200, 256)])
def test_fused_minima_equal_matrix_minima(emul, loss, n, m):
    """The contract of the GPU tests (tests/test_gpu_assign.py), on the CPU: fused row / column
    (min, argmin) == those of the matrix the same launch writes, bit for bit, lowest index on
    ties; the packed kernel agrees with the scalar one; persistent CTAs (cap 3) walk
    several tiles with the column minima kept in registers."""
    b1, b2 = boxes(n, m)
    res = {}
    for packed in (0, 1):
        mat, rmin, ridx, cmin, cidx = run(emul, loss, packed, 1, b1, b2)
        rv, ri = first_argmin(mat, 1)
        cv, ci = first_argmin(mat, 0)
        assert same_bits(rmin, rv) and np.array_equal(ridx, ri), (loss, packed, 'rows')
        assert same_bits(cmin, cv) and np.array_equal(cidx, ci), (loss, packed, 'cols')
        # minima only (no matrix): same answers
        _, rmin2, ridx2, cmin2, cidx2 = run(emul, loss, packed, 1, b1, b2, want_matrix=False)
        assert same_bits(rmin2, rmin) and np.array_equal(ridx2, ridx)
        assert same_bits(cmin2, cmin) and np.array_equal(cidx2, cidx)
        res[packed] = (mat, ridx, cidx)
    assert same_bits(res[0][0], res[1][0])
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])


@pytest.mark.parametrize('n,m,cpl', [(131, 70, 0), (131, 129, 1), (70, 300, 0), (200, 600, 1),
                                     (67, 513, 1), (130, 577, 1)])
def test_packed_chunked_columns(emul, n, m, cpl):
    """More columns than one pass of the CTA covers (m > 256 CPL) and ragged column counts:
    row minima carried across chunks in shared memory, column minima flushed per chunk."""
    b1, b2 = boxes(n, m, degenerate=False)
    mat, rmin, ridx, cmin, cidx = run(emul, 'gwd3d', cpl, 1, b1, b2, cap=2)
    rv, ri = first_argmin(mat, 1)
    cv, ci = first_argmin(mat, 0)
    assert not (mat == -7.0).any()
    assert same_bits(rmin, rv) and np.array_equal(ridx, ri)
    assert same_bits(cmin, cv) and np.array_equal(cidx, ci)
    only = run(emul, 'gwd3d', cpl, 0, b1, b2, cap=2)[0]                    # matrix-only launch
    assert same_bits(only, mat)


def test_nan_and_degenerate_boxes(emul):
    b1, b2 = boxes(130, 40)
    b1[7, 0] = float('nan')
    b2[11, 3] = float('nan')
    b1[20, 3:6] = 1e-7
    b2[5, 4] = -1.0
    for packed in (0, 1):
        mat, rmin, ridx, cmin, cidx = run(emul, 'gwd3d', packed, 1, b1, b2)
        rv, ri = first_argmin(mat, 1)
        cv, ci = first_argmin(mat, 0)
        assert np.isnan(rmin).all() and np.isnan(cmin).all()          # NaN is the minimum
        assert np.array_equal(ridx, ri) and np.array_equal(cidx, ci)
        assert ridx[0] == 11 and cidx[0] == 7


@pytest.mark.parametrize('loss', ['gwd3d', 'kld3d', 'bd3d'])
@pytest.mark.parametrize('n,m', [(1, 1), (63, 5), (65, 33), (131, 129), (200, 256), (700, 40), (90, 256)])
@pytest.mark.parametrize('mode', [3, 4, 5])
def test_rowlane_minima_equal_matrix_minima(emul, loss, n, m, mode):
    """The ROW-lane kernel of the fused reductions (lanes on rows, column Gaussians in shared
    memory; what gd_pairwise_assign / gd_pairwise_row_argmin launch when no matrix is asked
    for): row / column (min, argmin) == those of the matrix the column-lane kernel writes, bit
    for bit, lowest index on ties.  mode 3 / 4: two / one rows per lane with the dynamic unit
    counter; 5: no column minima, static unit schedule.  Covers partial units, dead lanes,
    duplicate rows / columns (ties) and rows that leave the FAST cores (general sweep)."""
    b1, b2 = boxes(n, m)
    mat = run(emul, loss, 1, 0, b1, b2)[0]
    rv, ri = first_argmin(mat, 1)
    cv, ci = first_argmin(mat, 0)
    for cap in (1, 3):
        _, rmin, ridx, cmin, cidx = run(emul, loss, 1, 1, b1, b2, want_matrix=False,
                                        force_cpl=mode, cap=cap)
        assert same_bits(rmin, rv) and np.array_equal(ridx, ri), (loss, mode, cap, 'rows')
        if mode != 5:
            assert same_bits(cmin, cv) and np.array_equal(cidx, ci), (loss, mode, cap, 'cols')


@pytest.mark.parametrize('mode', [3, 4])
def test_rowlane_nan_and_degenerate_boxes(emul, mode):
    b1, b2 = boxes(130, 40)
    b1[7, 0] = float('nan')
    b2[11, 3] = float('nan')
    b1[20, 3:6] = 1e-7
    b2[5, 4] = -1.0
    mat = run(emul, 'gwd3d', 1, 0, b1, b2)[0]
    rv, ri = first_argmin(mat, 1)
    cv, ci = first_argmin(mat, 0)
    _, rmin, ridx, cmin, cidx = run(emul, 'gwd3d', 1, 1, b1, b2, want_matrix=False, force_cpl=mode)
    assert np.isnan(rmin).all() and np.isnan(cmin).all()              # NaN is the minimum
    assert np.array_equal(ridx, ri) and np.array_equal(cidx, ci)
    assert ridx[0] == 11 and cidx[0] == 7
    # one NaN row only: the other rows keep their finite minima
    b1, b2 = boxes(130, 40)
    b1[7, 0] = float('nan')
    mat = run(emul, 'gwd3d', 1, 0, b1, b2)[0]
    rv, ri = first_argmin(mat, 1)
    cv, ci = first_argmin(mat, 0)
    _, rmin, ridx, cmin, cidx = run(emul, 'gwd3d', 1, 1, b1, b2, want_matrix=False, force_cpl=mode)
    assert np.isnan(rmin[7]) and ridx[7] == 0 and np.isfinite(np.delete(rmin, 7)).all()
    assert same_bits(np.delete(rmin, 7), np.delete(rv, 7)) and np.array_equal(ridx, ri)
    assert np.isnan(cmin).all() and (cidx == 7).all()
