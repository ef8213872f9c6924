"""CPU: the native MPS reader (libblp.so, blp_mps_*) against the pure-Python restatement of the
same dialect (oracle/mps_py.py) and against known model data.

The reference reads its fixtures through CLP's MPS reader (MILPInstance(file_name=...),
test_simple_mip_solver/helpers.py:42); all 64 scale_1_models files and the example models are
compared file by file when /root/reference is present, and the 64 models are always re-written
from the committed goldens in CLP's dialect and read back bit-exactly.
"""
import glob
import json
import os

import numpy as np
import pytest

from oracle.mps_py import read_mps as read_mps_py
from simple_mip_solver_b200.compat.mps import read_mps
from simple_mip_solver_b200.engine import BlpError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE1 = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'scale_1_models.json')))
REF = '/root/reference/test_simple_mip_solver'


def _same(a, b):
    b.A.sort_indices()
    assert a.name == b.name
    assert a.row_names == b.row_names and a.col_names == b.col_names and a.row_senses == b.row_senses
    assert np.array_equal(a.A.indptr, b.A.indptr) and np.array_equal(a.A.indices, b.A.indices)
    assert np.array_equal(a.A.data, b.A.data)
    for f in ('rhs', 'c', 'l', 'u'):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert a.integer_indices == b.integer_indices and a.obj_offset == b.obj_offset


def _clp_text(name, A, rhs, c, u):
    """The model in the layout CLP's writer produces for the reference's fixtures (L rows, UI bounds)."""
    m, n = A.shape
    out = ['NAME          %s' % name, 'ROWS', ' N  OBJROW'] + [' L  R_%d' % i for i in range(m)] + ['COLUMNS']
    for j in range(n):
        items = [('OBJROW', float(c[j]))] if c[j] != 0 else []
        items += [('R_%d' % i, float(A[i, j])) for i in range(m) if A[i, j] != 0]
        for k in range(0, len(items), 2):
            out.append('    x_%d  ' % j + '  '.join('%s  %r' % it for it in items[k:k + 2]))
    out.append('RHS')
    items = [('R_%d' % i, float(rhs[i])) for i in range(m) if rhs[i] != 0]
    for k in range(0, len(items), 2):
        out.append('    RHS  ' + '  '.join('%s  %r' % it for it in items[k:k + 2]))
    out += ['BOUNDS'] + [' UI BOUND  x_%d  %r' % (j, float(u[j])) for j in range(n)] + ['ENDATA']
    return '\n'.join(out) + '\n'


@pytest.mark.parametrize('label', sorted(SCALE1))
def test_golden_models_round_trip(tmp_path, label):
    g = SCALE1[label]
    # goldens hold the canonical form min c.x, A x >= b of "max obj.x, -A x <= -b"
    A, b, c, u = -np.array(g['A'], float), -np.array(g['b'], float), np.array(g['c'], float), np.array(g['u'], float)
    p = tmp_path / 'm.mps'
    p.write_text(_clp_text(label[:8], A, b, c, u))
    mdl = read_mps(str(p))
    _same(mdl, read_mps_py(str(p)))
    # columns without any entry in the file do not exist for an MPS reader: compare the ones that do
    keep = [j for j in range(A.shape[1])]
    assert mdl.A.shape == A.shape
    assert np.array_equal(mdl.A.toarray(), A) and np.array_equal(mdl.rhs, b) and np.array_equal(mdl.c, c)
    assert np.array_equal(mdl.u, u) and np.array_equal(mdl.l, np.zeros(len(u)))
    assert mdl.integer_indices == keep and mdl.row_senses == ['L'] * A.shape[0]


DIALECT = """* a comment line
NAME          dialect   extra
ROWS
 N  COST
 N  SECOND_OBJ
 G  lim1
 E  eq
 L  empty_row
 L  lim2
COLUMNS
    MARKER                 'MARKER'                 'INTORG'
    a   COST   1.5   lim1   2
    a   eq     -1e0
    b   lim1   1   SECOND_OBJ  7
    b   lim1   0.5
    MARKER                 'MARKER'                 'INTEND'
    c   COST   -2.   lim2   3.25
    d   eq     4
    unknown_row_user   nowhere   9
RANGES
    RNG   lim1   5
RHS
    RHS   COST   -3.5   lim1   4
    eq    -2
    lim2  1e30
BOUNDS
 UP BND  a   -1
 LO BND  b   1
 UP b   8
 MI BND  c
 FR d
 BV BND  e
 FX BND  f   2.5
 LI BND  g   -3
 PL BND  g
 SC BND  h   4
ENDATA
 L  after_the_end
"""


def test_dialect_details(tmp_path):
    p = tmp_path / 'd.mps'
    p.write_text(DIALECT)
    mdl = read_mps(str(p))
    _same(mdl, read_mps_py(str(p)))
    assert mdl.name == 'dialect'
    assert mdl.row_names == ['lim1', 'eq', 'empty_row', 'lim2'] and mdl.row_senses == ['G', 'E', 'L', 'L']
    assert mdl.col_names == ['a', 'b', 'c', 'd', 'unknown_row_user', 'e', 'f', 'g', 'h']
    A = mdl.A.toarray()
    assert np.array_equal(A[:, :4], [[2, 1.5, 0, 0], [-1, 0, 0, 4], [0, 0, 0, 0], [0, 0, 3.25, 0]])   # duplicates summed
    assert not A[:, 4:].any()
    assert list(mdl.c[:4]) == [1.5, 0, -2, 0] and mdl.obj_offset == 3.5
    assert list(mdl.rhs) == [4, -2, 0, 1e30]
    inf = np.inf
    assert list(mdl.l) == [-inf, 1, -inf, -inf, 0, 0, 2.5, -3, 0]        # UP < 0 without LO: lower becomes -inf
    assert list(mdl.u) == [-1, 8, inf, inf, inf, 1, 2.5, inf, inf]
    assert mdl.integer_indices == [0, 1, 5, 7]                           # marker columns, BV, LI


def test_errors(tmp_path):
    with pytest.raises(BlpError, match='cannot open'):
        read_mps(str(tmp_path / 'missing.mps'))
    p = tmp_path / 'bad.mps'
    p.write_text('ROWS\n N obj\n L r\nCOLUMNS\n    x  r  not_a_number\nENDATA\n')
    with pytest.raises(BlpError, match='line 5'):
        read_mps(str(p))


def test_large_file_is_fast(tmp_path):
    """A C4-sized model (10 000 x 5 000, ~100k entries) parses in well under a second."""
    import time
    from simple_mip_solver_b200.instances import numpy_random_mip
    d = numpy_random_mip(10000, 5000, density=2e-3, seed=2)
    A = d.A.tocsc()
    lines = ['NAME big', 'ROWS', ' N  OBJ'] + [' G  r%d' % i for i in range(d.m)] + ['COLUMNS']
    for j in range(d.n):
        lines.append('    x%d  OBJ  %r' % (j, float(d.c[j])))
        for k in range(A.indptr[j], A.indptr[j + 1]):
            lines.append('    x%d  r%d  %r' % (j, A.indices[k], float(A.data[k])))
    lines += ['RHS'] + ['    RHS  r%d  %r' % (i, float(d.b[i])) for i in range(d.m)]
    lines += ['BOUNDS'] + [' UI BOUND  x%d  %r' % (j, float(d.u[j])) for j in range(d.n)] + ['ENDATA']
    p = tmp_path / 'big.mps'
    p.write_text('\n'.join(lines) + '\n')
    t = time.perf_counter()
    mdl = read_mps(str(p))
    dt = time.perf_counter() - t
    B = d.A.copy()
    B.sort_indices()
    assert np.array_equal(mdl.A.indptr, B.indptr) and np.array_equal(mdl.A.indices, B.indices)
    assert np.array_equal(mdl.A.data, B.data) and np.array_equal(mdl.rhs, d.b) and np.array_equal(mdl.c, d.c)
    assert mdl.integer_indices == list(range(d.n))
    assert dt < 1.0, dt


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present')
def test_every_reference_fixture_reads_like_the_python_reader():
    files = sorted(glob.glob(os.path.join(REF, 'scale_1_models', '*.mps')) +
                   glob.glob(os.path.join(REF, 'example_models', '*.mps')))
    assert len(files) >= 64
    for f in files:
        _same(read_mps(f), read_mps_py(f))
