import ctypes
import json
import os
import re

import numpy as np
import pytest

from simple_mip_solver_b200.instances import frontier_nodes, grumpy_random_mip, numpy_random_mip

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE1 = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'scale_1_models.json')))


def test_generator_reproduces_the_reference_fixtures():
    """generate_random_variety(scale=1) (example_models.py:28-48): the file name says
    constraints/variables but the values land in (numVars, numCons) — see SURVEY.md section 4."""
    lv = {'low': 0, 'high': 1}
    for name, rec in SCALE1.items():
        t = name.split('_')
        f = dict(zip(('constraints', 'variables', 'density', 'obj', 'cons', 'tightness'),
                     (t[1], t[3], t[5], t[9], t[13], t[15])))
        d = grumpy_random_mip(numVars=(2, 4)[lv[f['constraints']]], numCons=(2, 4)[lv[f['variables']]],
                              density=(.2, .8)[lv[f['density']]], maxObjCoeff=(10, 100)[lv[f['obj']]],
                              maxConsCoeff=(10, 100)[lv[f['cons']]], tightness=(2, 8)[lv[f['tightness']]])
        assert np.array_equal(d.A.toarray(), np.array(rec['A'])), name
        assert np.array_equal(d.b, rec['b']) and np.array_equal(d.c, rec['c']), name
        assert np.array_equal(d.u, rec['u']) and d.integer_indices == rec['integer_indices'], name


def test_vectorised_generator_shapes():
    d = numpy_random_mip(2000, 1000, density=0.01, seed=2)
    assert d.A.shape == (1000, 2000) and abs(d.A.nnz - 20000) < 1500
    assert (d.A.data < 0).all() and (d.b < 0).all() and (d.c < 0).all()
    d2 = numpy_random_mip(2000, 1000, density=0.01, seed=2)
    assert (d.A != d2.A).nnz == 0


def test_frontier_nodes_are_split_invariant_and_screen_feasible():
    d = numpy_random_mip(300, 120, density=0.05, seed=2)
    x = np.random.default_rng(0).uniform(0, 3, 300)
    lb, ub, deltas = frontier_nodes(d, x, 0, 12, 6, seed=1)
    lb2, ub2, _ = frontier_nodes(d, x, 4, 4, 6, seed=1)
    assert np.array_equal(lb[4:8], lb2) and np.array_equal(ub[4:8], ub2)
    assert (lb <= ub).all()
    assert ((d.A @ lb.T).T >= d.b - 1e-9).all()      # all-lower-bound point satisfies every row
    for k, node in enumerate(deltas):
        for j, lo, hi in node:
            assert lb[k, j] == lo and ub[k, j] == hi


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports each function include/blp.h declares (no GPU needed)."""
    from simple_mip_solver_b200 import _build, engine
    _build.build_extension()
    header = open(os.path.join(ROOT, 'include', 'blp.h')).read()
    declared = set(re.findall(r'\b(blp_[a-z_]+)\s*\(', header))
    declared -= {'blp_handle_s'}
    lib = ctypes.CDLL(str(engine._LIB_PATH))
    for sym in declared:
        assert hasattr(lib, sym), f'{sym} declared in blp.h but not exported'
    assert declared == set(engine.EXPORTED_SYMBOLS)
    o = engine.default_opts()
    assert o.eps_rel == 1e-7 and o.eval_every == 64 and o.max_iters == 2000000
    assert lib.blp_ld(1) == 64 and lib.blp_ld(65) == 128
    assert engine.load_library().blp_version().decode().endswith('sm_100a')


def test_ctypes_structs_mirror_the_header():
    """blp_opts / blp_stats cross the C ABI by value layout: the ctypes mirrors in engine.py must list the same fields,
    in the same order and with the same C types as include/blp.h, and blp_default_opts must fill every field."""
    import ctypes as C
    from simple_mip_solver_b200 import engine
    header = open(os.path.join(ROOT, 'include', 'blp.h')).read()
    ctype = {'double': C.c_double, 'int': C.c_int}
    for name, mirror in (('blp_opts', engine.BlpOpts), ('blp_stats', engine.BlpStats)):
        body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (name, name), header, re.S).group(1)
        body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
        fields = re.findall(r'\b(double|int)\s+([a-z_0-9]+)\s*;', body)
        assert [(f, ctype[t]) for t, f in fields] == list(mirror._fields_), name
    o = engine.default_opts()
    assert o.freeze == 1 and o.freeze_margin == 0.05 and o.step_safety == 0.98 and o.obj_cutoff == float('inf')
    with pytest.raises(TypeError):
        engine.default_opts(no_such_option=1)


def test_product_fails_loudly_without_gpu_or_library(monkeypatch, tmp_path):
    import scipy.sparse as sp
    import torch
    from simple_mip_solver_b200 import engine
    if not torch.cuda.is_available():
        with pytest.raises(engine.BlpError):
            engine.BatchLP(sp.eye(2, format='csr'), np.zeros(2), np.ones(2))
    monkeypatch.setattr(engine, '_lib', None)
    monkeypatch.setattr(engine, '_LIB_PATH', tmp_path / 'missing.so')
    with pytest.raises(engine.BlpError, match='no CPU fallback'):
        engine.load_library()


def test_comm_entry_points_without_a_gpu():
    """blp_comm_unique_id binds NCCL at run time (no GPU needed); the handle-taking entry points
    reject a NULL handle with an error code and a message instead of crashing."""
    from simple_mip_solver_b200 import engine
    lib = engine.load_library()
    a, b = ctypes.create_string_buffer(128), ctypes.create_string_buffer(128)
    assert lib.blp_comm_unique_id(a) == 0 and lib.blp_comm_unique_id(b) == 0
    assert a.raw != b.raw and any(a.raw)
    assert lib.blp_comm_unique_id(None) == -1
    v = (ctypes.c_double * 2)(1.0, 2.0)
    assert lib.blp_allreduce_min(None, v) == -1 and b'NULL' in lib.blp_last_error()
    assert lib.blp_comm_init(None, 2, 0, a) == -1
    assert lib.blp_comm_destroy(None) == -1
    assert list(v) == [1.0, 2.0]
