"""More golden trees from the UNMODIFIED reference: GrUMPy-style random MILPs of 6-14 variables.

The 64 checked-in fixtures have 2-4 variables. These instances (the reference's own generator,
test_simple_mip_solver/example_models.py:12-25, restated in oracle/ref_stubs.GenerateRandomMIP and verified
against the checked-in MPS files) are large enough for deep trees, ties of the most-fractional rule, strong
branching over many candidates and several Gomory rounds per node. The reference's BranchAndBound runs on the
textbook dual simplex (ref_stubs 'dual_simplex'), whose pivoting the device reproduces bit for bit, so the
product has to build the SAME tree (tests/test_host_logic.py::test_fuzz_models_same_tree_as_reference).

Run from the repo root:  python tests/golden/make_fuzz_goldens.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_goldens as mg  # noqa: E402  (installs the stand-ins, imports the reference)

from test_simple_mip_solver.example_models import generate_random_MILPInstance  # noqa: E402

SHAPES = [(6, 4, .5), (8, 5, .4), (10, 6, .3), (12, 6, .3), (14, 8, .25)]
SEEDS = [3, 5, 8]


def instances():
    for (nv, nc, dens) in SHAPES:
        for seed in SEEDS:
            yield f'random_{nv}x{nc}_seed{seed}', (lambda nv=nv, nc=nc, dens=dens, seed=seed: generate_random_MILPInstance(
                numVars=nv, numCons=nc, density=dens, maxObjCoeff=10, maxConsCoeff=10, tightness=2, rand_seed=seed))


def main():
    out = {}
    # pass 1: HiGHS answers lp.dual() (what the host code of the first-order path is held to). A raw HiGHS
    # instance has to run BEFORE scipy.optimize.milp is first used in the process: the other way round, raw
    # instances return kNotset from run() (seen with scipy 1.18's bundled HiGHS; make_goldens.py happens to
    # solve a root LP first too)
    from oracle.highs_lp import HIGHS_INF, HighsLP
    assert HighsLP(np.eye(1), np.ones(1), np.ones(1), np.full(1, HIGHS_INF), np.zeros(1), np.full(1, 2.0)).solve().status == 0
    for name, make in instances():
        m = make()
        A = np.asarray(m.A, dtype=float)
        rec = dict(A=A.tolist(), b=[float(v) for v in m.b], c=[float(v) for v in np.asarray(m.lp.objective).ravel()],
                   l=[float(v) for v in m.l], u=[float(v) for v in m.u], integer_indices=list(m.integerIndices))
        rec['mip_optimum'] = mg.mip_optimum(A, np.asarray(m.b, float), np.asarray(rec['c']), np.asarray(m.l, float),
                                            np.asarray(m.u, float), list(m.integerIndices))
        rec['reference'] = {label: mg.run_bb(make, Node, kw) for label, (Node, kw) in mg.NODE_CASES.items()
                            if label != 'BaseNode_gomory'}
        assert all(r['status'] == 'optimal' for r in rec['reference'].values()), name
        out[name] = rec
    # pass 2: the textbook dual simplex, whose pivoting the device reproduces bit for bit
    for name, make in instances():
        rec = out[name]
        rec['reference_ds'] = {label: mg.run_bb(make, Node, kw, 'dual_simplex')
                               for label, (Node, kw) in mg.NODE_CASES.items()}
        print(name, 'optimum', rec['mip_optimum'],
              {k: (v['evaluated_nodes'], rec['reference'].get(k, {}).get('evaluated_nodes'))
               for k, v in rec['reference_ds'].items()})
    with open(os.path.join(HERE, 'fuzz_models.json'), 'w') as fh:
        json.dump(out, fh, indent=None, separators=(',', ':'))
    print('wrote', len(out), 'models,', os.path.getsize(os.path.join(HERE, 'fuzz_models.json')), 'bytes')


if __name__ == '__main__':
    main()
