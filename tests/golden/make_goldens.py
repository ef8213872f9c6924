"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE here.

The reference package (/root/reference) is imported on top of the HiGHS stand-in for CyLP/CLP
(oracle/ref_stubs.py) — its own BranchAndBound / BaseNode / PseudoCostBranchNode code drives every
result below. /root/reference is not available on the GPU box, so the outputs are committed:

  scale_1_models.json   the 64 fixtures of test_simple_mip_solver/scale_1_models parsed to arrays,
                        root LP optimum, MIP optimum (scipy.optimize.milp, independent check) and the
                        reference's B&B outcome per Node class (status, objective, solution,
                        evaluated nodes, tree: parent / branch variable / direction / LP value)
  example_models.json   the hand-written models of test_simple_mip_solver/example_models.py with
                        the reference's node-level and B&B-level answers
  floating_point.json   get_fraction / numerically_safe_cut input-output pairs

Run from the repo root:  python tests/golden/make_goldens.py
"""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp
from scipy.optimize import Bounds, LinearConstraint, milp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_stubs  # noqa: E402

ref_stubs.install()
ref_stubs.WARM_START = False        # cold solves: every LP answer is a function of the LP data only
import simple_mip_solver as ref  # noqa: E402
from coinor.cuppy.milpInstance import MILPInstance  # noqa: E402  (the stand-in)
from simple_mip_solver.utils.floating_point import get_fraction, numerically_safe_cut  # noqa: E402
from simple_mip_solver.algorithms.base_algorithm import BaseAlgorithm  # noqa: E402
from oracle.ref_lookalikes import CyLPArray  # noqa: E402
from oracle.mps_py import read_mps  # noqa: E402

REF_TESTS = '/root/reference/test_simple_mip_solver'
NODE_CASES = {
    'BaseNode': (ref.BaseNode, dict(gomory_cuts=False)),
    'DepthFirstSearchNode': (ref.DepthFirstSearchNode, dict(gomory_cuts=False)),
    'PseudoCostBranchNode': (ref.PseudoCostBranchNode, dict(pseudo_costs={}, gomory_cuts=False)),
    'PseudoCostBranchDepthFirstSearchNode': (ref.PseudoCostBranchDepthFirstSearchNode,
                                             dict(pseudo_costs={}, gomory_cuts=False)),
    'BaseNode_gomory': (ref.BaseNode, dict()),
}


def fl(v):
    if v is None:
        return None
    v = float(v)
    return v if np.isfinite(v) else ('inf' if v > 0 else '-inf')


def run_bb(make_model, Node, kwargs, backend='highs'):
    # strong branching is only meaningful from the parent's basis (5 warm pivots, base_node.py:
    # 608, 645), so pseudo-cost runs use the warm-started stand-in; the others solve cold.
    # The textbook dual simplex (backend 'dual_simplex') is deterministic, so there every LP starts
    # from the basis the reference hands it, exactly as with CLP (base_node.py:589, 608).
    ref_stubs.LP_BACKEND = backend
    ref_stubs.WARM_START = 'pseudo_costs' in kwargs or backend == 'dual_simplex'
    kw = {k: (dict(v) if isinstance(v, dict) else v) for k, v in kwargs.items()}
    bb = ref.BranchAndBound(make_model(), Node, **kw)
    bb.solve()
    tree = {}
    for idx, vert in bb.tree.nodes.items():
        n = vert.attr['node']
        parent = bb.tree.get_parent(idx)
        tree[str(idx)] = [parent, n._b_idx, n._b_dir, fl(n.objective_value),
                          None if n.lp_feasible is None else bool(n.lp_feasible),
                          None if n.mip_feasible is None else bool(n.mip_feasible)]
    out = dict(status=bb.status, objective=fl(bb.objective_value),
               solution=None if bb.solution is None else [float(v) for v in bb.solution],
               evaluated_nodes=bb.evaluated_nodes, tree=tree)
    if 'pseudo_costs' in bb._kwargs:
        out['pseudo_costs'] = {str(i): {d: dict(cost=float(e['cost']), times=int(e['times']))
                                        for d, e in v.items()} for i, v in bb._kwargs['pseudo_costs'].items()}
    ref_stubs.LP_BACKEND = 'highs'
    return out


def mip_optimum(A, b, c, l, u, ints):
    integrality = np.zeros(len(c)); integrality[ints] = 1
    r = milp(c, constraints=LinearConstraint(A, lb=b, ub=np.inf), bounds=Bounds(l, u), integrality=integrality)
    return fl(r.fun) if r.status == 0 else None


def scale_1():
    out = {}
    folder = os.path.join(REF_TESTS, 'scale_1_models')
    for name in sorted(os.listdir(folder)):
        if not name.endswith('.mps'):
            continue
        path = os.path.join(folder, name)
        mdl = read_mps(path)
        assert set(mdl.row_senses) <= {'L'}
        # canonical form as BaseAlgorithm produces it: -A x >= -b (base_algorithm.py:47-61)
        A, b = -mdl.A.toarray(), -mdl.rhs
        u = np.where(np.isinf(mdl.u), 1e308, mdl.u)
        rec = dict(A=A.tolist(), b=b.tolist(), c=mdl.c.tolist(), l=mdl.l.tolist(), u=u.tolist(),
                   integer_indices=list(mdl.integer_indices))
        root = ref.BranchAndBound(MILPInstance(file_name=path)).root_node
        root._bound_lp()
        rec['root_lp'] = dict(objective=fl(root.objective_value), solution=[float(v) for v in root.solution],
                              mip_feasible=bool(root.mip_feasible))
        rec['mip_optimum'] = mip_optimum(A, b, mdl.c, mdl.l, mdl.u, mdl.integer_indices)
        rec['reference'] = {}
        rec['reference_ds'] = {}       # the same runs with oracle/dual_simplex.py answering lp.dual()
        for label, (Node, kw) in NODE_CASES.items():
            rec['reference'][label] = run_bb(lambda: MILPInstance(file_name=path), Node, kw)
            rec['reference_ds'][label] = run_bb(lambda: MILPInstance(file_name=path), Node, kw, 'dual_simplex')
        out[name[:-4]] = rec
    return out


def example_models():
    import importlib
    from test_simple_mip_solver import example_models as em
    out = {}
    for name in ('no_branch', 'small_branch', 'infeasible', 'infeasible2', 'unbounded', 'random',
                 'cut1', 'cut2', 'cut3', 'square', 'negative', 'h3p1', 'h3p1_0', 'h3p1_1', 'h3p1_2',
                 'h3p1_3', 'h3p1_4', 'h3p1_5', 'lift_project'):
        if not hasattr(em, name):
            continue

        def fresh(nm=name):
            importlib.reload(em)
            return getattr(em, nm)
        m = BaseAlgorithm._convert_constraints_to_greq(fresh())
        A = m.A.toarray() if sp.issparse(m.A) else np.asarray(m.A, dtype=float)
        rec = dict(A=A.tolist(), b=[float(v) for v in m.lp.constraintsLower],   # ('unbounded' passes a short b)
                   c=[float(v) for v in np.asarray(m.lp.objective).ravel()],
                   l=[float(v) for v in m.l], u=[float(min(v, 1e308)) for v in m.u],
                   integer_indices=list(m.integerIndices))
        try:
            node = ref.BaseNode(m.lp, m.integerIndices, idx=0)
        except AssertionError as e:     # e.g. 'negative': the reference rejects x < 0 models
            out[name] = dict(rec, rejected=str(e))
            continue
        node._bound_lp()
        rec['root_lp'] = dict(lp_feasible=bool(node.lp_feasible), unbounded=bool(node.unbounded),
                              objective=fl(node.objective_value),
                              solution=None if node.solution is None else [float(v) for v in node.solution],
                              mip_feasible=bool(node.mip_feasible),
                              most_fractional_index=node._most_fractional_index)
        rec['reference'] = {}
        rec['reference_ds'] = {}
        if name != 'unbounded':
            for label, (Node, kw) in NODE_CASES.items():
                for key, backend in (('reference', 'highs'), ('reference_ds', 'dual_simplex')):
                    try:
                        rec[key][label] = run_bb(fresh, Node, kw, backend)
                    except Exception as e:       # a reference failure on this model is recorded, not hidden
                        rec[key][label] = dict(error=f'{type(e).__name__}: {e}')
        out[name] = rec
    return out


def floating_point():
    rng = np.random.default_rng(7)
    fr = []
    xs = list(rng.uniform(-5, 5, 60)) + [0.5, 0.25, 1 / 3, 0.9991, 1e-8, 123456.7, -0.75, 2.0, 0.0, 1e-3, 0.999]
    for x in xs:
        for est in (None, 'over', 'under'):
            for mt in (1000, 50):
                n, d = get_fraction(float(x), max_term=mt, estimate=est)
                fr.append(dict(x=float(x), max_term=mt, estimate=est, n=int(n), d=int(d)))
    cuts = []
    for _ in range(40):
        k = int(rng.integers(2, 8))
        pi = rng.uniform(-10, 10, k) * (rng.random(k) > 0.2)
        pi0 = float(rng.uniform(-10, 10))
        for est in ('over', 'under'):
            sp_, sp0 = numerically_safe_cut(CyLPArray(pi), pi0, estimate=est)
            cuts.append(dict(pi=pi.tolist(), pi0=pi0, estimate=est, safe_pi=[float(v) for v in sp_], safe_pi0=float(sp0)))
    return dict(get_fraction=fr, numerically_safe_cut=cuts)


if __name__ == '__main__':
    for fname, fn in (('example_models.json', example_models), ('floating_point.json', floating_point),
                      ('scale_1_models.json', scale_1)):
        data = fn()
        with open(os.path.join(HERE, fname), 'w') as fh:
            json.dump(data, fh, indent=None, separators=(',', ':'))
        print(fname, len(data), 'entries', os.path.getsize(os.path.join(HERE, fname)), 'bytes')
