"""Golden values of the cut generating LP, produced by RUNNING THE UNMODIFIED REFERENCE here.

The reference's ``CutGeneratingLP`` (simple_mip_solver/utils/cut_generating_lp.py) builds its LP with
CyLP's multi-variable modelling algebra and solves it with CLP's primal simplex. Here that same code
runs on oracle/cylp_multivar.py (the algebra, HiGHS behind ``primal()``) on top of a disjunction
made by the reference's own ``BranchAndBound`` on the textbook dual simplex (ref_stubs 'dual_simplex':
the trees the device reproduces node for node). Recorded per model: the leaves that form the
disjunction (bounds), the size of the reference's LP, and for several points ``x_star`` the OPTIMUM
``x_star.pi - pi0`` the reference finds (the cut itself is one of possibly many optimal ones, kept
for inspection). tests/test_disjunctive_cut.py and tests/test_gpu_cglp.py hold the product to them.

Run from the repo root:  python tests/golden/make_cglp_goldens.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_stubs  # noqa: E402

ref_stubs.install()
import simple_mip_solver as ref  # noqa: E402
from coinor.cuppy.milpInstance import MILPInstance  # noqa: E402  (the stand-in)
from oracle.ref_lookalikes import CyLPArray  # noqa: E402

ref_cglp = ref_stubs.enable_cglp()           # rebinds the one name the reference's CGLP module takes from cylp

NODE_LIMIT = 8


def model_from(rec):
    # the fixtures store "no upper bound" as 1e308; the reference compares against getCoinInfinity() exactly
    u = np.array(rec['u'], dtype=float)
    u = np.where(u >= 1e300, ref_stubs.COIN_INFINITY, u)
    return MILPInstance(A=np.array(rec['A']), b=CyLPArray(rec['b']), c=CyLPArray(rec['c']), l=CyLPArray(rec['l']),
                        u=CyLPArray(u), sense=['Min', '>='], integerIndices=list(rec['integer_indices']),
                        numVars=len(rec['c']))


def points_of(x, seed):
    rng = np.random.default_rng(seed)
    return [x, np.maximum(x * 0.9, 0), np.maximum(x + 0.25, 0), np.maximum(x * rng.uniform(.7, 1.2, len(x)), 0)]


def main():
    scale1 = json.load(open(os.path.join(HERE, 'scale_1_models.json')))
    examples = json.load(open(os.path.join(HERE, 'example_models.json')))
    cases = [(k, examples[k]) for k in ('square', 'small_branch', 'cut1', 'cut2', 'lift_project', 'random')]
    cases += list(scale1.items())[::5]
    out = {}
    for seed, (name, rec) in enumerate(cases):
        ref_stubs.LP_BACKEND, ref_stubs.WARM_START = 'dual_simplex', True
        bb = ref.BranchAndBound(model_from(rec), ref.BaseNode, node_limit=NODE_LIMIT, gomory_cuts=False)
        bb.solve()
        root = bb.root_node
        if root.solution is None:
            continue
        cglp = ref_cglp.CutGeneratingLP(bb, root.idx)
        leaves = {str(n.idx): dict(lower=[float(v) for v in n.lp.variablesLower],
                                   upper=[float(min(v, 1e308)) for v in n.lp.variablesUpper])
                  for n in bb.tree.get_leaves(root.idx) if n.lp_feasible is not False}
        pts = []
        for p in points_of(np.asarray(root.solution, dtype=float), seed):
            pi, pi0 = cglp.solve(x_star=CyLPArray(p))
            ok = pi is not None
            pts.append(dict(x_star=[float(v) for v in p], status=int(cglp.lp.getStatusCode()),
                            optimum=float(np.dot(pi, p) - pi0) if ok else None,
                            pi=[float(v) for v in pi] if ok else None, pi0=float(pi0) if ok else None))
        out[name] = dict(node_limit=NODE_LIMIT, evaluated_nodes=bb.evaluated_nodes, terms=leaves,
                         n_variables=int(cglp.lp.nVariables), n_constraints=int(cglp.lp.nConstraints), points=pts)
        # the constructor's options (reference :13-50): a depth limit on the disjunction, other constraints for
        # every term, extra variable bounds (terms they empty are dropped, :113-121)
        x = np.asarray(root.solution, dtype=float)
        lo_x, hi_x = np.floor(x), np.floor(x) + 1
        options = dict(
            depth_1=dict(depth=1),
            root_rows=dict(A=root.lp.coefMatrix.copy(), b=CyLPArray(np.asarray(root.lp.constraintsLower).copy())),
            unit_box=dict(var_lb=CyLPArray(lo_x), var_ub=CyLPArray(hi_x)),
        )
        variants = {}
        for key, kw in options.items():
            try:
                other = ref_cglp.CutGeneratingLP(bb, root.idx, **kw)
                pi, pi0 = other.solve(x_star=CyLPArray(x))
            except Exception as exc:          # e.g. every term emptied by the box: recorded, the product must agree
                variants[key] = dict(error=type(exc).__name__)
                continue
            ok = pi is not None
            variants[key] = dict(n_variables=int(other.lp.nVariables), n_constraints=int(other.lp.nConstraints),
                                 status=int(other.lp.getStatusCode()),
                                 optimum=float(np.dot(pi, x) - pi0) if ok else None)
        out[name]['variants'] = variants
        print(name, 'terms', len(leaves), 'lp', cglp.lp.nVariables, 'x', cglp.lp.nConstraints,
              'optima', [None if q['optimum'] is None else round(q['optimum'], 6) for q in pts])
    ref_stubs.LP_BACKEND = 'highs'
    json.dump(out, open(os.path.join(HERE, 'cglp.json'), 'w'), indent=0)
    print('wrote', len(out), 'models to tests/golden/cglp.json')


if __name__ == '__main__':
    main()
