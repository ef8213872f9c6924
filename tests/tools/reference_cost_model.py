"""The reference's own strong-branching cost model on config 3 ("report the
faithful reference cost model once").

Runs the UNMODIFIED reference (``/root/reference``) on the HiGHS stand-in for CyLP/CLP
(oracle/ref_stubs.py) — only possible in the authoring container — and times, for the candidates of
the 500 x 300 instance, the two halves of ``BaseNode._strong_branch`` (base_node.py:629-647):
  * ``_base_branch``: both child models are REBUILT in Python (variables, bounds, every constraint
    re-added, basis copied; base_node.py:592-608),
  * ``n.lp.dual()`` with ``maxNumIteration = 5`` for the two children.
The LP half runs HiGHS instead of CLP and the modelling half the stub's CyLP look-alike instead of
CyLP's Cython layer, so the figures are indicative, not CLP timings. Prints one JSON line.
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
from oracle import ref_stubs
ref_stubs.install()
from simple_mip_solver.nodes.base_node import BaseNode                       # the reference's class
from coinor.cuppy.milpInstance import MILPInstance                            # stub
from simple_mip_solver_b200.instances import grumpy_random_mip

d = grumpy_random_mip(500, 300, density=0.1, maxObjCoeff=10, maxConsCoeff=10, tightness=2, rand_seed=2)
# canonical form of the reference's nodes (base_node.py:111): min c.x, A x >= b
model = MILPInstance(A=np.asarray(d.A.todense()), b=d.b, c=d.c, l=d.l, u=d.u, sense=['Min', '>='],
                     integerIndices=d.integer_indices, numVars=d.n)
root = BaseNode(lp=model.lp, integer_indices=model.integerIndices, idx=0)
t = time.perf_counter(); root.bound(gomory_cuts=False) if 'gomory_cuts' in BaseNode.bound.__code__.co_varnames else root.bound(); t_root = time.perf_counter() - t
x = root.solution
ints = np.asarray(model.integerIndices)
frac = np.minimum(x[ints] - np.floor(x[ints]), np.ceil(x[ints]) - x[ints])
cand = [int(j) for j in ints[np.argsort(-frac, kind='stable')][:16] if frac[list(ints).index(j)] > 1e-4]
t_build = t_solve = 0.0
for j in cand:
    t = time.perf_counter()
    kids = {k: v for k, v in root._base_branch(j).items() if k in ('left', 'right')}
    t_build += time.perf_counter() - t
    t = time.perf_counter()
    for n in kids.values():
        n.lp.maxNumIteration = 5
        n.lp.dual()
    t_solve += time.perf_counter() - t
k = 2 * len(cand)
print(json.dumps(dict(instance='config 3: 500 x 300, density 0.1', children=k, root_bound_s=round(t_root, 4),
                      rebuild_s_per_child=round(t_build / k, 5), dual5_s_per_child=round(t_solve / k, 5),
                      rebuild_share=round(t_build / (t_build + t_solve), 3),
                      note='reference control flow on HiGHS/CyLP look-alike stubs; indicative only')))
