"""Run the Node API on a C4-sized MPS file in a process of its own and report what happened as one
JSON line (used by tests/test_gpu_configs.py::test_config4_shape_through_the_node_api)."""
import json
import os
import resource
import sys

import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from simple_mip_solver_b200 import BaseNode, BranchAndBound, MILPInstance, PseudoCostBranchNode  # noqa: E402

path = sys.argv[1]
model = MILPInstance(file_name=path)
out = dict(sparse_model=bool(sp.issparse(model.A)), n=model.numVars, n_int=len(model.integerIndices))
node = BaseNode(model.lp, model.integerIndices, idx=0)
node.bound(max_cut_generation_iterations=2, max_gomory_cuts=16)
out.update(bound_feasible=bool(node.lp_feasible), bound_objective=float(node.objective_value),
           cut_rounds=node.cut_generation_iterations, terminator=node.cut_generation_terminator,
           rows_in_lp=int(node.lp.nConstraints), gmic_added=node.number_gmic_added,
           exact_basis=bool(node.lp.has_exact_basis), first_lp_value=float(node.cut_generation_dual_bound.get(0, float('nan'))))
model.lp._shared.close()
model = MILPInstance(file_name=path)
bb = BranchAndBound(model, PseudoCostBranchNode, node_limit=3, pseudo_costs={}, gomory_cuts=False,
                    strong_branch_iters=5, frontier_batch=8)
bb.solve()
sh = bb.model.lp._shared
out.update(bb_status=bb.status, bb_nodes=bb.evaluated_nodes, bb_root_objective=float(bb.root_node.objective_value),
           bb_dual_bound=float(bb.dual_bound), pseudo_costs=len(bb._kwargs['pseudo_costs']),
           lps_solved=sh.lps_solved, gpu_calls=sh.solve_calls, kernel_launches=sh.kernel_launches,
           peak_rss_gb=resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6)
sh.close()
print(json.dumps(out), flush=True)
