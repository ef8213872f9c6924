"""GPU: batches of node LPs sharded over several handles through the PRODUCT API (MultiGpuBatchLP
behind SharedLP, used by BranchAndBound._prefetch_frontier and the strong-branching batch).

With one visible GPU two handles share device 0 (the sharding, threading and merging are the same; the
bound exchange falls back to the host because NCCL needs distinct devices). With two or more GPUs the
shards run on different devices and agree on [incumbent, dual bound] through blp_allreduce_min."""
import json
import os

import numpy as np
import pytest
import torch

from simple_mip_solver_b200 import (BaseNode, BranchAndBound, CyLPArray, MILPInstance, PseudoCostBranchNode,
                                    engine)
from simple_mip_solver_b200.compat.cylp_like import SharedLP
from simple_mip_solver_b200.instances import frontier_nodes, grumpy_random_mip

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
SCALE1 = json.load(open(os.path.join(GOLD, 'scale_1_models.json')))
EXAMPLES = json.load(open(os.path.join(GOLD, 'example_models.json')))


def devices():
    return [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]


def model_from(rec):
    return MILPInstance(A=np.array(rec['A']), b=CyLPArray(rec['b']), c=CyLPArray(rec['c']),
                        l=CyLPArray(rec['l']), u=CyLPArray(rec['u']), sense=['Min', '>='],
                        integerIndices=list(rec['integer_indices']), numVars=len(rec['c']))


def tree_of(bb):
    return {idx: (bb.tree.get_parent(idx), v.attr['node']._b_idx, v.attr['node']._b_dir,
                  v.attr['node'].objective_value) for idx, v in bb.tree.nodes.items()}


@pytest.mark.parametrize('method', ['auto', 'pdhg'])
def test_same_tree_as_one_gpu(blp_lib, monkeypatch, method):
    monkeypatch.setattr(SharedLP, 'default_method', method)
    recs = [EXAMPLES['random'], EXAMPLES['small_branch']] + [r for r in list(SCALE1.values())[::8]]
    for rec in recs:
        for Node, kw in ((BaseNode, dict(gomory_cuts=False)),
                         (PseudoCostBranchNode, dict(pseudo_costs={}, gomory_cuts=False))):
            trees = []
            for devs in (None, devices()):
                monkeypatch.setattr(SharedLP, 'default_devices', devs)
                bb = BranchAndBound(model_from(rec), Node, frontier_batch=8,
                                    **{k: (dict(v) if isinstance(v, dict) else v) for k, v in kw.items()})
                bb.solve()
                trees.append((bb.status, bb.objective_value, tree_of(bb)))
                sh = bb.model.lp._shared
                if devs is not None:
                    assert isinstance(sh.engine, engine.MultiGpuBatchLP)
                    assert sh.engine.uses_nccl == (len(set(devs)) > 1)
                sh.close()
            assert trees[0][0] == trees[1][0]
            if method == 'auto':            # exact simplex: identical search
                assert trees[0] == trees[1]
            else:
                assert abs(trees[0][1] - trees[1][1]) <= 1e-6 * max(1.0, abs(trees[0][1]))


def test_sharded_frontier_and_bound_exchange(blp_lib):
    """A PDHG batch of 96 C4-shaped frontier nodes split over the handles equals the single-handle
    result node for node (each node LP is solved independently of its batch), and the exchanged pair is
    the min over all nodes."""
    from simple_mip_solver_b200.instances import numpy_random_mip
    d = numpy_random_mip(2000, 1000, density=5e-3, seed=3)
    one = engine.BatchLP(d.A, d.b, d.c)
    root = one.solve_batch(d.l[None], d.u[None])
    _, _, deltas = frontier_nodes(d, root.x[0], 0, 96, 8, seed=1, dense=False)
    ints = list(range(0, d.n, 2))
    a = one.solve_children(d.l, d.u, deltas, x0=root.x[0], y0=root.y[0], integer_indices=ints)
    multi = engine.MultiGpuBatchLP(d.A, d.b, d.c, devices=devices())
    b = multi.solve_children(d.l, d.u, deltas, x0=root.x[0], y0=root.y[0], integer_indices=ints)
    assert np.array_equal(a.status, b.status) and np.array_equal(a.frac_idx, b.frac_idx)
    assert np.allclose(a.objective, b.objective, rtol=1e-6, atol=1e-9)
    integral = (b.status == 0) & (b.frac_idx < 0)
    open_ = (b.status == 0) & ~integral
    inc = b.objective[integral].min() if integral.any() else np.inf
    low = b.lower_bound[open_].min() if open_.any() else np.inf
    assert b.stats['global_incumbent'] == inc and b.stats['global_lower_bound'] == low
    assert b.stats['bound_exchange'].startswith('blp_allreduce_min') == multi.uses_nccl
    print('devices', b.stats['devices'], 'exchange', b.stats['bound_exchange'])
    one.close()
    multi.close()
