"""Synthetic MILP instances of the shapes named in BASELINE.json.

``grumpy_random_mip`` restates the draw order of GrUMPy's ``GenerateRandomMIP`` (not vendored
in the reference; used at ``test_simple_mip_solver/example_models.py:12-25``) so that the
checked-in ``scale_1_models/*.mps`` fixtures are reproduced exactly (tests/test_instances.py).
``numpy_random_mip`` draws from the same distributions with a vectorised generator for the
large configurations (C4/C5), where the python-loop generator would need n*m draws.

Everything here returns plain arrays in the solver's canonical form

    min c.x   s.t.  A x >= b,  l <= x <= u

which is what ``generate_random_MILPInstance`` hands to ``MILPInstance`` in the reference
(``A=-A, b=-b, c=-objective``; example_models.py:24).
"""
from __future__ import annotations

import random as _pyrandom
from dataclasses import dataclass
from typing import List

import numpy as np
import scipy.sparse as sp


@dataclass
class MipData:
    A: sp.csr_matrix          # m x n, rows are ">= b"
    b: np.ndarray             # m
    c: np.ndarray             # n (minimise)
    l: np.ndarray             # n
    u: np.ndarray             # n
    integer_indices: List[int]

    @property
    def n(self) -> int:
        return self.A.shape[1]

    @property
    def m(self) -> int:
        return self.A.shape[0]


def grumpy_random_mip(numVars=40, numCons=20, density=0.2, maxObjCoeff=10, maxConsCoeff=10,
                      tightness=2, rand_seed=2) -> MipData:
    """Exact restatement of GrUMPy's generator as wrapped by example_models.py:12-25.

    Draw order: objective coefficients for every variable; then, variable by variable and row
    by row, one uniform draw and (only on a hit) one integer coefficient; then the right-hand
    sides. The packing problem ``max obj.x, A x <= rhs`` is returned negated into canonical form.
    """
    rng = _pyrandom.Random()
    rng.seed(rand_seed)
    obj = [rng.randint(1, maxObjCoeff) for _ in range(numVars)]
    dense = np.zeros((numCons, numVars))
    for j in range(numVars):
        for i in range(numCons):
            if rng.random() <= density:
                dense[i, j] = rng.randint(1, maxConsCoeff)
    lo = int(numVars * density * maxConsCoeff / tightness)
    hi = int(numVars * density * maxConsCoeff / 1.5)
    rhs = [rng.randint(lo, hi) for _ in range(numCons)]
    return MipData(A=sp.csr_matrix(-dense), b=-np.asarray(rhs, dtype=float),
                   c=-np.asarray(obj, dtype=float), l=np.zeros(numVars),
                   u=np.full(numVars, float(maxObjCoeff)),
                   integer_indices=list(range(numVars)))


def numpy_random_mip(numVars, numCons, density, maxObjCoeff=10, maxConsCoeff=10, tightness=2,
                     seed=2) -> MipData:
    """Same distributions as ``grumpy_random_mip``, vectorised (PCG64), for large shapes.

    The number of nonzeros is drawn binomially and positions are sampled without replacement,
    which is the same law as an independent Bernoulli(density) per entry.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    obj = rng.integers(1, maxObjCoeff + 1, size=numVars).astype(float)
    total = numVars * numCons
    nnz = int(rng.binomial(total, density))
    flat = np.unique(rng.integers(0, total, size=int(nnz * 1.02) + 16))
    rng.shuffle(flat)
    flat = np.sort(flat[:nnz])
    rows = (flat // numVars).astype(np.int32)
    cols = (flat % numVars).astype(np.int32)
    vals = rng.integers(1, maxConsCoeff + 1, size=flat.size).astype(float)
    A = sp.csr_matrix((-vals, (rows, cols)), shape=(numCons, numVars))
    lo = int(numVars * density * maxConsCoeff / tightness)
    hi = int(numVars * density * maxConsCoeff / 1.5)
    rhs = rng.integers(lo, max(hi, lo) + 1, size=numCons).astype(float)
    return MipData(A=A, b=-rhs, c=-obj, l=np.zeros(numVars),
                   u=np.full(numVars, float(maxObjCoeff)),
                   integer_indices=list(range(numVars)))


def random_dive_bounds(data: MipData, x_root: np.ndarray, batch: int, max_depth: int,
                       seed: int = 0):
    """Per-node bounds of ``batch`` open nodes reached by seeded random dives from the root.

    Node k applies depth_k ~ U{1..max_depth} branching decisions; each picks an integer variable
    (fractional ones at the root LP solution first, as a B&B dive would) and moves one bound to
    floor/ceil of its root value, the way ``BaseNode._base_branch`` does (base_node.py:595-600).
    Returns (lb, ub) as [batch, n] arrays plus the list of (node, var, lb, ub) deltas.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    n = data.n
    ints = np.asarray(data.integer_indices)
    frac = np.abs(x_root[ints] - np.round(x_root[ints]))
    order = ints[np.argsort(-frac, kind='stable')]
    n_frac = max(int((frac > 1e-4).sum()), 1)
    lb = np.tile(data.l, (batch, 1))
    ub = np.tile(data.u, (batch, 1))
    deltas = []
    for k in range(batch):
        depth = int(rng.integers(1, max_depth + 1))
        pool = order[:max(n_frac, depth)]
        picks = rng.choice(pool, size=min(depth, pool.size), replace=False)
        for j in picks:
            v = x_root[j]
            if rng.random() < 0.5:
                ub[k, j] = min(ub[k, j], np.floor(v))
            else:
                lb[k, j] = max(lb[k, j], np.ceil(v)) if np.ceil(v) <= ub[k, j] else lb[k, j]
            deltas.append((k, int(j), lb[k, j], ub[k, j]))
    return lb, ub, deltas


# Frontier nodes with committed golden answers (bench_data/<workload>_children.npz, made by
# tests/tools/make_child_goldens.py): ids GOLD_FIRST .. GOLD_FIRST + GOLD_COUNT - 1 of frontier_nodes(seed=0).
GOLD_FIRST = 20_000_000
GOLD_COUNT = 64


def frontier_nodes(data: MipData, x_root: np.ndarray, first: int, count: int, max_depth: int,
                   seed: int = 0, p_down: float = 0.75, dense: bool = True):
    """Bounds of open nodes ``first .. first+count-1`` of a synthetic frontier.

    Node k is reached from the root by depth_k ~ U{1..max_depth} branching decisions on integer
    variables that are fractional at the root LP vertex ``x_root`` (most fractional first, as
    ``BaseNode._most_fractional_index`` would pick them), each moving one bound to floor/ceil of
    the root value as ``BaseNode._base_branch`` does (base_node.py:595-600). An up-branch that
    would make a row unsatisfiable even with every other variable at its lower bound is turned
    into a down-branch, so every node passes the row-activity screen and needs a real LP solve.
    Each node has its own generator seeded by (seed, k): the frontier does not depend on how it
    is split into batches or over GPUs. Returns (lb, ub) of shape [count, n] and the deltas: per node the
    list of ``(var, lb, ub)`` changes against the root bounds; with ``dense=False`` lb and ub are None."""
    n = data.n
    ints = np.asarray(data.integer_indices)
    vals = x_root[ints]
    frac = np.minimum(vals - np.floor(vals), np.ceil(vals) - vals)
    order = ints[np.argsort(-frac, kind='stable')]
    n_frac = int((frac > 1e-4).sum())
    A = data.A.tocsc()
    base_act = data.A @ data.l              # activity with everything at its lower bound
    lb = np.tile(data.l, (count, 1)) if dense else None
    ub = np.tile(data.u, (count, 1)) if dense else None
    deltas = []
    for t in range(count):
        k = first + t
        rng = np.random.Generator(np.random.PCG64([seed, k]))
        cur = {}                                  # bounds this node has moved so far
        depth = int(rng.integers(1, max_depth + 1))
        pool = order[:max(n_frac, depth)]
        picks = rng.choice(pool, size=min(depth, pool.size), replace=False)
        act = base_act.copy()
        node = []
        for j in picks:
            v = x_root[j]
            lo, hi = cur.get(int(j), (data.l[j], data.u[j]))
            up = rng.random() >= p_down and np.ceil(v) <= hi and np.ceil(v) > lo
            if up:
                col = A.getcol(j)
                trial = act[col.indices] + col.data * (np.ceil(v) - lo)
                if (trial < data.b[col.indices] - 1e-9).any():      # rows are A x >= b
                    up = False
                else:
                    act[col.indices] = trial
                    lo = np.ceil(v)
            if not up:
                hi = max(min(hi, np.floor(v)), lo)
            cur[int(j)] = (float(lo), float(hi))
            if dense:
                lb[t, j], ub[t, j] = lo, hi
            node.append((int(j), float(lo), float(hi)))
        deltas.append(node)
    return lb, ub, deltas
