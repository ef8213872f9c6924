"""Solve the root LP of a bench configuration with the HiGHS oracle and store its vertex, row
duals and basis under bench_data/ (the warm start both bench arms branch from)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from simple_mip_solver_b200.instances import grumpy_random_mip, numpy_random_mip
from oracle.highs_lp import HighsLP, HIGHS_INF

name, n, m, dens = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
if name == 'c3':     # the reference's own generator (bench.load_instance does the same)
    d = grumpy_random_mip(n, m, density=dens, maxObjCoeff=10, maxConsCoeff=10, tightness=2, rand_seed=2)
else:
    d = numpy_random_mip(n, m, density=dens, seed=2)
t = time.time()
r = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), d.l, d.u).solve()
print(name, 'status', r.status, 'obj', r.objective, 'iters', r.iterations, 'wall', time.time() - t)
np.savez_compressed(f'bench_data/{name}_root.npz', x=r.x, y=r.row_dual, col_basis=r.col_basis.astype(np.int8),
                    row_basis=r.row_basis.astype(np.int8), objective=r.objective, n=n, m=m, density=dens, seed=2,
                    nnz=d.A.nnz)
