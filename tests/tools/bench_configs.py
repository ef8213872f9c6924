"""Timing of BASELINE.json configs 3 and 4 on one GPU next to the host simplex (same run).

config 3: 500 x 300, 10 % density: root + strong branching on 64 candidates (128 child LPs, one batch)
config 4: 10 000 x 5 000 sparse: 256 dive nodes, then 3 cut rounds of 32 appended dense rows each
          (shared pool, per-node row masks), re-solved warm.
Writes gpurun_out/configs.json.
"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import frontier_nodes, grumpy_random_mip
from oracle.highs_lp import HIGHS_INF, HighsLP

out = {}
cores = bench.host_cores()

# ---------------- config 3
d = grumpy_random_mip(500, 300, density=0.1, maxObjCoeff=10, maxConsCoeff=10, tightness=2, rand_seed=2)
t = time.perf_counter(); h = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), d.l, d.u); root = h.solve(); t_root_cpu = time.perf_counter() - t
x = root.x
ints = np.asarray(d.integer_indices)
frac = np.minimum(x[ints] - np.floor(x[ints]), np.ceil(x[ints]) - x[ints])
cand = ints[np.argsort(-frac, kind='stable')][:64]
deltas = []
for j in cand:
    deltas.append([(int(j), d.l[j], float(np.floor(x[j])))])
    deltas.append([(int(j), float(np.ceil(x[j])), d.u[j])])
lp = engine.BatchLP(d.A, d.b, d.c)
# round 2: LPs of this size go to the dual simplex kernels (one CTA per node LP)
lp.simplex_children(d.l, d.u, deltas[:4])                                                   # warm the library
t = time.perf_counter(); rr = lp.simplex_batch(d.l[None], d.u[None]); t_root_gpu = time.perf_counter() - t
sx = {}
for label, kw in (('children_from_status', dict(col_status=rr.col_status[0], row_status=rr.row_status[0])),
                  ('children_from_stored_factor', dict(col_status=rr.col_status[0], row_status=rr.row_status[0], parent_slot=0)),
                  ('children_5_pivots_from_stored_factor', dict(col_status=rr.col_status[0], row_status=rr.row_status[0], parent_slot=0, max_pivots=5))):
    best = None
    for rep in range(3):
        if 'parent_slot' in kw:
            lp.simplex_batch(d.l[None], d.u[None])            # the store holds the last call: the parent again
        t = time.perf_counter(); r = lp.simplex_children(d.l, d.u, deltas, **kw); dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    sx[label] = dict(seconds=best, lps_per_s=len(deltas) / best, kernel_ms=r.stats['step_kernel_ms'], max_pivots=int(r.pivots.max()),
                     mean_pivots=float(r.pivots.mean()))
r = lp.simplex_children(d.l, d.u, deltas, col_status=rr.col_status[0], row_status=rr.row_status[0])
t_gpu = sx['children_from_status']['seconds']
# the previous path (PDHG) for the record
t = time.perf_counter(); rp = lp.solve_children(d.l, d.u, deltas, x0=rr.x[0], y0=rr.y[0], integer_indices=d.integer_indices); t_pdhg = time.perf_counter() - t
arm = bench.CpuArm(d, dict(col_basis=root.col_basis, row_basis=root.row_basis), cores)
arm.run(deltas[:cores])
ref, t_cpu = arm.run(deltas)
arm.close()
t = time.perf_counter()
for dl in deltas:
    l, u = d.l.copy(), d.u.copy(); l[dl[0][0]], u[dl[0][0]] = dl[0][1], dl[0][2]
    h.set_col_bounds(l, u); h.set_basis(root.col_basis, root.row_basis); h.solve()
t_cpu1 = time.perf_counter() - t
err = max(abs(a - b[1]) / max(1, abs(b[1])) for a, b, s in zip(r.objective, ref, r.status) if s == 0 and b[0] == 0)
out['config3'] = dict(lps=len(deltas), simplex=sx, pdhg_s=t_pdhg, gpu_s=t_gpu, gpu_lps_per_s=len(deltas) / t_gpu, gpu_root_s=t_root_gpu, cpu_root_s=t_root_cpu,
                      cpu_pool_s=t_cpu, cpu_pool_lps_per_s=len(deltas) / t_cpu, cpu_cores=cores, cpu_1core_s=t_cpu1,
                      cpu_1core_lps_per_s=len(deltas) / t_cpu1, max_rel_obj_err=err, status_match=bool(all(int(s) == b[0] for s, b in zip(r.status, ref))),
                      gpu_pivots_max=int(r.pivots.max()), kernel_launches=r.stats['kernel_launches'])
print(out['config3'], flush=True)
lp.close()

# ---------------- config 4
d, depth, root4 = bench.load_instance('c4')
B = 256
lbs, ubs, dl = frontier_nodes(d, root4['x'], 0, B, depth, seed=0)
lp = engine.BatchLP(d.A, d.b, d.c)
rng = np.random.default_rng(5)
x0 = np.tile(root4['x'], (B, 1)); y0 = np.tile(root4['y'], (B, 1))
lp.solve_batch(lbs[:8], ubs[:8], x0=x0[:8], y0=y0[:8], opts=engine.default_opts(max_iters=128))
t = time.perf_counter(); res = lp.solve_batch(lbs, ubs, x0=x0, y0=y0); t0 = time.perf_counter() - t
rounds = [dict(rows=0, gpu_s=t0, lps_per_s=B / t0, iters_mean=float(res.iterations.mean()), unsolved=int((res.status == 3).sum()))]
masks = np.zeros((B, 0), dtype=np.uint8)
cuts, rhs = [], []
for rnd in range(3):
    new_rows, new_rhs = [], []
    for k in range(32):
        S = rng.choice(d.n, size=400, replace=False)
        row = np.zeros(d.n); row[S] = -1.0
        new_rows.append(row); new_rhs.append(-np.floor(res.x[k % B][S].sum()))
    t = time.perf_counter(); lp.append_rows(np.array(new_rows), np.array(new_rhs)); t_app = time.perf_counter() - t
    cuts += new_rows; rhs += new_rhs
    masks = np.hstack([masks, (rng.random((B, 32)) < 0.25).astype(np.uint8)])
    t = time.perf_counter()
    res = lp.solve_batch(lbs, ubs, row_mask=masks, x0=res.x, y0=np.hstack([res.y, np.zeros((B, 32))]))
    dt = time.perf_counter() - t
    rounds.append(dict(rows=32 * (rnd + 1), append_s=t_app, gpu_s=dt, lps_per_s=B / dt, iters_mean=float(res.iterations.mean()),
                       unsolved=int((res.status == 3).sum())))
    print(rounds[-1], flush=True)
# CPU: the final-round LPs of `cores` nodes, warm from the root basis (+ basic slack for the cut rows)
def cpu_node(k):
    hh = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), lbs[k], ubs[k])
    on = np.flatnonzero(masks[k])
    for tt in on:
        hh.add_row(cuts[tt], rhs[tt])
    hh.set_basis(root4['col_basis'], np.concatenate([root4['row_basis'], np.ones(len(on), dtype=np.int32)]))
    t = time.perf_counter(); s = hh.solve(); return s.status, s.objective, time.perf_counter() - t
import multiprocessing as mp
with mp.get_context('fork').Pool(cores) as pool:
    t = time.perf_counter(); refs = pool.map(cpu_node, range(cores)); t_cpu = time.perf_counter() - t
err = max(abs(res.objective[k] - refs[k][1]) / max(1, abs(refs[k][1])) for k in range(cores) if refs[k][0] == 0 and res.status[k] == 0)
out['config4'] = dict(batch=B, rounds=rounds, cpu_sample_nodes=cores, cpu_cores=cores, cpu_pool_s=t_cpu, cpu_lps_per_s=cores / t_cpu,
                      cpu_mean_s_per_lp=float(np.mean([r[2] for r in refs])), max_rel_obj_err_last_round=err,
                      status_match=bool(all(int(res.status[k]) == refs[k][0] for k in range(cores))))
print(out['config4'], flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/configs.json', 'w'), indent=1)
