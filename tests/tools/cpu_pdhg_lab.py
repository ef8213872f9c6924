"""TEST INFRASTRUCTURE — CPU laboratory for the PDHG iteration the CUDA kernels run.

A lean single-node restatement (1-D vectors, ~0.5 ms per iteration at the C4 shape, ~2 ms at C5) of
``oracle/pdhg_numpy.BatchPDHG.solve`` with every constant of ``k_decide`` exposed, so that changes
to the restart rule / primal weight / preconditioner are screened on the host cores before they
cost GPU minutes. Uses the oracle's scaling and the bench fixtures; never imported by the product.

    python tests/tools/cpu_pdhg_lab.py c4 0:8 base frozen tb0.3      # nodes 0..7, three variants
    python tests/tools/cpu_pdhg_lab.py c5 1:2 trace                  # convergence trace of node 1

Variants: base (the shipped rule: smoothing .05, balance .3, dead zone .25) | nobal (balance off) |
frozen (weight never moves) | theta<v> (smoothing v, no balance) | tb<k> (balance k, no dead zone) |
bal<k> (same with theta 0) | dz<d>_<k> (balance k, dead zone d) | rho<v> (reflection) |
ex<b> (restart point extrapolated by b) | eta<f> (per-node step up to f/||A||) | hs<c> (Halpern weight (k+c)/(k+c+1)) | art<v> (artificial restart constant) |
om<f> (frozen weight f x omega0) | cold (no warm start) | trace (base, with the convergence trace)

Findings of round 1: on the C4/C5 frontiers the iteration is in its
sublinear O(1/k) regime — the primal objective is right to 1e-8 after ~3 k iterations, the other
~20 k only close the duality gap and the primal residual, both at rate 1/k; a warm start from the
root optimum is forgotten within ~8 k iterations; Ruiz(10)+Pock-Chambolle is the best of seven
scalings; reflection 1.0 and the artificial-restart constant 0.36 are at their optimum; implied
bound tightening shrinks the gap by 30 % at a fixed y but does not shorten the run (the primal
residual binds); a primal weight that is pushed towards the lagging criterion at restarts
(``BLP_OMEGA_BALANCE``) leaves the mean iteration count alone and cuts the slowest nodes by 12-20 %.
Iteration counts of this model match the device's (config 3, 128 children: mean 11 504 / max 30 976
without balancing on both; 11 056 / 34 816 here against 11 052 / 34 816 on the GPU with balance .3).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                                                    # noqa: E402
from oracle.pdhg_numpy import BatchPDHG                         # noqa: E402
from simple_mip_solver_b200.instances import frontier_nodes     # noqa: E402

INF = float('inf')


def solve1(P, lb, ub, x0=None, y0=None, eps=1e-7, max_iters=150000, K=64, theta=0.05, omega_init=None,
           art=0.36, suff=0.2, nec=0.8, trace=False, long_after=32, balance=0.3, bal_clip=1.0,
           bal_dead=0.25, rho=1.0, extrap=0.0, eta_max=1.0, eta_safety=0.9, hshift=1.0):
    """One node, the device algorithm: reflected Halpern PDHG, evaluation every K (4K after
    ``long_after`` periods) iterations, restart to T(z), primal weight updated at restarts."""
    n, m = P.n, P.m
    l = lb / P.dc * P.sb
    u = ub / P.dc * P.sb
    A, AT, b, c = P.A, P.AT, P.b, P.c
    x = np.clip(np.zeros(n) if x0 is None else x0 / P.dc * P.sb, l, u)
    y = np.zeros(m) if y0 is None else np.maximum(y0 / P.dr * P.sc, 0)
    xa, ya = x.copy(), y.copy()
    omega = P.omega0 if omega_init is None else omega_init
    eta = eta0 = P.eta
    eta_resets = 0
    t = 0
    fpe0 = fpe_prev = INF
    rowscale = 1.0 / (P.dr * P.sb)
    objscale = 1.0 / (P.sb * P.sc)
    total = periods = 0
    hist = []
    pobj = dobj = np.nan
    while total < max_iters:
        Kp = K if periods < long_after else 4 * K
        for it in range(Kp):
            tau, sig = eta / omega, eta * omega
            w = (t + hshift) / (t + hshift + 1)          # Halpern weight; hshift 1 is the device's
            g = AT @ y
            xp = np.clip(x - tau * (c - g), l, u)
            xbar = 2 * xp - x
            yp = np.maximum(y + sig * (b - A @ xbar), 0)
            xn = w * ((1 + rho) * xp - rho * x) + (1 - w) * xa
            yn = w * ((1 + rho) * yp - rho * y) + (1 - w) * ya
            if it == Kp - 1:
                dx, dy = xp - x, yp - y
                gp = AT @ yp
                fpe = np.sqrt(max(dx @ dx / tau + dy @ dy / sig + 2 * dx @ (gp - g), 0))
                r = c - gp
                dobj = (b @ yp + (np.maximum(r, 0) * l + np.minimum(r, 0) * u).sum()) * objscale
                pobj = c @ xp * objscale
                rp = np.linalg.norm(np.maximum(b - A @ xp, 0) * rowscale) / (1 + P.bnorm0)
                rg = abs(pobj - dobj) / (1 + abs(pobj) + abs(dobj))
                tot_it = total + it + 1
                if trace:
                    hist.append([tot_it, rp, rg, pobj, dobj, fpe, omega, t + 1, ''])
                if rp <= eps and rg <= eps:
                    return dict(iters=tot_it, obj=pobj, dobj=dobj, hist=hist, omega=omega, eta=eta / eta0,
                                eta_resets=eta_resets,
                                x=xp * P.dc / P.sb, y=yp * P.dr / P.sc)
                cross = dx @ (gp - g)
                if eta_max > 1.0 and eta > eta0 and fpe > 1.5 * min(fpe_prev, fpe0):
                    eta, eta_resets = eta0, eta_resets + 1      # the larger step lost contraction: back to 1/||A||
                    fpe0 = INF                                   # ... and restart from here
                why = 's' if fpe <= suff * fpe0 else 'n' if (fpe <= nec * fpe0 and fpe > fpe_prev) \
                    else 'a' if t + 1 >= art * tot_it else 'i' if not np.isfinite(fpe0) else ''
                if why:
                    ddx, ddy = np.linalg.norm(xp - xa), np.linalg.norm(yp - ya)
                    if np.isfinite(fpe0):
                        if ddx > 1e-10 and ddy > 1e-10:
                            omega = np.exp(theta * np.log(ddy / ddx) + (1 - theta) * np.log(omega))
                        if balance > 0 and rp > 0 and rg > 0:      # k_decide: BLP_OMEGA_BALANCE
                            lr = np.log(rp / rg)
                            lr = np.sign(lr) * max(abs(lr) - bal_dead, 0.0)
                            omega *= np.exp(balance * np.clip(lr, -bal_clip, bal_clip))
                    if eta_max > 1.0 and np.isfinite(fpe0) and abs(cross) > 0:
                        # per-node step: PDLP's bound ||dz||^2_omega / (2 |dx.A'dy|) >= 1/||A||,
                        # evaluated on the last step of the phase; the iteration only feels the active block
                        lim = eta_safety * (omega * (dx @ dx) + (dy @ dy) / omega) / (2 * abs(cross))
                        eta = float(np.clip(lim, eta0, eta_max * eta0))
                    if extrap > 0 and np.isfinite(fpe0):       # extrapolated restart point (tried: see DESIGN)
                        zx = np.clip(xp + extrap * (xp - xa), l, u)
                        zy = np.maximum(yp + extrap * (yp - ya), 0)
                    else:
                        zx, zy = xp, yp
                    xn, yn = zx, zy
                    xa, ya = zx.copy(), zy.copy()
                    fpe0, fpe_prev, t = fpe, INF, -1
                    if trace:
                        hist[-1][-1] = why
                else:
                    fpe_prev = fpe
            x, y = xn, yn
            t += 1
        total += Kp
        periods += 1
    return dict(iters=total, obj=pobj, dobj=dobj, hist=hist, omega=omega, eta=eta / eta0, eta_resets=eta_resets,
                x=xp * P.dc / P.sb, y=yp * P.dr / P.sc)


def variant_kwargs(v, P):
    off = dict(balance=0.0, bal_dead=0.0)
    if v in ('base', 'trace', 'cold'):
        return {}
    if v == 'nobal':
        return off
    if v == 'frozen':
        return dict(off, theta=0.0)
    for prefix, make in (('theta', lambda s: dict(off, theta=float(s))),
                         ('tb', lambda s: dict(balance=float(s), bal_dead=0.0)),
                         ('bal', lambda s: dict(theta=0.0, balance=float(s), bal_dead=0.0)),
                         ('dz', lambda s: dict(bal_dead=float(s.split('_')[0]), balance=float(s.split('_')[1]))),
                         ('rho', lambda s: dict(rho=float(s))),
                         ('ex', lambda s: dict(extrap=float(s))),
                         ('art', lambda s: dict(art=float(s))),
                         ('eta', lambda s: dict(eta_max=float(s))),
                         ('hs', lambda s: dict(hshift=float(s))),
                         ('om', lambda s: dict(off, theta=0.0, omega_init=P.omega0 * float(s)))):
        if v.startswith(prefix):
            return make(v[len(prefix):])
    raise SystemExit(f'unknown variant {v}')


def main():
    wl = sys.argv[1]
    k0, k1 = (int(a) for a in sys.argv[2].split(':'))
    d, depth, root = bench.load_instance(wl)
    P = BatchPDHG(d.A, d.b, d.c)
    for k in range(k0, k1):
        lb, ub, _ = frontier_nodes(d, root['x'], k, 1, depth, seed=0)
        for v in sys.argv[3:]:
            warm = {} if v == 'cold' else dict(x0=root['x'], y0=root['y'])
            t0 = time.time()
            r = solve1(P, lb[0], ub[0], trace=v == 'trace', **warm, **variant_kwargs(v, P))
            print(v, 'node', k, 'iters', r['iters'], 'obj %.6f' % r['obj'], 'omega0 %.3f end %.3f' % (P.omega0, r['omega']), 'eta x%.2f resets %d' % (r['eta'], r['eta_resets']),
                  'time %.0f' % (time.time() - t0), flush=True)
            for h in r['hist']:
                if h[-1] or h[0] % 2048 == 0:
                    print('%7d rp %.2e rg %.2e p %.4f d %.4f fpe %.3e omega %.3f since_restart %d %s' % tuple(h))


if __name__ == '__main__':
    main()
