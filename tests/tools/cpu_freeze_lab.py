"""TEST INFRASTRUCTURE — CPU laboratory for FREEZING settled coordinates of the PDHG iteration.

Same iteration as ``cpu_pdhg_lab.solve1`` (the device's), for a small LOCK-STEP batch of nodes (columns of X, Y), plus
the device's freezing rule: a column that rests at a bound with a reduced cost of the right sign and a safe margin, and a
row with zero multiplier and a safe slack, are not updated (the full iteration would not move them either); the margins
are re-checked at every evaluation on the FULL problem, with hysteresis, and a set is shared by all nodes of a tile (a
coordinate is frozen only if every node of the tile agrees). Reports iterations and the share of coordinate updates
that were skipped, i.e. the HBM stream the device would save.

    python tests/tools/cpu_freeze_lab.py c5 0:8 0.25 0.5      # nodes 0..7 as one tile, margins (unfreeze, freeze)
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


def solve_tile(P, lb, ub, x0, y0, eps=1e-7, max_iters=150000, K=64, theta=0.05, art=0.36, suff=0.2, nec=0.8,
               long_after=32, balance=0.3, bal_dead=0.25, m_lo=None, m_hi=None, log=None, sub_step=0.0, imb=0.0, imb_min=0.1):
    """lb, ub: [B, n]. Returns per-node iteration counts and the skipped share. m_lo/m_hi None: no freezing.
    sub_step > 0: the step is sub_step / ||A_UU||, A_UU = the rows and columns that are not frozen (power iteration,
    re-estimated at every change of the sets), instead of 0.998 / ||A||."""
    B = lb.shape[0]
    n, m = P.n, P.m
    l = (lb / P.dc * P.sb).T.copy()
    u = (ub / P.dc * P.sb).T.copy()
    A, AT, b, c = P.A, P.AT, P.b[:, None], P.c[:, None]
    x = np.clip(np.repeat((x0 / P.dc * P.sb)[:, None], B, 1), l, u)
    y = np.repeat(np.maximum(y0 / P.dr * P.sc, 0)[:, None], B, 1)
    xa, ya = x.copy(), y.copy()
    omega = np.full(B, P.omega0)
    eta = P.eta
    t = np.zeros(B, dtype=np.int64)
    fpe0 = np.full(B, INF)
    fpe_prev = np.full(B, INF)
    rowscale = (1.0 / (P.dr * P.sb))[:, None]
    objscale = 1.0 / (P.sb * P.sc)
    total = periods = 0
    done = np.zeros(B, dtype=bool)
    iters = np.zeros(B, dtype=np.int64)
    obj = np.zeros(B)
    fc = np.zeros(n, dtype=bool)        # tile-wide frozen columns / rows
    fr = np.zeros(m, dtype=bool)
    work = skipped = 0.0
    runaway = np.zeros(B, dtype=np.int64)
    freeze = m_lo is not None
    cabs = np.abs(P.c)[:, None] + 1e-300

    def margins(xp, yp, gp, axp, xi, yi):
        # signed safety margin of a column: how far its reduced cost is from letting it move; only for a coordinate
        # whose iterate, anchor and T(z) all rest on the bound (then the full iteration reproduces it exactly)
        r = c - gp
        colm = np.where(xp <= l, r, np.where(xp >= u, -r, -INF))
        colm = np.where((xi == xp) & (xa == xp), colm, -INF)
        colm = np.where(l >= u, np.where((xi == l) & (xa == l), INF, -INF), colm)
        rowm = np.where((yp <= 0.0) & (yi == 0.0) & (ya == 0.0), axp - b, -INF)
        return colm, rowm

    def update_sets(colm, rowm, cscale, rscale, live):
        nonlocal fc, fr
        cm = colm[:, live].min(axis=1) / cscale
        rm = rowm[:, live].min(axis=1) / rscale
        fc = np.where(fc, cm > m_lo, cm > m_hi)
        fr = np.where(fr, rm > m_lo, rm > m_hi)

    pv = np.random.default_rng(0).standard_normal(n)

    def sub_norm(iters):
        nonlocal pv
        keep_c, keep_r = ~fc, ~fr
        v = pv * keep_c
        v /= max(np.linalg.norm(v), 1e-300)
        s = 0.0
        for _ in range(iters):
            w = (A @ v) * keep_r
            v = (AT @ w) * keep_c
            s = np.sqrt(np.linalg.norm(v))
            v /= max(np.linalg.norm(v), 1e-300)
        pv = v
        return s

    def new_eta(iters):
        if sub_step <= 0:
            return P.eta
        return min(sub_step / max(sub_norm(iters), 1e-300), 4.0 * P.eta)

    cscale = rscale = 1.0
    if freeze:
        g0 = AT @ y
        ax0 = A @ x
        cscale = np.median(np.abs(c - g0)[np.abs(c - g0) > 1e-12])
        sl0 = (ax0 - b)
        rscale = np.median(sl0[sl0 > 1e-12])
        colm, rowm = margins(x, y, g0, ax0, x, y)
        update_sets(colm, rowm, cscale, rscale, ~done)
        eta = new_eta(60)
        if log:
            log(f'scales col {cscale:.3e} row {rscale:.3e}; frozen at start: cols {fc.mean():.3f} rows {fr.mean():.3f} '
                f'eta x{eta / P.eta:.3f}')
    while total < max_iters and not done.all():
        Kp = K if periods < long_after else 4 * K
        live = ~done
        for it in range(Kp):
            tau, sig = eta / omega, eta * omega
            w = (t + 1.0) / (t + 2.0)
            g = AT @ y
            xp = np.clip(x - tau * (c - g), l, u)
            if freeze:
                xp[fc] = x[fc]
            xbar = 2 * xp - x
            yp = np.maximum(y + sig * (b - A @ xbar), 0)
            if freeze:
                yp[fr] = y[fr]
            xn = w * (2 * xp - x) + (1 - w) * xa
            yn = w * (2 * yp - y) + (1 - w) * ya
            nl = live.sum()
            work += nl * (n + m)
            skipped += nl * (fc.sum() + fr.sum())
            if it == Kp - 1:
                dx, dy = xp - x, yp - y
                gp = AT @ yp
                axp = A @ xp
                fpe = np.sqrt(np.maximum((dx * dx).sum(0) / tau + (dy * dy).sum(0) / sig + 2 * (dx * (gp - g)).sum(0), 0))
                r = c - gp
                dobj = ((b * yp).sum(0) + (np.maximum(r, 0) * l + np.minimum(r, 0) * u).sum(0)) * objscale
                pobj = (c * xp).sum(0) * objscale
                rp = np.linalg.norm(np.maximum(b - axp, 0) * rowscale, axis=0) / (1 + P.bnorm0)
                rg = np.abs(pobj - dobj) / (1 + np.abs(pobj) + np.abs(dobj))
                tot_it = total + it + 1
                fin = live & (rp <= eps) & (rg <= eps)
                iters[fin] = tot_it
                obj[fin] = pobj[fin]
                done |= fin
                for k in np.nonzero(~done)[0]:
                    if np.isfinite(fpe0[k]) and fpe[k] > 1.5 * min(fpe0[k], fpe_prev[k]):
                        runaway[k] += 1            # the device's watchdog would send the node back to 1 / ||A|| here
                    why = fpe[k] <= suff * fpe0[k] or (fpe[k] <= nec * fpe0[k] and fpe[k] > fpe_prev[k]) \
                        or t[k] + 1 >= art * tot_it or not np.isfinite(fpe0[k])
                    # experiment: restart early when one criterion lags the other by more than a factor exp(imb)
                    if not why and imb > 0 and rp[k] > 0 and rg[k] > 0 and t[k] + 1 >= imb_min * tot_it \
                            and abs(np.log(rp[k] / rg[k])) > imb and max(rp[k], rg[k]) > eps:
                        why = True
                    if why:
                        ddx, ddy = np.linalg.norm(xp[:, k] - xa[:, k]), np.linalg.norm(yp[:, k] - ya[:, k])
                        if np.isfinite(fpe0[k]):
                            if ddx > 1e-10 and ddy > 1e-10:
                                omega[k] = np.exp(theta * np.log(ddy / ddx) + (1 - theta) * np.log(omega[k]))
                            if balance > 0 and rp[k] > 0 and rg[k] > 0:
                                lr = np.log(rp[k] / rg[k])
                                lr = np.sign(lr) * max(abs(lr) - bal_dead, 0.0)
                                omega[k] *= np.exp(balance * np.clip(lr, -1.0, 1.0))
                        xn[:, k], yn[:, k] = xp[:, k], yp[:, k]
                        xa[:, k], ya[:, k] = xp[:, k], yp[:, k]
                        fpe0[k], fpe_prev[k], t[k] = fpe[k], INF, -1
                    else:
                        fpe_prev[k] = fpe[k]
                if freeze:
                    colm, rowm = margins(xp, yp, gp, axp, xn, yn)   # after the restarts, as on the device
                    if (~done).any():
                        update_sets(colm, rowm, cscale, rscale, ~done)
                        eta = new_eta(6)
                if log and (periods % 16 == 0):
                    log(f'{tot_it:7d} live {int((~done).sum())} eta x{eta / P.eta:.3f} frozen cols {fc.mean():.3f} rows {fr.mean():.3f} '
                        f'rp {rp[~done].max() if (~done).any() else 0:.2e} rg {rg[~done].max() if (~done).any() else 0:.2e}')
            x, y = xn, yn
            t = t + 1
        total += Kp
        periods += 1
    iters[~done] = total
    return dict(iters=iters, obj=obj, skipped=skipped / max(work, 1.0), done=done, runaway=runaway)


def main():
    wl = sys.argv[1]
    k0, k1 = (int(a) for a in sys.argv[2].split(':'))
    m_lo = float(sys.argv[3]) if len(sys.argv) > 3 else None
    m_hi = float(sys.argv[4]) if len(sys.argv) > 4 else None
    sub_step = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
    imb = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
    imb_min = float(sys.argv[7]) if len(sys.argv) > 7 else 0.1
    d, depth, root = bench.load_instance(wl)
    P = BatchPDHG(d.A, d.b, d.c)
    lb, ub, _ = frontier_nodes(d, root['x'], k0, k1 - k0, depth, seed=0)
    t0 = time.time()
    r = solve_tile(P, lb, ub, root['x'], root['y'], m_lo=m_lo, m_hi=m_hi, sub_step=sub_step, imb=imb, imb_min=imb_min, log=None)
    print('margins', m_lo, m_hi, 'sub_step', sub_step, 'imb', imb, imb_min, 'max', int(r['iters'].max()), 'iters', r['iters'].tolist(), 'mean %.0f' % r['iters'].mean(), 'skipped %.3f' % r['skipped'],
          'runaway', r['runaway'].tolist(), 'obj', np.round(r['obj'], 6).tolist(), 'time %.0f' % (time.time() - t0))


if __name__ == '__main__':
    main()
