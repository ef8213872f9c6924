"""TEST INFRASTRUCTURE — numpy model of the batched restarted Halpern PDHG the CUDA kernels run.

Not the product and never imported by it: only ``tests/`` use this file, to localise a
divergence between the device kernels (simple_mip_solver_b200/csrc/blp_kernels.cu) and the
algorithm they are meant to implement. The exact-LP oracle for parity is ``oracle/highs_lp.py``.

Problem (the canonical node LP of the reference, base_node.py:259-286):

    min c.x  s.t.  A x >= b (row-masked), l_k <= x <= u_k        for every node k of a batch

Vectors are stored node-fastest, X[n, B], exactly as on the device. One deliberate difference: the
device rounds the Halpern anchors to fp32 at every restart (any fixed anchor is a valid Halpern
anchor); this model keeps them in fp64.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

INF = float('inf')


def ruiz_pc_scaling(A: sp.csr_matrix, ruiz_iters=10):
    """Ruiz equilibration followed by one Pock-Chambolle (alpha=1) pass. Returns (Dr, Dc)."""
    m, n = A.shape
    dr = np.ones(m)
    dc = np.ones(n)
    absA = abs(A).tocsr().astype(float)
    cur = absA.copy()
    for _ in range(ruiz_iters):
        rmax = np.asarray(cur.max(axis=1).todense()).ravel() if cur.nnz else np.zeros(m)
        cmax = np.asarray(cur.max(axis=0).todense()).ravel() if cur.nnz else np.zeros(n)
        sr = np.where(rmax > 0, 1.0 / np.sqrt(np.where(rmax > 0, rmax, 1.0)), 1.0)
        sc = np.where(cmax > 0, 1.0 / np.sqrt(np.where(cmax > 0, cmax, 1.0)), 1.0)
        dr *= sr
        dc *= sc
        cur = sp.diags(sr) @ cur @ sp.diags(sc)
    rsum = np.asarray(cur.sum(axis=1)).ravel()
    csum = np.asarray(cur.sum(axis=0)).ravel()
    sr = np.where(rsum > 0, 1.0 / np.sqrt(np.where(rsum > 0, rsum, 1.0)), 1.0)
    sc = np.where(csum > 0, 1.0 / np.sqrt(np.where(csum > 0, csum, 1.0)), 1.0)
    return dr * sr, dc * sc


def power_iteration_norm(A: sp.csr_matrix, iters=60, seed=1):
    m, n = A.shape
    if A.nnz == 0:
        return 1.0
    rng = np.random.Generator(np.random.PCG64(seed))
    v = rng.standard_normal(n)
    v /= np.linalg.norm(v)
    s = 1.0
    for _ in range(iters):
        w = A @ v
        v = A.T @ w
        nv = np.linalg.norm(v)
        if nv == 0:
            return 1.0
        s = np.sqrt(nv)
        v /= nv
    return s


class BatchPDHG:
    def __init__(self, A, b, c, ruiz_iters=10, bound_obj_rescale=True):
        A = sp.csr_matrix(A, dtype=float)
        self.m, self.n = A.shape
        self.A0, self.b0, self.c0 = A, np.asarray(b, float), np.asarray(c, float)
        self.dr, self.dc = ruiz_pc_scaling(A, ruiz_iters)
        As = (sp.diags(self.dr) @ A @ sp.diags(self.dc)).tocsr()
        bs = self.b0 * self.dr
        cs = self.c0 * self.dc
        if bound_obj_rescale:
            self.sb = 1.0 / (np.linalg.norm(bs) + 1.0)
            self.sc = 1.0 / (np.linalg.norm(cs) + 1.0)
        else:
            self.sb = self.sc = 1.0
        self.A = As
        self.AT = As.T.tocsr()
        self.b = bs * self.sb
        self.c = cs * self.sc
        self.eta = 0.998 / power_iteration_norm(As)
        nb, nc = np.linalg.norm(self.b), np.linalg.norm(self.c)
        self.omega0 = nc / nb if nb > 1e-12 and nc > 1e-12 else 1.0
        self.bnorm0 = np.linalg.norm(self.b0)
        self.cnorm0 = np.linalg.norm(self.c0)

    def solve(self, lb, ub, eps=1e-8, max_iters=200000, K=64, row_mask=None, x0=None, y0=None,
              reflect=True, restart_to='pdhg', theta=0.5, verbose=False, eps_inf=1e-9, omega_init=None, balance=0.3, balance_dead=0.25):
        """lb, ub: [n, B] (unscaled). Returns dict of per-node arrays."""
        n, m = self.n, self.m
        lb = np.asarray(lb, float).reshape(n, -1)
        ub = np.asarray(ub, float).reshape(n, -1)
        B = lb.shape[1]
        l = lb / self.dc[:, None] * self.sb
        u = ub / self.dc[:, None] * self.sb
        fin_l, fin_u = np.isfinite(l), np.isfinite(u)
        A, AT, b, c = self.A, self.AT, self.b[:, None], self.c[:, None]
        mask = np.ones((m, B)) if row_mask is None else np.asarray(row_mask, float).reshape(m, B)
        x = np.clip(np.zeros((n, B)) if x0 is None else x0 / self.dc[:, None] * self.sb, l, u)
        y = np.zeros((m, B)) if y0 is None else np.maximum(y0 / self.dr[:, None] * self.sc, 0)
        y *= mask
        xa, ya = x.copy(), y.copy()
        omega = np.full(B, self.omega0) if omega_init is None else np.broadcast_to(np.asarray(omega_init, float), (B,)).copy()
        t = np.zeros(B)                       # iterations since restart, per node
        fpe0 = np.full(B, INF)
        fpe_prev = np.full(B, INF)
        active = np.ones(B, bool)
        status = np.full(B, 3)                # CLP codes: 0 opt, 1 primal infeasible, 2 dual inf, 3 limit
        out_x = np.zeros((n, B)); out_y = np.zeros((m, B))
        out_obj = np.full(B, np.nan); out_dobj = np.full(B, np.nan); iters = np.zeros(B, int)
        total = 0
        rowscale = 1.0 / (self.dr * self.sb)        # scaled primal residual -> unscaled
        colscale = 1.0 / (self.dc * self.sc)        # scaled dual residual -> unscaled
        objscale = 1.0 / (self.sb * self.sc)
        while total < max_iters and active.any():
            for it in range(K):
                major = it == K - 1
                tau = (self.eta / omega)[None, :]
                sig = (self.eta * omega)[None, :]
                w = ((t + 1) / (t + 2))[None, :]
                g = AT @ y
                xp = np.clip(x - tau * (c - g), l, u)
                xbar = 2 * xp - x
                yp = np.maximum(y + sig * (b - A @ xbar), 0) * mask
                if reflect:
                    xn = w * xbar + (1 - w) * xa
                    yn = w * (2 * yp - y) + (1 - w) * ya
                else:
                    xn = w * xp + (1 - w) * xa
                    yn = w * yp + (1 - w) * ya
                if major:
                    dx, dy = xp - x, yp - y
                    gp = AT @ yp
                    Axp = A @ xp
                    # fixed point error in the M norm (scaled space)
                    fpe2 = (dx * dx).sum(0) / tau[0] + (dy * dy).sum(0) / sig[0] \
                        + 2 * (dx * (gp - g)).sum(0)
                    fpe = np.sqrt(np.maximum(fpe2, 0))
                    # KKT in the unscaled space
                    pres = np.maximum(b - Axp, 0) * mask * rowscale[:, None]
                    r = c - gp
                    lam = np.where(r > 0, np.where(fin_l, r, 0), np.where(fin_u, r, 0))
                    dres = (r - lam) * colscale[:, None]
                    pobj = (c * xp).sum(0) * objscale
                    lterm = np.where(fin_l, l, 0) * np.maximum(lam, 0)
                    uterm = np.where(fin_u, u, 0) * np.minimum(lam, 0)
                    dobj = ((b * yp).sum(0) + (lterm + uterm).sum(0)) * objscale
                    rp = np.linalg.norm(pres, axis=0) / (1 + self.bnorm0)
                    rd = np.linalg.norm(dres, axis=0) / (1 + self.cnorm0)
                    rg = np.abs(pobj - dobj) / (1 + np.abs(pobj) + np.abs(dobj))
                    conv = (rp <= eps) & (rd <= eps) & (rg <= eps) & active
                    # primal infeasibility certificate from the dual iterate: b.y - max_box (A^T y).x > 0
                    gpos, gneg = np.maximum(gp, 0), np.minimum(gp, 0)
                    box = np.where(gpos > 0, gpos * np.where(fin_u, u, INF), 0) \
                        + np.where(gneg < 0, gneg * np.where(fin_l, l, -INF), 0)
                    with np.errstate(invalid='ignore'):
                        farkas = (b * yp).sum(0) - box.sum(0)
                    ynorm = np.abs(yp).sum(0) * np.abs(b).max() + 1e-300
                    infeas = (farkas > eps_inf * (np.abs(b * yp).sum(0) + np.abs(box).sum(0) + 1e-300)) \
                        & (farkas > 0) & active & ~conv
                    # the same certificate from the dual step dy = y' - y (offset-free ray estimate)
                    gd = gp - g
                    dpos, dneg = np.maximum(gd, 0), np.minimum(gd, 0)
                    dbox = np.where(dpos > 0, dpos * np.where(fin_u, u, INF), 0) \
                        + np.where(dneg < 0, dneg * np.where(fin_l, l, -INF), 0)
                    dymax = np.abs(dy).max(0)
                    with np.errstate(invalid='ignore'):
                        fstep = (b * dy).sum(0) - dbox.sum(0)
                        step_ok = np.isfinite(dbox).all(0) & (dymax > 0) \
                            & (fstep > 1e-6 * (np.abs(b * dy).sum(0) + np.abs(np.where(np.isfinite(dbox), dbox, 0)).sum(0))) \
                            & ((-dy).max(0) <= 1e-8 * dymax)
                    infeas |= step_ok & active & ~conv
                    # dual infeasibility (unbounded) certificate from the primal iterate
                    xnorm = np.abs(xp).max(0) + 1e-300
                    d = xp / xnorm
                    ray_ok = ((np.minimum(A @ d, 0) * mask) ** 2).sum(0) <= (1e-9) ** 2
                    ray_ok &= ((np.where(fin_l, np.minimum(d, 0), 0) ** 2).sum(0) <= 1e-18)
                    ray_ok &= ((np.where(fin_u, np.maximum(d, 0), 0) ** 2).sum(0) <= 1e-18)
                    unb = ray_ok & ((c * d).sum(0) < -1e-9) & (xnorm > 1e6) & active & ~conv
                    done = conv | infeas | unb
                    for k in np.where(done)[0]:
                        status[k] = 0 if conv[k] else (1 if infeas[k] else 2)
                        out_x[:, k] = xp[:, k] * self.dc / self.sb
                        out_y[:, k] = yp[:, k] * self.dr / self.sc
                        out_obj[k] = pobj[k]; out_dobj[k] = dobj[k]
                        iters[k] = total + it + 1
                    active &= ~done
                    if verbose:
                        print(total + it + 1, 'active', active.sum(), 'rp %.2e rd %.2e rg %.2e' %
                              (rp[active].max(initial=0), rd[active].max(initial=0),
                               rg[active].max(initial=0)))
                    # restart decision
                    tt = t + 1
                    do_restart = (fpe <= 0.2 * fpe0) | ((fpe <= 0.8 * fpe0) & (fpe > fpe_prev)) \
                        | (tt >= 0.36 * (total + it + 1)) | ~np.isfinite(fpe0)
                    fpe_prev = np.where(do_restart, INF, fpe)
                    if restart_to == 'pdhg':
                        zx, zy = xp, yp
                    else:
                        zx, zy = xn, yn
                    ddx = np.linalg.norm(zx - xa, axis=0)
                    ddy = np.linalg.norm(zy - ya, axis=0)
                    good = do_restart & (ddx > 1e-10) & (ddy > 1e-10) & np.isfinite(fpe0)
                    omega = np.where(good, np.exp(theta * np.log(np.where(good, ddy / np.maximum(ddx, 1e-300), 1))
                                                  + (1 - theta) * np.log(omega)), omega)
                    # residual balancing (k_decide, BLP_OMEGA_BALANCE): push the weight towards the lagging
                    # criterion — the primal residual shrinks with the dual step, the gap with the primal step
                    if balance > 0:
                        with np.errstate(divide='ignore', invalid='ignore'):
                            lr = np.log(rp / rg)
                            lr = np.sign(lr) * np.clip(np.abs(lr) - balance_dead, 0.0, 1.0)
                        fb = omega * np.exp(balance * np.where(np.isfinite(lr), lr, 0.0))
                        ok = do_restart & np.isfinite(fpe0) & (rp > 0) & (rg > 0) \
                            & (fb <= 1e4 * self.omega0) & (fb >= 1e-4 * self.omega0)
                        omega = np.where(ok, fb, omega)
                    rs = do_restart[None, :]
                    xn = np.where(rs, zx, xn); yn = np.where(rs, zy, yn)
                    xa = np.where(rs, zx, xa); ya = np.where(rs, zy, ya)
                    fpe0 = np.where(do_restart, fpe, fpe0)
                    t = np.where(do_restart, -1.0, t)
                x, y = xn, yn
                t = t + 1
            total += K
        for k in np.where(active)[0]:
            out_x[:, k] = xp[:, k] * self.dc / self.sb
            out_y[:, k] = yp[:, k] * self.dr / self.sc
            out_obj[k] = pobj[k]; out_dobj[k] = dobj[k]; iters[k] = total
        return dict(status=status, obj=out_obj, dobj=out_dobj, x=out_x, y=out_y, iters=iters, omega=omega)
