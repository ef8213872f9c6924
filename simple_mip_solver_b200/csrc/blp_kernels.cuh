// Device kernels of the batched node-LP bound step (sm_100a, fp64, no tensor cores).
//
// Layout: every batched vector is node-fastest, V[row][ld]; a warp owns one node tile
// (NT nodes, NT in {1,2,4,8,16,32}) and 32/NT consecutive matrix rows at a time, so every
// vector access of a warp is one contiguous 256-byte segment and the CSR entries of a row are
// warp-uniform. Work is ordered tile-major (blockIdx.y = tile, blockIdx.x = row chunk) so the
// CTAs resident at any moment share a few node tiles and the gathered vector of those tiles is
// served from L2 after its first (compulsory) read from HBM.
//
// Algorithm: reflected, restarted Halpern PDHG. With T the PDHG map
//     x' = clip(x - tau (c - A'y), l, u),   y' = max(0, y + sigma (b - A (2x' - x)))
// one step is  z+ = w (2 T(z) - z) + (1 - w) z_anchor,  w = (s+1)/(s+2), s = steps since restart.
// State kept per node column: xbar = 2x' - x (gathered by the dual step; the next primal step
// rebuilds x = w xbar + (1-w) xa from it), xa, l, u, y, ya.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace blp {

constexpr int kCtaThreads = 256;
constexpr int kWarps = kCtaThreads / 32;

// column-pass accumulators (sums first, then maxima)
enum { C_DX2, C_CROSS, C_DRES2, C_CX, C_BND, C_DXA2, C_BOX, C_BOXABS, C_CD, C_NSUM,
       C_DMAX = C_NSUM, C_DVIOL, C_RAYVIOL, C_N };
// row-pass accumulators
enum { R_PRES2, R_BY, R_DY2, R_DYA2, R_BYABS, R_NSUM, R_ADNEG = R_NSUM, R_YMAX, R_N };

struct DevProb {
    int m, m_base, n;
    const int32_t* rowptr; const int32_t* colidx; const double* val;     // scaled A   (m x n)
    const int32_t* cptr;   const int32_t* ridx;   const double* cval;    // scaled A^T (n x m)
    const double* c; const double* b;            // scaled objective / row lower bounds
    const double* rowscale; const double* colscale;   // scaled residual -> unscaled
    const double* dr; const double* dc;
    double eta, sb, sc, objscale, bnorm0, cnorm0, cinf_s, omega0;
};

struct DevState {
    int B, ld;
    double *xbar, *xa, *l, *u, *X1, *DX, *G;       // [n][ld]
    double *y, *ya, *Y1, *DY;                      // [m][ld]
    const uint8_t* rowmask;                        // [m - m_base][ld] or null
    double *omega, *fpe0, *fpe_prev, *pobj, *dobj; // [ld]
    int32_t *sbase, *fin, *status, *iters, *restart;   // [ld]
    double *partC, *partR;                         // [chunks][C_N][ld], [chunks][R_N][ld]
    int32_t* counters;                             // [0] active nodes, [1] nodes restarting
};

__device__ __forceinline__ bool is_inf(double v) { return fabs(v) >= 1e30; }

// ---------------------------------------------------------------------------------------------
// gather-dot of one CSR row with a batched vector: sum_p val[p] * V[idx[p]][node]
// NT == 32: the row is warp-uniform; lanes fetch 32 (index, value) pairs with one coalesced load
// each and broadcast them by shuffle, so the dependent index->gather chain is one load deep.
// NT < 32: every lane walks its own row.
template <int NT>
__device__ __forceinline__ double row_dot(const int32_t* __restrict__ ptr,
                                          const int32_t* __restrict__ idx,
                                          const double* __restrict__ val, int row, bool row_ok,
                                          const double* __restrict__ V, int ld, int node,
                                          bool node_ok, int lane) {
    double acc = 0.0;
    if constexpr (NT == 32) {
        const int p0 = __ldg(ptr + row), p1 = __ldg(ptr + row + 1);
        for (int base = p0; base < p1; base += 32) {
            const int cnt = min(32, p1 - base);
            int myi = 0;
            double mya = 0.0;
            if (lane < cnt) {
                myi = __ldg(idx + base + lane);
                mya = __ldg(val + base + lane);
            }
#pragma unroll 4
            for (int q = 0; q < cnt; ++q) {
                const int i = __shfl_sync(0xffffffffu, myi, q);
                const double a = __shfl_sync(0xffffffffu, mya, q);
                if (node_ok) acc = fma(a, V[(size_t)i * ld + node], acc);
            }
        }
    } else {
        if (row_ok && node_ok) {
            const int p0 = __ldg(ptr + row), p1 = __ldg(ptr + row + 1);
            for (int p = p0; p < p1; ++p)
                acc = fma(__ldg(val + p), V[(size_t)__ldg(idx + p) * ld + node], acc);
        }
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// Primal half step, fused  G = A'y  ->  x' = clip(x - tau (c - G), l, u)  ->  xbar = 2x' - x.
template <int NT, bool MAJOR>
__global__ void __launch_bounds__(kCtaThreads)
k_primal(const DevProb P, const DevState S, const int it, const int rows_per_cta) {
    constexpr int RW = 32 / NT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int node = blockIdx.y * NT + (lane % NT);
    const int sub = lane / NT;
    const bool node_ok = node < S.B && S.fin[node] == 0;
    if (__ballot_sync(0xffffffffu, node_ok) == 0) return;     // whole node tile retired
    double w = 0.0, tau = 0.0;
    if (node_ok) {
        const int s = S.sbase[node] + it;
        w = (double)s / (double)(s + 1);
        tau = P.eta / S.omega[node];
    }
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = min(P.n, r0 + rows_per_cta);
    for (int jb = r0 + warp * RW; jb < r1; jb += kWarps * RW) {
        const int j = jb + sub;
        const bool row_ok = j < r1;
        const size_t e = (size_t)j * S.ld + node;
        double xb = 0, a = 0, lo = 0, hi = 0;
        if (row_ok && node_ok) {       // issue the streaming loads before the gather
            xb = S.xbar[e];
            a = __ldcs(S.xa + e);
            lo = __ldcs(S.l + e);
            hi = __ldcs(S.u + e);
        }
        const double g = row_dot<NT>(P.cptr, P.ridx, P.cval, row_ok ? j : 0, row_ok, S.y, S.ld,
                                     node, node_ok && row_ok, lane);
        if (row_ok && node_ok) {
            const double xc = fma(w, xb - a, a);                  // w xbar + (1-w) xa
            const double xp = fmin(fmax(xc - tau * (__ldg(P.c + j) - g), lo), hi);
            S.xbar[e] = 2.0 * xp - xc;
            if constexpr (MAJOR) {
                S.X1[e] = xp;
                S.DX[e] = xp - xc;
                S.G[e] = g;
            }
        }
    }
}

// Dual half step, fused  s = A xbar  ->  y' = max(0, y + sigma (b - s))  ->  Halpern update of y.
template <int NT, bool MAJOR>
__global__ void __launch_bounds__(kCtaThreads)
k_dual(const DevProb P, const DevState S, const int it, const int rows_per_cta) {
    constexpr int RW = 32 / NT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int node = blockIdx.y * NT + (lane % NT);
    const int sub = lane / NT;
    const bool node_ok = node < S.B && S.fin[node] == 0;
    if (__ballot_sync(0xffffffffu, node_ok) == 0) return;     // whole node tile retired
    double w = 0.0, sig = 0.0;
    if (node_ok) {
        const int s = S.sbase[node] + it;
        w = (double)(s + 1) / (double)(s + 2);
        sig = P.eta * S.omega[node];
    }
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = min(P.m, r0 + rows_per_cta);
    for (int ib = r0 + warp * RW; ib < r1; ib += kWarps * RW) {
        const int i = ib + sub;
        const bool row_ok = i < r1;
        const size_t e = (size_t)i * S.ld + node;
        double yc = 0, a = 0;
        bool on = true;
        if (row_ok && node_ok) {
            yc = S.y[e];
            a = __ldcs(S.ya + e);
            if (i >= P.m_base && S.rowmask) on = S.rowmask[(size_t)(i - P.m_base) * S.ld + node] != 0;
        }
        const double ax = row_dot<NT>(P.rowptr, P.colidx, P.val, row_ok ? i : 0, row_ok, S.xbar,
                                      S.ld, node, node_ok && row_ok, lane);
        if (row_ok && node_ok) {
            const double yp = on ? fmax(0.0, yc + sig * (__ldg(P.b + i) - ax)) : 0.0;
            S.y[e] = fma(w, (2.0 * yp - yc) - a, a);
            if constexpr (MAJOR) {
                S.Y1[e] = yp;
                S.DY[e] = yp - yc;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// CTA-level, order-fixed reduction of per-thread accumulators into the chunk's partial slot.
template <int NT, int NACC, int NSUM>
__device__ __forceinline__ void cta_reduce_store(double (&acc)[NACC], double* __restrict__ part,
                                                 int ld, int node_base) {
    __shared__ double sm[NACC][kCtaThreads];
#pragma unroll
    for (int a = 0; a < NACC; ++a) sm[a][threadIdx.x] = acc[a];
    __syncthreads();
    // thread t < NT*NACC reduces accumulator (t / NT) of local node (t % NT)
    for (int t = threadIdx.x; t < NT * NACC; t += kCtaThreads) {
        const int a = t / NT, nl = t % NT;
        double r = sm[a][nl];
        for (int q = nl + NT; q < kCtaThreads; q += NT)
            r = (a < NSUM) ? r + sm[a][q] : fmax(r, sm[a][q]);
        part[((size_t)blockIdx.x * NACC + a) * ld + node_base + nl] = r;
    }
}

// Column pass of an evaluation: G' = A'y', reduced costs, dual residual, objectives, Farkas box
// term, ray direction d = x' - xa (left in G for the row pass).
template <int NT>
__global__ void __launch_bounds__(kCtaThreads)
k_eval_cols(const DevProb P, const DevState S, const int rows_per_cta) {
    constexpr int RW = 32 / NT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int node = blockIdx.y * NT + (lane % NT);
    const int sub = lane / NT;
    const bool node_ok = node < S.B && S.fin[node] == 0;
    if (__ballot_sync(0xffffffffu, node_ok) == 0) return;     // uniform over the CTA
    double acc[C_N];
#pragma unroll
    for (int a = 0; a < C_N; ++a) acc[a] = 0.0;
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = min(P.n, r0 + rows_per_cta);
    for (int jb = r0 + warp * RW; jb < r1; jb += kWarps * RW) {
        const int j = jb + sub;
        const bool row_ok = j < r1;
        const size_t e = (size_t)j * S.ld + node;
        const double gp = row_dot<NT>(P.cptr, P.ridx, P.cval, row_ok ? j : 0, row_ok, S.Y1, S.ld,
                                      node, node_ok && row_ok, lane);
        if (row_ok && node_ok) {
            const double xp = S.X1[e], dx = S.DX[e], g = S.G[e];
            const double lo = S.l[e], hi = S.u[e], xa = S.xa[e];
            const double cj = __ldg(P.c + j), cs = __ldg(P.colscale + j);
            const bool fl = !is_inf(lo), fu = !is_inf(hi);
            acc[C_DX2] = fma(dx, dx, acc[C_DX2]);
            acc[C_CROSS] = fma(dx, gp - g, acc[C_CROSS]);
            const double r = cj - gp;
            const double lam = (r > 0.0) ? (fl ? r : 0.0) : (fu ? r : 0.0);
            const double dres = (r - lam) * cs;
            acc[C_DRES2] = fma(dres, dres, acc[C_DRES2]);
            acc[C_CX] = fma(cj, xp, acc[C_CX]);
            if (lam > 0.0) acc[C_BND] = fma(lo, lam, acc[C_BND]);
            else if (lam < 0.0) acc[C_BND] = fma(hi, lam, acc[C_BND]);
            const double d = xp - xa;
            acc[C_DXA2] = fma(d, d, acc[C_DXA2]);
            double box = 0.0;
            if (gp > 0.0) { if (fu) box = gp * hi; else acc[C_RAYVIOL] = fmax(acc[C_RAYVIOL], gp); }
            else if (gp < 0.0) { if (fl) box = gp * lo; else acc[C_RAYVIOL] = fmax(acc[C_RAYVIOL], -gp); }
            acc[C_BOX] += box;
            acc[C_BOXABS] += fabs(box);
            acc[C_CD] = fma(cj, d, acc[C_CD]);
            acc[C_DMAX] = fmax(acc[C_DMAX], fabs(d));
            if (fl) acc[C_DVIOL] = fmax(acc[C_DVIOL], -d);
            if (fu) acc[C_DVIOL] = fmax(acc[C_DVIOL], d);
            S.G[e] = d;
        }
    }
    cta_reduce_store<NT, C_N, C_NSUM>(acc, S.partC, S.ld, blockIdx.y * NT);
}

// Row pass of an evaluation: A x' and A d, primal residual, b.y', iterate movement.
template <int NT>
__global__ void __launch_bounds__(kCtaThreads)
k_eval_rows(const DevProb P, const DevState S, const int rows_per_cta) {
    constexpr int RW = 32 / NT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int node = blockIdx.y * NT + (lane % NT);
    const int sub = lane / NT;
    const bool node_ok = node < S.B && S.fin[node] == 0;
    if (__ballot_sync(0xffffffffu, node_ok) == 0) return;     // uniform over the CTA
    double acc[R_N];
#pragma unroll
    for (int a = 0; a < R_N; ++a) acc[a] = 0.0;
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = min(P.m, r0 + rows_per_cta);
    for (int ib = r0 + warp * RW; ib < r1; ib += kWarps * RW) {
        const int i = ib + sub;
        const bool row_ok = i < r1;
        const size_t e = (size_t)i * S.ld + node;
        const bool ok = row_ok && node_ok;
        const double ax = row_dot<NT>(P.rowptr, P.colidx, P.val, row_ok ? i : 0, row_ok, S.X1, S.ld,
                                      node, ok, lane);
        const double ad = row_dot<NT>(P.rowptr, P.colidx, P.val, row_ok ? i : 0, row_ok, S.G, S.ld,
                                      node, ok, lane);
        if (ok) {
            bool on = true;
            if (i >= P.m_base && S.rowmask) on = S.rowmask[(size_t)(i - P.m_base) * S.ld + node] != 0;
            if (on) {
                const double yp = S.Y1[e], dy = S.DY[e], ya = S.ya[e];
                const double bi = __ldg(P.b + i);
                const double pr = fmax(bi - ax, 0.0) * __ldg(P.rowscale + i);
                acc[R_PRES2] = fma(pr, pr, acc[R_PRES2]);
                acc[R_BY] = fma(bi, yp, acc[R_BY]);
                acc[R_BYABS] += fabs(bi * yp);
                acc[R_DY2] = fma(dy, dy, acc[R_DY2]);
                const double t = yp - ya;
                acc[R_DYA2] = fma(t, t, acc[R_DYA2]);
                acc[R_ADNEG] = fmax(acc[R_ADNEG], -ad);
                acc[R_YMAX] = fmax(acc[R_YMAX], yp);
            }
        }
    }
    cta_reduce_store<NT, R_N, R_NSUM>(acc, S.partR, S.ld, blockIdx.y * NT);
}

// ---------------------------------------------------------------------------------------------
// One thread per node: fold the chunk partials in fixed order, test termination and the
// infeasibility / unboundedness certificates, decide restarts, update the primal weight.
struct DecideArgs {
    int chunksC, chunksR, steps_in_period, max_iters;
    double eps, eps_inf;
};

// counters: [0] nodes still running after the evaluation, [1] nodes restarting,
//           [2] PDHG iterations executed so far (advanced by k_tick at the start of a period,
//               so one captured period graph can be replayed unchanged)
__global__ void k_tick(const DevState S, const int steps) {
    S.counters[0] = 0;
    S.counters[1] = 0;
    S.counters[2] += steps;
}

__global__ void k_decide(const DevProb P, const DevState S, const DecideArgs D) {
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= S.B) return;
    S.restart[node] = 0;
    if (S.fin[node] != 0) return;
    double c[C_N], r[R_N];
#pragma unroll
    for (int a = 0; a < C_N; ++a) c[a] = 0.0;
#pragma unroll
    for (int a = 0; a < R_N; ++a) r[a] = 0.0;
    for (int ch = 0; ch < D.chunksC; ++ch)
#pragma unroll
        for (int a = 0; a < C_N; ++a) {
            const double v = S.partC[((size_t)ch * C_N + a) * S.ld + node];
            c[a] = (a < C_NSUM) ? c[a] + v : fmax(c[a], v);
        }
    for (int ch = 0; ch < D.chunksR; ++ch)
#pragma unroll
        for (int a = 0; a < R_N; ++a) {
            const double v = S.partR[((size_t)ch * R_N + a) * S.ld + node];
            r[a] = (a < R_NSUM) ? r[a] + v : fmax(r[a], v);
        }
    const double omega = S.omega[node];
    const double tau = P.eta / omega, sig = P.eta * omega;
    const double fpe = sqrt(fmax(c[C_DX2] / tau + r[R_DY2] / sig + 2.0 * c[C_CROSS], 0.0));
    const double pobj = c[C_CX] * P.objscale;
    const double dobj = (r[R_BY] + c[C_BND]) * P.objscale;
    const double rp = sqrt(r[R_PRES2]) / (1.0 + P.bnorm0);
    const double rd = sqrt(c[C_DRES2]) / (1.0 + P.cnorm0);
    const double rg = fabs(pobj - dobj) / (1.0 + fabs(pobj) + fabs(dobj));
    S.pobj[node] = pobj;
    S.dobj[node] = dobj;
    const int total = S.counters[2];
    const bool last = total >= D.max_iters;
    int st = -1;
    if (rp <= D.eps && rd <= D.eps && rg <= D.eps) {
        st = 0;
    } else {
        // Farkas certificate of primal infeasibility from the dual iterate y' >= 0:
        //   b.y' > max_{l<=x<=u} (A'y').x     (columns with an infinite bound must not need it)
        const double farkas = r[R_BY] - c[C_BOX];
        if (farkas > 0.0 && farkas > D.eps_inf * (r[R_BYABS] + c[C_BOXABS]) &&
            c[C_RAYVIOL] <= 1e-8 * r[R_YMAX])
            st = 1;
        // primal ray d = x' - xa: c.d < 0, A d >= 0, d respects finite bounds => unbounded
        const double dmax = c[C_DMAX];
        if (st < 0 && dmax > 0.0 && c[C_CD] < -1e-6 * dmax * P.cinf_s &&
            r[R_ADNEG] <= 1e-8 * dmax && c[C_DVIOL] <= 1e-8 * dmax)
            st = 2;
    }
    if (st < 0 && last) st = 3;
    if (st >= 0) {
        S.status[node] = st;
        S.fin[node] = 1;
        S.iters[node] = total;
        if (st == 1) S.pobj[node] = INFINITY;
        return;
    }
    atomicAdd(S.counters + 0, 1);
    // restart test on the fixed-point error (sufficient / necessary / artificial)
    const int s_now = S.sbase[node] + D.steps_in_period;
    const double f0 = S.fpe0[node], fprev = S.fpe_prev[node];
    const bool first = !(f0 < INFINITY);
    const bool do_restart = first || fpe <= 0.2 * f0 || (fpe <= 0.8 * f0 && fpe > fprev) ||
                            (double)s_now >= 0.36 * (double)total;
    if (do_restart) {
        const double ddx = sqrt(c[C_DXA2]), ddy = sqrt(r[R_DYA2]);
        if (!first && ddx > 1e-10 && ddy > 1e-10)
            S.omega[node] = exp(0.5 * log(ddy / ddx) + 0.5 * log(omega));
        S.fpe0[node] = fpe;
        S.fpe_prev[node] = INFINITY;
        S.sbase[node] = 0;
        S.restart[node] = 1;
        atomicAdd(S.counters + 1, 1);
    } else {
        S.fpe_prev[node] = fpe;
        S.sbase[node] = s_now;
    }
}

// Restart: the anchor and the current point both become T(z) of the evaluated iterate.
// grid = (row chunks, ceil(ld/32)); a warp owns 32 node columns and strides over the rows.
__global__ void __launch_bounds__(kCtaThreads)
k_apply_restart(const DevProb P, const DevState S) {
    if (S.counters[1] == 0) return;
    const int lane = threadIdx.x & 31;
    const int node = blockIdx.y * 32 + lane;
    const bool mine = node < S.B && S.restart[node] != 0;
    if (__ballot_sync(0xffffffffu, mine) == 0) return;
    if (!mine) return;
    const int gwarp = (blockIdx.x * kCtaThreads + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * kCtaThreads) >> 5;
    for (int j = gwarp; j < P.n; j += nwarps) {
        const size_t e = (size_t)j * S.ld + node;
        const double v = S.X1[e];
        S.xa[e] = v;
        S.xbar[e] = v;
    }
    for (int i = gwarp; i < P.m; i += nwarps) {
        const size_t e = (size_t)i * S.ld + node;
        const double v = S.Y1[e];
        S.ya[e] = v;
        S.y[e] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Set-up: scale the per-node bounds into the solver's space and build the start point.
__global__ void __launch_bounds__(kCtaThreads)
k_init_cols(const DevProb P, const DevState S, const double* __restrict__ lb,
            const double* __restrict__ ub, const double* __restrict__ x0) {
    const size_t total = (size_t)P.n * S.ld;
    for (size_t e = (size_t)blockIdx.x * kCtaThreads + threadIdx.x; e < total;
         e += (size_t)gridDim.x * kCtaThreads) {
        const int j = (int)(e / S.ld), node = (int)(e % S.ld);
        double lo = 0.0, hi = 0.0, x = 0.0;
        if (node < S.B) {
            const double f = P.sb / P.dc[j];
            const double a = lb[e], b = ub[e];
            lo = is_inf(a) ? (a > 0 ? INFINITY : -INFINITY) : a * f;
            hi = is_inf(b) ? (b > 0 ? INFINITY : -INFINITY) : b * f;
            x = x0 ? x0[e] * f : 0.0;
            x = fmin(fmax(x, lo), hi);
        }
        if (lo > hi) {              // empty box: primal infeasible without any iteration
            S.fin[node] = 1; S.status[node] = 1; S.pobj[node] = INFINITY; S.dobj[node] = INFINITY;
        }
        S.l[e] = lo; S.u[e] = hi; S.xa[e] = x; S.xbar[e] = x; S.X1[e] = x;
        S.DX[e] = 0.0; S.G[e] = 0.0;
    }
}

__global__ void __launch_bounds__(kCtaThreads)
k_init_rows(const DevProb P, const DevState S, const double* __restrict__ y0) {
    const size_t total = (size_t)P.m * S.ld;
    for (size_t e = (size_t)blockIdx.x * kCtaThreads + threadIdx.x; e < total;
         e += (size_t)gridDim.x * kCtaThreads) {
        const int i = (int)(e / S.ld), node = (int)(e % S.ld);
        double y = 0.0;
        if (node < S.B && y0) {
            y = fmax(0.0, y0[e] * P.sc / P.dr[i]);
            if (i >= P.m_base && S.rowmask && S.rowmask[(size_t)(i - P.m_base) * S.ld + node] == 0)
                y = 0.0;
        }
        S.y[e] = y; S.ya[e] = y; S.Y1[e] = y; S.DY[e] = 0.0;
    }
}

// Row-activity bound test (the cheap infeasibility screen for branching children, e.g. the
// right child x2 >= 2 of small_branch against x0 + x2 <= 1.5): a row whose largest possible
// activity over the node's box stays below its lower bound proves the node LP infeasible.
template <int NT>
__global__ void __launch_bounds__(kCtaThreads)
k_check_rows(const DevProb P, const DevState S, const int rows_per_cta) {
    constexpr int RW = 32 / NT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int node = blockIdx.y * NT + (lane % NT);
    const int sub = lane / NT;
    const bool node_ok = node < S.B;
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = min(P.m, r0 + rows_per_cta);
    for (int ib = r0 + warp * RW; ib < r1; ib += kWarps * RW) {
        const int i = ib + sub;
        if (i >= r1 || !node_ok) continue;
        if (i >= P.m_base && S.rowmask && S.rowmask[(size_t)(i - P.m_base) * S.ld + node] == 0)
            continue;
        double act = 0.0, mag = 0.0;
        for (int p = __ldg(P.rowptr + i); p < __ldg(P.rowptr + i + 1); ++p) {
            const double a = __ldg(P.val + p);
            const size_t e = (size_t)__ldg(P.colidx + p) * S.ld + node;
            const double t = a * (a > 0.0 ? S.u[e] : S.l[e]);
            act += t;
            mag += fabs(t);
        }
        const double bi = __ldg(P.b + i);
        if (bi - act > 1e-9 * (fabs(bi) + mag) + 1e-300) {
            S.fin[node] = 1; S.status[node] = 1; S.pobj[node] = INFINITY; S.dobj[node] = INFINITY;
        }
    }
}

__global__ void k_count_active(const DevState S) {
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node < S.B && S.fin[node] == 0) atomicAdd(S.counters + 0, 1);
}

__global__ void k_init_nodes(const DevProb P, const DevState S) {
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= S.ld) return;
    S.omega[node] = P.omega0;
    S.fpe0[node] = INFINITY;
    S.fpe_prev[node] = INFINITY;
    S.pobj[node] = 0.0;
    S.dobj[node] = -INFINITY;
    S.sbase[node] = 0;
    S.fin[node] = node < S.B ? 0 : 2;
    S.status[node] = 3;
    S.iters[node] = 0;
    S.restart[node] = 0;
}

// Epilogue: unscale x', y' into the caller's arrays, per-node scalars, most fractional column.
__global__ void __launch_bounds__(kCtaThreads)
k_out_vec(const double* __restrict__ src, const double* __restrict__ diag, const double inv,
          const int rows, const int ld, const int B, double* __restrict__ dst) {
    const size_t total = (size_t)rows * ld;
    for (size_t e = (size_t)blockIdx.x * kCtaThreads + threadIdx.x; e < total;
         e += (size_t)gridDim.x * kCtaThreads) {
        const int r = (int)(e / ld), node = (int)(e % ld);
        dst[e] = node < B ? src[e] * diag[r] * inv : 0.0;
    }
}

__global__ void k_out_nodes(const DevProb P, const DevState S, const int32_t* __restrict__ int_idx,
                            const int n_int, const double frac_eps, double* __restrict__ obj,
                            double* __restrict__ lower, int32_t* __restrict__ status,
                            int32_t* __restrict__ iters, int32_t* __restrict__ frac_idx) {
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= S.ld) return;
    const bool real = node < S.B;
    const int st = real ? S.status[node] : -1;
    if (obj) obj[node] = real ? S.pobj[node] : 0.0;
    if (lower) lower[node] = real ? S.dobj[node] : 0.0;
    if (status) status[node] = st;
    if (iters) iters[node] = real ? S.iters[node] : 0;
    if (frac_idx) {
        int best = -1;
        if (st == 0 && int_idx) {
            double far = frac_eps;
            const double inv = 1.0 / P.sb;
            for (int q = 0; q < n_int; ++q) {
                const int j = int_idx[q];
                const double v = S.X1[(size_t)j * S.ld + node] * P.dc[j] * inv;
                const double dist = fmin(v - floor(v), ceil(v) - v);
                if (dist > far) { far = dist; best = j; }
            }
        }
        frac_idx[node] = best;
    }
}

// ---------------------------------------------------------------------------------------------
// Plain batched SpMV on the unscaled matrix (parity tests, SpMV roofline measurement).
template <int NT>
__global__ void __launch_bounds__(kCtaThreads)
k_spmv(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
       const double* __restrict__ val, const int rows, const int B, const int ld,
       const double* __restrict__ X, double* __restrict__ Y, const int rows_per_cta) {
    constexpr int RW = 32 / NT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int node = blockIdx.y * NT + (lane % NT);
    const int sub = lane / NT;
    const bool node_ok = node < B;
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = min(rows, r0 + rows_per_cta);
    for (int ib = r0 + warp * RW; ib < r1; ib += kWarps * RW) {
        const int i = ib + sub;
        const bool row_ok = i < r1;
        const double s = row_dot<NT>(ptr, idx, val, row_ok ? i : 0, row_ok, X, ld, node,
                                     node_ok && row_ok, lane);
        if (row_ok && node_ok) Y[(size_t)i * ld + node] = s;
    }
}

// node-major host order [B][rows]  <->  node-fastest device order [rows][ld] (tiled transpose)
__global__ void k_transpose_in(const double* __restrict__ src, const int B, const int rows,
                               const int ld, double* __restrict__ dst) {
    __shared__ double tile[32][33];
    const int r0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        const int k = k0 + q, r = r0 + threadIdx.x;
        tile[q][threadIdx.x] = (k < B && r < rows) ? src[(size_t)k * rows + r] : 0.0;
    }
    __syncthreads();
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        const int r = r0 + q, k = k0 + threadIdx.x;
        if (r < rows && k < ld) dst[(size_t)r * ld + k] = tile[threadIdx.x][q];
    }
}

__global__ void k_transpose_out(const double* __restrict__ src, const int B, const int rows,
                                const int ld, double* __restrict__ dst) {
    __shared__ double tile[32][33];
    const int r0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        const int r = r0 + q, k = k0 + threadIdx.x;
        tile[q][threadIdx.x] = (r < rows && k < ld) ? src[(size_t)r * ld + k] : 0.0;
    }
    __syncthreads();
    for (int q = threadIdx.y; q < 32; q += blockDim.y) {
        const int k = k0 + q, r = r0 + threadIdx.x;
        if (k < B && r < rows) dst[(size_t)k * rows + r] = tile[threadIdx.x][q];
    }
}

__global__ void k_transpose_in_u8(const uint8_t* __restrict__ src, const int B, const int rows,
                                  const int ld, uint8_t* __restrict__ dst) {
    const size_t total = (size_t)rows * ld;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / ld), k = (int)(e % ld);
        dst[e] = k < B ? src[(size_t)k * rows + r] : 0;
    }
}

// children-of-one-parent bounds: broadcast the parent's vectors, then patch the deltas
__global__ void k_broadcast_rows(const double* __restrict__ v, const int rows, const int ld,
                                 double* __restrict__ dst) {
    const size_t total = (size_t)rows * ld;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x)
        dst[e] = v[e / ld];
}

__global__ void k_patch_bounds(const int B, const int ld, const int32_t* __restrict__ dptr,
                               const int32_t* __restrict__ dvar, const double* __restrict__ dlb,
                               const double* __restrict__ dub, double* __restrict__ lb,
                               double* __restrict__ ub) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= B) return;
    for (int p = dptr[k]; p < dptr[k + 1]; ++p) {
        lb[(size_t)dvar[p] * ld + k] = dlb[p];
        ub[(size_t)dvar[p] * ld + k] = dub[p];
    }
}

}  // namespace blp
