// Host-side set-up of the shared LP data: diagonal scaling, transpose, step size.
// Runs once per instance (and once per cut-row append); not on the per-iteration path.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace blp {

struct HostCsr {
    int rows = 0, cols = 0;
    std::vector<int32_t> ptr, idx;
    std::vector<double> val;
    int64_t nnz() const { return (int64_t)idx.size(); }
};

inline HostCsr transpose(const HostCsr& a) {
    HostCsr t;
    t.rows = a.cols;
    t.cols = a.rows;
    t.ptr.assign(a.cols + 1, 0);
    t.idx.resize(a.idx.size());
    t.val.resize(a.val.size());
    for (int32_t j : a.idx) t.ptr[j + 1]++;
    for (int j = 0; j < a.cols; ++j) t.ptr[j + 1] += t.ptr[j];
    std::vector<int32_t> fill(t.ptr.begin(), t.ptr.end() - 1);
    for (int i = 0; i < a.rows; ++i)
        for (int32_t p = a.ptr[i]; p < a.ptr[i + 1]; ++p) {
            int32_t q = fill[a.idx[p]]++;
            t.idx[q] = i;
            t.val[q] = a.val[p];
        }
    return t;
}

// Ruiz equilibration (iters passes) followed by one Pock-Chambolle alpha=1 pass.
// On return a.val holds the scaled matrix, dr/dc the accumulated scalings.
// Rows [row_begin, rows) are scaled; with row_begin > 0 the column scaling dc is kept fixed
// (cut rows appended to an already scaled problem).
inline void ruiz_pc(HostCsr& a, std::vector<double>& dr, std::vector<double>& dc, int iters,
                    int row_begin) {
    const int m = a.rows, n = a.cols;
    const bool cols_free = (row_begin == 0);
    std::vector<double> rs(m), cs(n);
    auto apply = [&]() {
        for (int i = row_begin; i < m; ++i)
            for (int32_t p = a.ptr[i]; p < a.ptr[i + 1]; ++p) a.val[p] *= rs[i] * cs[a.idx[p]];
        for (int i = row_begin; i < m; ++i) dr[i] *= rs[i];
        if (cols_free)
            for (int j = 0; j < n; ++j) dc[j] *= cs[j];
    };
    for (int it = 0; it < iters + 1; ++it) {
        const bool pc = (it == iters);
        std::fill(rs.begin(), rs.end(), 0.0);
        std::fill(cs.begin(), cs.end(), 0.0);
        for (int i = row_begin; i < m; ++i)
            for (int32_t p = a.ptr[i]; p < a.ptr[i + 1]; ++p) {
                double v = std::fabs(a.val[p]);
                if (pc) {
                    rs[i] += v;
                    cs[a.idx[p]] += v;
                } else {
                    rs[i] = std::max(rs[i], v);
                    cs[a.idx[p]] = std::max(cs[a.idx[p]], v);
                }
            }
        for (int i = 0; i < m; ++i) rs[i] = (i >= row_begin && rs[i] > 0) ? 1.0 / std::sqrt(rs[i]) : 1.0;
        for (int j = 0; j < n; ++j) cs[j] = (cols_free && cs[j] > 0) ? 1.0 / std::sqrt(cs[j]) : 1.0;
        apply();
    }
}

inline void spmv(const HostCsr& a, const double* x, double* y) {
    for (int i = 0; i < a.rows; ++i) {
        double s = 0;
        for (int32_t p = a.ptr[i]; p < a.ptr[i + 1]; ++p) s += a.val[p] * x[a.idx[p]];
        y[i] = s;
    }
}

// Largest singular value by power iteration on A'A with a fixed deterministic start vector.
inline double sigma_max(const HostCsr& a, const HostCsr& at, int iters = 80) {
    if (a.nnz() == 0) return 1.0;
    std::vector<double> v(a.cols), w(a.rows);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (int j = 0; j < a.cols; ++j) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        v[j] = 0.5 + (double)(s >> 11) / 9007199254740992.0;   // in [0.5, 1.5): not orthogonal to
    }                                                           // the dominant (Perron-like) vector
    double sig = 1.0, nv = 0;
    for (double t : v) nv += t * t;
    nv = std::sqrt(nv);
    for (double& t : v) t /= nv;
    for (int it = 0; it < iters; ++it) {
        spmv(a, v.data(), w.data());
        spmv(at, w.data(), v.data());
        nv = 0;
        for (double t : v) nv += t * t;
        nv = std::sqrt(nv);
        if (nv == 0) return 1.0;
        double ns = std::sqrt(nv);
        for (double& t : v) t /= nv;
        if (std::fabs(ns - sig) <= 1e-10 * ns) { sig = ns; break; }
        sig = ns;
    }
    return sig;
}

inline double norm2(const std::vector<double>& v) {
    double s = 0;
    for (double t : v) s += t * t;
    return std::sqrt(s);
}

}  // namespace blp
