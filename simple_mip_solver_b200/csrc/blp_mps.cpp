// MPS reader of libblp.so (include/blp.h: blp_mps_*), host only.
//
// The reference loads its test models with MILPInstance(file_name=...)
// (test_simple_mip_solver/helpers.py:42, example_models.py:43-48), i.e. through CLP's MPS reader;
// SURVEY section 8(f) #3 asks for a native reader for instances of the C4 / C5 size, where a
// line-by-line Python reader takes minutes. Dialect: the one CLP writes (free, whitespace separated
// fields; sections NAME / ROWS / COLUMNS / RHS / BOUNDS / ENDATA; integer columns flagged by UI / LI /
// BV bounds or by MARKER lines; an optional set name in RHS and BOUNDS lines). RANGES are not
// produced by the reference's writer and are ignored. Empty rows and columns are kept.
//
// The result is raw model data as written in the file: rows with their sense (L / G / E), not yet
// the solver's canonical ">=" form — MILPInstance does that conversion as the reference does
// (algorithms/base_algorithm.py:53-59).
#include "../../include/blp.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

struct blp_mps_s {
    std::string name;
    std::vector<std::string> row_names, col_names;
    std::vector<char> sense;
    std::vector<int32_t> ptr, idx;
    std::vector<double> val, rhs, c, l, u;
    std::vector<int32_t> ints;
    double obj_offset = 0.0;
};

namespace {

thread_local std::string g_mps_err;

struct Triple {
    int32_t r, c;
    double v;
};

std::string upper(std::string s) {
    for (char& ch : s) ch = (char)std::toupper((unsigned char)ch);
    return s;
}

// whitespace-separated tokens of a line
void split(const char* line, std::vector<std::string>& tok) {
    tok.clear();
    const char* p = line;
    while (*p) {
        while (*p && std::isspace((unsigned char)*p)) ++p;
        if (!*p) break;
        const char* q = p;
        while (*q && !std::isspace((unsigned char)*q)) ++q;
        tok.emplace_back(p, q - p);
        p = q;
    }
}

bool to_double(const std::string& s, double* out) {
    char* end = nullptr;
    const double v = std::strtod(s.c_str(), &end);
    if (end == s.c_str() || *end != '\0') return false;
    *out = v;
    return true;
}

}  // namespace

extern "C" {

const char* blp_mps_last_error(void) { return g_mps_err.c_str(); }

int blp_mps_read(const char* path, blp_mps* out) {
    if (!out) { g_mps_err = "blp_mps_read: out is NULL"; return BLP_ERR_ARG; }
    *out = nullptr;
    if (!path) { g_mps_err = "blp_mps_read: path is NULL"; return BLP_ERR_ARG; }
    FILE* fh = std::fopen(path, "r");
    if (!fh) { g_mps_err = std::string("blp_mps_read: cannot open ") + path; return BLP_ERR_ARG; }

    blp_mps_s* M = new (std::nothrow) blp_mps_s;
    if (!M) { std::fclose(fh); g_mps_err = "blp_mps_read: out of memory"; return BLP_ERR_NOMEM; }
    std::unordered_map<std::string, int32_t> row_of, col_of;
    std::vector<Triple> ent;
    std::vector<char> has_l, has_u, is_int;
    std::string obj_row;
    bool have_obj = false, in_marker = false;
    enum { NONE, ROWS, COLUMNS, RHS, BOUNDS, OTHER } sec = NONE;
    const double inf = HUGE_VAL;

    auto column = [&](const std::string& nm) -> int32_t {
        auto it = col_of.find(nm);
        if (it != col_of.end()) return it->second;
        const int32_t j = (int32_t)M->col_names.size();
        col_of.emplace(nm, j);
        M->col_names.push_back(nm);
        M->c.push_back(0.0);
        M->l.push_back(0.0);
        M->u.push_back(inf);
        has_l.push_back(0);
        has_u.push_back(0);
        is_int.push_back(in_marker ? 1 : 0);
        return j;
    };
    auto bad = [&](long lineno, const char* what) {
        char buf[256];
        std::snprintf(buf, sizeof buf, "blp_mps_read: %s line %ld: %s", path, lineno, what);
        g_mps_err = buf;
        std::fclose(fh);
        delete M;
        return (int)BLP_ERR_ARG;
    };

    std::vector<std::string> tok;
    std::string line;
    char buf[4096];
    long lineno = 0;
    bool done = false;
    while (!done && std::fgets(buf, sizeof buf, fh)) {
        line = buf;
        while (!line.empty() && line.back() != '\n' && std::fgets(buf, sizeof buf, fh)) line += buf;   // long lines
        ++lineno;
        split(line.c_str(), tok);
        if (tok.empty() || tok[0][0] == '*') continue;
        if (!std::isspace((unsigned char)line[0])) {                  // section header
            const std::string s = upper(tok[0]);
            if (s == "NAME") { if (tok.size() > 1) M->name = tok[1]; sec = NONE; }
            else if (s == "ROWS") sec = ROWS;
            else if (s == "COLUMNS") sec = COLUMNS;
            else if (s == "RHS") sec = RHS;
            else if (s == "BOUNDS") sec = BOUNDS;
            else if (s == "ENDATA") done = true;
            else sec = OTHER;                                         // RANGES, OBJSENSE, ...: ignored
            continue;
        }
        double v = 0.0;
        switch (sec) {
            case ROWS: {
                if (tok.size() < 2) return bad(lineno, "ROWS line needs a sense and a name");
                const char s = (char)std::toupper((unsigned char)tok[0][0]);
                if (s == 'N') {
                    if (!have_obj) { obj_row = tok[1]; have_obj = true; }
                } else {
                    row_of[tok[1]] = (int32_t)M->row_names.size();
                    M->row_names.push_back(tok[1]);
                    M->sense.push_back(s);
                }
                break;
            }
            case COLUMNS: {
                if (tok.size() >= 3 && upper(tok[1]) == "'MARKER'") {
                    in_marker = upper(tok[2]) == "'INTORG'";
                    break;
                }
                const int32_t j = column(tok[0]);
                for (size_t k = 1; k + 1 < tok.size(); k += 2) {
                    if (!to_double(tok[k + 1], &v)) return bad(lineno, "bad number in COLUMNS");
                    if (have_obj && tok[k] == obj_row) M->c[j] = v;
                    else {
                        auto it = row_of.find(tok[k]);
                        if (it != row_of.end()) ent.push_back(Triple{it->second, j, v});
                    }
                }
                break;
            }
            case RHS: {
                for (size_t k = (tok.size() % 2 == 1) ? 1 : 0; k + 1 < tok.size(); k += 2) {   // optional set name
                    if (!to_double(tok[k + 1], &v)) return bad(lineno, "bad number in RHS");
                    if (have_obj && tok[k] == obj_row) M->obj_offset = -v;
                    else {
                        auto it = row_of.find(tok[k]);
                        if (it != row_of.end()) {
                            if (M->rhs.size() < M->row_names.size()) M->rhs.resize(M->row_names.size(), 0.0);
                            M->rhs[it->second] = v;
                        }
                    }
                }
                break;
            }
            case BOUNDS: {
                const std::string kind = upper(tok[0]);
                const bool no_value = kind == "FR" || kind == "MI" || kind == "PL" || kind == "BV";
                if (tok.size() < 2) return bad(lineno, "BOUNDS line too short");
                std::string cname;
                if (no_value) cname = tok.size() >= 3 ? tok[2] : tok[1];
                else {
                    cname = tok.size() >= 4 ? tok[2] : tok[1];
                    if (!to_double(tok.back(), &v)) return bad(lineno, "bad number in BOUNDS");
                }
                const int32_t j = column(cname);
                if (kind == "UP") {
                    M->u[j] = v; has_u[j] = 1;
                    if (v < 0 && !has_l[j]) { M->l[j] = -inf; has_l[j] = 1; }
                } else if (kind == "UI") { M->u[j] = v; has_u[j] = 1; is_int[j] = 1; }
                else if (kind == "LO") { M->l[j] = v; has_l[j] = 1; }
                else if (kind == "LI") { M->l[j] = v; has_l[j] = 1; is_int[j] = 1; }
                else if (kind == "FX") { M->l[j] = M->u[j] = v; has_l[j] = has_u[j] = 1; }
                else if (kind == "FR") { M->l[j] = -inf; M->u[j] = inf; has_l[j] = has_u[j] = 1; }
                else if (kind == "MI") { M->l[j] = -inf; has_l[j] = 1; }
                else if (kind == "PL") { M->u[j] = inf; has_u[j] = 1; }
                else if (kind == "BV") { M->l[j] = 0.0; M->u[j] = 1.0; has_l[j] = has_u[j] = 1; is_int[j] = 1; }
                break;
            }
            default: break;
        }
    }
    std::fclose(fh);

    const int32_t m = (int32_t)M->row_names.size(), n = (int32_t)M->col_names.size();
    M->rhs.resize(m, 0.0);
    // CSR with sorted column indices; duplicate entries are summed
    std::stable_sort(ent.begin(), ent.end(), [](const Triple& a, const Triple& b) {
        return a.r != b.r ? a.r < b.r : a.c < b.c;
    });
    M->ptr.assign(m + 1, 0);
    for (size_t k = 0; k < ent.size(); ++k) {
        if (k && ent[k].r == ent[k - 1].r && ent[k].c == ent[k - 1].c) {
            M->val.back() += ent[k].v;
            continue;
        }
        M->idx.push_back(ent[k].c);
        M->val.push_back(ent[k].v);
        M->ptr[ent[k].r + 1] += 1;
    }
    for (int32_t i = 0; i < m; ++i) M->ptr[i + 1] += M->ptr[i];
    for (int32_t j = 0; j < n; ++j)
        if (is_int[j]) M->ints.push_back(j);
    *out = M;
    return BLP_OK;
}

int blp_mps_dims(blp_mps M, int32_t* m, int32_t* n, int64_t* nnz, int32_t* n_int) {
    if (!M) { g_mps_err = "blp_mps_dims: NULL model"; return BLP_ERR_ARG; }
    if (m) *m = (int32_t)M->row_names.size();
    if (n) *n = (int32_t)M->col_names.size();
    if (nnz) *nnz = (int64_t)M->idx.size();
    if (n_int) *n_int = (int32_t)M->ints.size();
    return BLP_OK;
}

int blp_mps_copy(blp_mps M, int32_t* rowptr, int32_t* colidx, double* val, double* rhs, char* sense,
                 double* c, double* obj_offset, double* l, double* u, int32_t* int_idx) {
    if (!M) { g_mps_err = "blp_mps_copy: NULL model"; return BLP_ERR_ARG; }
    auto cp = [](auto* dst, const auto& src) {
        if (dst && !src.empty()) std::memcpy(dst, src.data(), src.size() * sizeof(src[0]));
    };
    cp(rowptr, M->ptr);
    cp(colidx, M->idx);
    cp(val, M->val);
    cp(rhs, M->rhs);
    cp(sense, M->sense);
    cp(c, M->c);
    cp(l, M->l);
    cp(u, M->u);
    cp(int_idx, M->ints);
    if (obj_offset) *obj_offset = M->obj_offset;
    return BLP_OK;
}

const char* blp_mps_name(blp_mps M) { return M ? M->name.c_str() : ""; }

const char* blp_mps_row_name(blp_mps M, int32_t i) {
    return (M && i >= 0 && i < (int32_t)M->row_names.size()) ? M->row_names[i].c_str() : "";
}

const char* blp_mps_col_name(blp_mps M, int32_t j) {
    return (M && j >= 0 && j < (int32_t)M->col_names.size()) ? M->col_names[j].c_str() : "";
}

int blp_mps_free(blp_mps M) {
    delete M;
    return BLP_OK;
}

}  // extern "C"
