// ldpc_code.cpp -- H-file loaders and code constructors (host only, no CUDA).
//
// Format A is what FP_Decoder::ReadH parses (ArrayLDPC_Decoder.cpp:650-671): "n m / dv dc /
// vdeg[n] / cdeg[m] / n variable rows / m check rows", zero based, rows not padded.
// Format C is the legacy one-based check-list layout of H2212_316_array_cut79.txt and
// H_array_2209_235_old.txt ("n m / cdeg[m] / m check rows"), for which the reference has no
// reader.  Unlike ReadH every read is checked and the two adjacency lists are cross-validated.
#include "ldpc_code.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "../../include/ldpc_capi.h"

namespace ldpc {

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }
const char *last_error() { return g_last_error.c_str(); }

// Derive vlist / vslot / bounds from sorted check rows.
static int finish_tables(ldpc_code &c)
{
    c.edges = 0;
    c.vdeg.assign(c.n, 0);
    for (int r = 0; r < c.m; ++r) {
        int d = c.cdeg[r];
        if (d < 2) {  // the recursion reads Backward[1] / Forward[d-2] (ArrayLDPC_Decoder.cpp:92-105)
            set_error("check " + std::to_string(r) + " has degree < 2");
            return LDPC_ERR_FORMAT;
        }
        int *row = &c.clist[(size_t)r * c.dc_max];
        std::sort(row, row + d);
        for (int k = 0; k < d; ++k) {
            if (row[k] < 0 || row[k] >= c.n || (k && row[k] == row[k - 1])) {
                set_error("check " + std::to_string(r) + ": variable index out of range or repeated");
                return LDPC_ERR_FORMAT;
            }
            c.vdeg[row[k]]++;
        }
        c.edges += d;
    }
    c.dv_max = 0;
    for (int v = 0; v < c.n; ++v) c.dv_max = std::max(c.dv_max, c.vdeg[v]);
    if (c.dv_max == 0) { set_error("code has no edges"); return LDPC_ERR_FORMAT; }
    c.vlist.assign((size_t)c.n * c.dv_max, -1);
    c.vslot.assign((size_t)c.n * c.dv_max, -1);
    std::vector<int> fill(c.n, 0);
    for (int r = 0; r < c.m; ++r)  // ascending r => ascending vlist rows
        for (int k = 0; k < c.cdeg[r]; ++k) {
            int v = c.clist[(size_t)r * c.dc_max + k];
            c.vlist[(size_t)v * c.dv_max + fill[v]] = r;
            c.vslot[(size_t)v * c.dv_max + fill[v]] = k;
            fill[v]++;
        }
    if (c.rate == 0.0) c.rate = double(c.n - c.m) / c.n;
    return LDPC_OK;
}

int build_from_checks(int n, int m, const int *cdeg, const int *clist, int cstride, ldpc_code &out)
{
    if (n <= 0 || m <= 0 || !cdeg || !clist || cstride <= 0) { set_error("bad argument"); return LDPC_ERR_ARG; }
    ldpc_code c;
    c.n = n; c.m = m;
    c.cdeg.assign(cdeg, cdeg + m);
    c.dc_max = 0;
    for (int r = 0; r < m; ++r) {
        if (cdeg[r] < 0 || cdeg[r] > cstride) { set_error("check degree exceeds row stride"); return LDPC_ERR_ARG; }
        c.dc_max = std::max(c.dc_max, cdeg[r]);
    }
    c.clist.assign((size_t)m * c.dc_max, -1);
    for (int r = 0; r < m; ++r)
        for (int k = 0; k < cdeg[r]; ++k) c.clist[(size_t)r * c.dc_max + k] = clist[(size_t)r * cstride + k];
    int st = finish_tables(c);
    if (st == LDPC_OK) out = std::move(c);
    return st;
}

static bool read_tokens(const char *path, std::vector<long> &tok)
{
    std::ifstream in(path);
    if (!in) return false;
    std::stringstream ss;
    ss << in.rdbuf();
    std::string s;
    while (ss >> s) {
        char *end = nullptr;
        long v = std::strtol(s.c_str(), &end, 10);
        if (*end != '\0') { set_error(std::string("non-integer token '") + s + "' in " + path); tok.clear(); return true; }
        tok.push_back(v);
    }
    return true;
}

static int parse_format_a(const std::vector<long> &t, ldpc_code &out)
{
    if (t.size() < 4) return LDPC_ERR_FORMAT;
    long n = t[0], m = t[1];
    if (n <= 0 || m <= 0 || n > (1 << 24) || m > (1 << 24)) return LDPC_ERR_FORMAT;
    size_t pos = 4;
    if (t.size() < pos + (size_t)n + (size_t)m) return LDPC_ERR_FORMAT;
    std::vector<int> vdeg(t.begin() + pos, t.begin() + pos + n); pos += n;
    std::vector<int> cdeg(t.begin() + pos, t.begin() + pos + m); pos += m;
    long ev = 0, ec = 0, dcm = 0;
    for (int d : vdeg) { if (d < 0) return LDPC_ERR_FORMAT; ev += d; }
    for (int d : cdeg) { if (d < 0) return LDPC_ERR_FORMAT; ec += d; dcm = std::max<long>(dcm, d); }
    if (ev != ec || t.size() != pos + (size_t)ev + (size_t)ec) return LDPC_ERR_FORMAT;
    std::vector<std::vector<int>> vrows(n);
    for (long v = 0; v < n; ++v)
        for (int j = 0; j < vdeg[v]; ++j) vrows[v].push_back((int)t[pos++]);
    std::vector<int> clist((size_t)m * dcm, -1);
    for (long c = 0; c < m; ++c)
        for (int k = 0; k < cdeg[c]; ++k) clist[(size_t)c * dcm + k] = (int)t[pos++];
    ldpc_code c;
    int st = build_from_checks((int)n, (int)m, cdeg.data(), clist.data(), (int)dcm, c);
    if (st != LDPC_OK) return st;
    // cross-check the file's variable rows against the ones implied by its check rows
    for (long v = 0; v < n; ++v) {
        std::sort(vrows[v].begin(), vrows[v].end());
        if ((int)vrows[v].size() != c.vdeg[v] ||
            !std::equal(vrows[v].begin(), vrows[v].end(), &c.vlist[(size_t)v * c.dv_max])) {
            set_error("variable row " + std::to_string(v) + " disagrees with the check rows");
            return LDPC_ERR_FORMAT;
        }
    }
    out = std::move(c);
    return LDPC_OK;
}

static int parse_format_c(const std::vector<long> &t, ldpc_code &out)
{
    if (t.size() < 2) return LDPC_ERR_FORMAT;
    long n = t[0], m = t[1];
    if (n <= 0 || m <= 0 || t.size() < 2 + (size_t)m) return LDPC_ERR_FORMAT;
    std::vector<int> cdeg(t.begin() + 2, t.begin() + 2 + m);
    long e = 0, dcm = 0;
    for (int d : cdeg) { if (d < 0) return LDPC_ERR_FORMAT; e += d; dcm = std::max<long>(dcm, d); }
    if (t.size() != 2 + (size_t)m + (size_t)e) return LDPC_ERR_FORMAT;
    std::vector<int> clist((size_t)m * dcm, -1);
    size_t pos = 2 + m;
    for (long c = 0; c < m; ++c)
        for (int k = 0; k < cdeg[c]; ++k) clist[(size_t)c * dcm + k] = (int)t[pos++] - 1;  // one-based
    return build_from_checks((int)n, (int)m, cdeg.data(), clist.data(), (int)dcm, out);
}

int load_file(const char *path, int format, ldpc_code &out)
{
    if (!path) { set_error("NULL path"); return LDPC_ERR_ARG; }
    std::vector<long> tok;
    set_error("");
    if (!read_tokens(path, tok)) { set_error(std::string("cannot open ") + path); return LDPC_ERR_IO; }
    if (tok.empty()) { if (!*last_error()) set_error(std::string("empty file ") + path); return LDPC_ERR_FORMAT; }
    int st;
    if (format == LDPC_FMT_A) st = parse_format_a(tok, out);
    else if (format == LDPC_FMT_C) st = parse_format_c(tok, out);
    else if (format == LDPC_FMT_AUTO) {
        st = parse_format_a(tok, out);
        if (st != LDPC_OK) st = parse_format_c(tok, out);
    } else { set_error("unknown format id"); return LDPC_ERR_ARG; }
    if (st == LDPC_ERR_FORMAT && !*last_error()) set_error(std::string("token count does not match the header in ") + path);
    return st;
}

int build_array(int p, int nrows, const int *row_mult, int ncols, const int *col_sel, int backward,
                ldpc_code &out)
{
    if (p < 2 || nrows < 1 || ncols < 2) { set_error("bad array-code shape"); return LDPC_ERR_ARG; }
    const int n = ncols * p, m = nrows * p;
    std::vector<int> cdeg(m, ncols), clist((size_t)m * ncols);
    for (int i = 0; i < nrows; ++i) {
        long a = row_mult ? row_mult[i] : i;
        for (int t = 0; t < p; ++t)
            for (int b = 0; b < ncols; ++b) {
                long sel = col_sel ? col_sel[b] : b;
                long shift = (a * sel) % p;  // ROM::CirShift, ArrayLDPCMacro.h:57
                long off = backward ? ((t - shift) % p + p) % p : (t + shift) % p;
                clist[((size_t)i * p + t) * ncols + b] = b * p + (int)off;
            }
    }
    ldpc_code c;
    int st = build_from_checks(n, m, cdeg.data(), clist.data(), ncols, c);
    if (st != LDPC_OK) return st;
    c.array_p = p; c.array_rows = nrows;
    // ROM::getRate, ArrayLDPCMacro.h:60 (meaningful for the full p x p column set)
    c.rate = 1.0 - double(nrows * p - nrows + 1) / (double(p) * p);
    out = std::move(c);
    return LDPC_OK;
}

int save_format_a(const ldpc_code &c, const char *path)
{
    FILE *f = std::fopen(path, "w");
    if (!f) { set_error(std::string("cannot write ") + path); return LDPC_ERR_IO; }
    std::fprintf(f, "%d %d\n%d %d\n", c.n, c.m, c.dv_max, c.dc_max);
    for (int v = 0; v < c.n; ++v) std::fprintf(f, "%d ", c.vdeg[v]);
    std::fprintf(f, "\n");
    for (int r = 0; r < c.m; ++r) std::fprintf(f, "%d ", c.cdeg[r]);
    std::fprintf(f, "\n");
    for (int v = 0; v < c.n; ++v) {
        for (int j = 0; j < c.vdeg[v]; ++j) std::fprintf(f, "%d ", c.vlist[(size_t)v * c.dv_max + j]);
        std::fprintf(f, "\n");
    }
    for (int r = 0; r < c.m; ++r) {
        for (int k = 0; k < c.cdeg[r]; ++k) std::fprintf(f, "%d ", c.clist[(size_t)r * c.dc_max + k]);
        std::fprintf(f, "\n");
    }
    std::fclose(f);
    return LDPC_OK;
}

// ---- generator (Format B) ------------------------------------------------------------------
// Reference: ArrayLDPC_Encoder.cpp:45-83.  Header "n rows", two ignored degree bounds, n column flags
// (1 = parity / pivot column), `rows` equation lengths, then the equations.
int load_generator(const char *path, ldpc_gen &out)
{
    if (!path) { set_error("NULL path"); return LDPC_ERR_ARG; }
    std::vector<long> t;
    set_error("");
    if (!read_tokens(path, t)) { set_error(std::string("cannot open ") + path); return LDPC_ERR_IO; }
    if (t.size() < 4) { if (!*last_error()) set_error("generator file too short"); return LDPC_ERR_FORMAT; }
    const long n = t[0], rows = t[1];
    if (n <= 0 || rows <= 0 || rows >= n || t.size() < 4 + (size_t)n + (size_t)rows) {
        set_error("bad generator header"); return LDPC_ERR_FORMAT;
    }
    ldpc_gen g;
    g.n = (int)n; g.rows = (int)rows;
    size_t pos = 4;
    g.flag.assign(t.begin() + pos, t.begin() + pos + n); pos += n;
    for (int v = 0; v < g.n; ++v) (g.flag[v] == 1 ? g.parity_index : g.info_index).push_back(v);
    if ((int)g.parity_index.size() != g.rows) { set_error("number of flagged columns != number of equations"); return LDPC_ERR_FORMAT; }
    std::vector<long> deg(t.begin() + pos, t.begin() + pos + rows); pos += rows;
    long total = 0;
    for (long d : deg) { if (d < 0) return LDPC_ERR_FORMAT; total += d; }
    if (t.size() != pos + (size_t)total) { set_error("equation lengths do not match the file size"); return LDPC_ERR_FORMAT; }
    g.eq.resize(rows);
    for (long r = 0; r < rows; ++r)
        for (long j = 0; j < deg[r]; ++j) {
            long v = t[pos++];
            if (v < 0 || v >= n) { set_error("equation entry out of range"); return LDPC_ERR_FORMAT; }
            g.eq[r].push_back((int)v);
        }
    out = std::move(g);
    return LDPC_OK;
}

// GF(2) Gauss-Jordan elimination on bit-packed rows of H.  The result has one equation per pivot:
// parity column + the information columns it depends on, i.e. exactly the shape of the reference's
// Format B rows (every row holds one flagged column).
int derive_generator(const ldpc_code &code, const int *parity_cols, int nparity, ldpc_gen &out)
{
    const int n = code.n, m = code.m, words = (n + 63) / 64;
    std::vector<std::vector<unsigned long long>> row(m, std::vector<unsigned long long>(words, 0ull));
    for (int r = 0; r < m; ++r)
        for (int k = 0; k < code.cdeg[r]; ++k) {
            int v = code.clist[(size_t)r * code.dc_max + k];
            row[r][v >> 6] ^= 1ull << (v & 63);
        }
    std::vector<int> order;  // candidate pivot columns in the order they are tried
    if (parity_cols) {
        for (int i = 0; i < nparity; ++i) {
            if (parity_cols[i] < 0 || parity_cols[i] >= n) { set_error("parity column out of range"); return LDPC_ERR_ARG; }
            order.push_back(parity_cols[i]);
        }
    } else {
        for (int v = n - 1; v >= 0; --v) order.push_back(v);
    }
    std::vector<int> pivot_col;  // pivot column of row i after elimination (rows 0..rank-1)
    int rank = 0;
    for (int col : order) {
        if (rank == m) break;
        int sel = -1;
        for (int r = rank; r < m; ++r)
            if ((row[r][col >> 6] >> (col & 63)) & 1ull) { sel = r; break; }
        if (sel < 0) {
            if (parity_cols) continue;  // dependent within the requested set: tolerated, reported below
            continue;
        }
        std::swap(row[rank], row[sel]);
        for (int r = 0; r < m; ++r)
            if (r != rank && ((row[r][col >> 6] >> (col & 63)) & 1ull))
                for (int w = 0; w < words; ++w) row[r][w] ^= row[rank][w];
        pivot_col.push_back(col);
        ++rank;
    }
    if (parity_cols) {
        // the requested columns must explain every independent check
        for (int r = rank; r < m; ++r)
            for (int w = 0; w < words; ++w)
                if (row[r][w]) { set_error("the given parity columns do not span the row space of H"); return LDPC_ERR_ARG; }
    }
    ldpc_gen g;
    g.n = n; g.rows = rank;
    g.flag.assign(n, 0);
    for (int c : pivot_col) g.flag[c] = 1;
    for (int v = 0; v < n; ++v) (g.flag[v] ? g.parity_index : g.info_index).push_back(v);
    // equation i belongs to the i-th flagged column in ascending order (ArrayLDPC_Encoder.cpp:56-69)
    std::vector<int> row_of_col(n, -1);
    for (int i = 0; i < rank; ++i) row_of_col[pivot_col[i]] = i;
    g.eq.resize(rank);
    for (int i = 0; i < rank; ++i) {
        const std::vector<unsigned long long> &bits = row[row_of_col[g.parity_index[i]]];
        for (int v = 0; v < n; ++v)
            if ((bits[v >> 6] >> (v & 63)) & 1ull) g.eq[i].push_back(v);
    }
    out = std::move(g);
    return LDPC_OK;
}

int save_generator(const ldpc_gen &g, const char *path)
{
    FILE *f = std::fopen(path, "w");
    if (!f) { set_error(std::string("cannot write ") + path); return LDPC_ERR_IO; }
    size_t dmax = 0;
    for (const auto &e : g.eq) dmax = std::max(dmax, e.size());
    std::fprintf(f, "%d %d \n%d %d \n", g.n, g.rows, 0, (int)dmax);
    for (int v = 0; v < g.n; ++v) std::fprintf(f, "%d ", g.flag[v]);
    std::fprintf(f, "\n");
    for (const auto &e : g.eq) std::fprintf(f, "%d ", (int)e.size());
    std::fprintf(f, "\n");
    for (const auto &e : g.eq) {
        for (int v : e) std::fprintf(f, "%d ", v);
        std::fprintf(f, "\n");
    }
    std::fclose(f);
    return LDPC_OK;
}

// Reference: ArrayLDPC_Encoder.cpp:160-225.
int encode(const ldpc_gen &g, const char *info, int info_len, uint8_t *codeword)
{
    const int k = g.n - g.rows;
    if (!info || !codeword || info_len < 1 || (info_len - 1) * 8 + k % 8 < k) {
        set_error("info buffer shorter than k bits"); return LDPC_ERR_ARG;
    }
    std::vector<uint8_t> bit((size_t)info_len * 8);
    int cnt = 0;
    for (int i = 0; i < info_len - 1; ++i)
        for (int j = 0; j < 8; ++j) bit[cnt++] = (info[i] >> j) & 1;
    for (int j = 0; j < k % 8; ++j) bit[cnt++] = (info[info_len - 1] >> j) & 1;
    for (int i = 0; i < k; ++i) codeword[g.info_index[i]] = bit[i];
    for (int r = 0; r < g.rows; ++r) {
        uint8_t par = 0;
        for (int v : g.eq[r]) if (g.flag[v] == 0) par ^= codeword[v];
        codeword[g.parity_index[r]] = par;
    }
    return LDPC_OK;
}

}  // namespace ldpc
