// ldpc_decoder.cu -- device-resident decode engine behind the C ABI (include/ldpc_capi.h).
//
// Owns the device copies of the code tables, picks the kernel instantiation and launch shape
// for the code, and runs the packed int16x2 kernel with an exact int32 re-decode of the
// frames whose values left the packed guard range.  There is no CPU decode path here.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ldpc_capi.h"
#include "ldpc_code.hpp"
#include "ldpc_kernels.cuh"

namespace ldpc {

#define CUDA_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            ldpc::set_error(std::string(#expr) + ": " + cudaGetErrorString(e_));                 \
            return LDPC_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

typedef void (*kernel_fn)(const KParams);

struct KernelChoice {
    kernel_fn fn = nullptr;
    kernel_fn fn_ws = nullptr;  // the same kernel with the word-set count `ws` and the CTA size `ts` compiled in (named codes)
    int ws = 0, ts = 0;
    int max_threads = 0;
    int ni = 1;
    int cs = 1;  // checks of one word set interleaved per thread (CS of the kernel)
    int ctas_per_sm = 1;
    bool cdeg_in_smem = false, vdeg_in_smem = false;
};

template <class T, int DC, bool REG, int DV, int NI, int MAXT, int NCTA, int M, int N, int EA = 0, int ARRP = 0, unsigned VMASK = 0xffffffffu,
          unsigned CMASK = 0xffffffffu, int WS16 = 0, int WS32 = 0, int TS = 0, int CS = 1>
static KernelChoice make_choice()
{
    KernelChoice k;
    k.fn = decode_kernel<T, DC, REG, DV, NI, MAXT, NCTA, M, N, EA, ARRP, VMASK, CMASK, 0, 0, CS>;
    k.cs = CS;
    // word sets per CTA the plan arrives at without overrides (packed and int32 kernels have the same count: a word set
    // is E + n words either way)
    constexpr int WS = T::LANES == 2 ? WS16 : WS32;
    if (WS > 0) {
        k.fn_ws = decode_kernel<T, DC, REG, DV, NI, MAXT, NCTA, M, N, EA, ARRP, VMASK, CMASK, WS, TS, CS>;
        k.ws = WS;
        k.ts = TS;
    }
    k.max_threads = MAXT;
    k.ni = NI;
    k.ctas_per_sm = NCTA;
    // shared-memory copies of the degree tables (must mirror the kernel's REG / REGV conditions)
    k.cdeg_in_smem = !REG;
    k.vdeg_in_smem = !(PREFETCH_VEDGE && M != 0 && REG && DV <= 8) && ARRP == 0;
    return k;
}

// Forward square array code (class ROM, ArrayLDPCMacro.h:42-82): n = p*p, m = r*p, variable b*p + j meets check
// a*p + ((j - a*b) mod p) of every row group a and sits in slot b there.  Kernels instantiated with ARRP compute the
// edge addresses from this formula, so it is verified edge by edge before one of them is chosen.
static bool is_forward_array(const ldpc_code &c, int p)
{
    if (p <= 0 || c.n != p * p || c.m % p != 0 || c.dv_max != c.m / p || c.dc_max != p) return false;
    for (int v = 0; v < c.n; ++v) {
        if (c.vdeg[v] != c.dv_max) return false;
        const int b = v / p, j = v % p;
        for (int a = 0; a < c.dv_max; ++a) {
            const int t = ((j - a * b) % p + p) % p;
            if (c.vlist[(size_t)v * c.dv_max + a] != a * p + t || c.vslot[(size_t)v * c.dv_max + a] != b) return false;
        }
    }
    return true;
}

// Exact instantiations for the four named codes plus a generic bucket.  NI = word sets interleaved per thread
// in the check phase (instruction-level parallelism across independent chains); MAXT bounds the CTA so that the
// NI forward arrays stay in registers.
template <class T> static KernelChoice pick_kernel(const ldpc_code &c)
{
    bool regular = true;
    for (int d : c.cdeg) regular &= (d == c.dc_max);
    bool vregular = true;
    for (int d : c.vdeg) vregular &= (d == c.dv_max);
    // the four named codes: dimensions are compile-time constants (the REGV variable phase also assumes a
    // uniform variable degree, which all three array codes have)
    // Launch shapes below were chosen by measurement on B200 (profiles/r01/launch_shape_sweep.txt): for the long
    // checks more, smaller CTAs with one chain per thread beat two interleaved chains per thread (fewer registers,
    // half the unrolled code, four independent phase streams per SM); the short 802.11 checks prefer interleaved
    // chains, and with the compact word sets three of them fit a CTA (NI = 3: +5 % over NI = 2).
    if (regular && vregular && c.dc_max == 47 && c.dv_max == 5 && c.m == 235 && c.n == 2209) {        // array p47 r5
        // closed-form edge addresses measured +1 % at 30 iterations, +2 % at the operating point over the table with
        // its one-variable prefetch (profiles/r02/launch_shape_sweep.txt); LDPC_A5_TABLE=1 keeps the table
        // (two word sets per CTA on 480 threads, two CTAs per SM: same rate at 30 iterations, 3 % slower at 4.5 dB)
        if (!getenv("LDPC_A5_TABLE") && is_forward_array(c, 47))
            return make_choice<T, 47, true, 5, 1, 256, 4, 235, 2209, 0, 47, 0xffffffffu, 0xffffffffu, 1, 1, 256>();
        return make_choice<T, 47, true, 5, 1, 256, 4, 235, 2209>();
    }
    if (regular && vregular && c.dc_max == 47 && c.dv_max == 24 && c.m == 1128 && c.n == 2209 && is_forward_array(c, 47))  // array p47 r24
    {
        // One 221 KB word set per SM, 384 threads: 12 warps, three per scheduler.  Every thread interleaves three checks
        // (c, c + 376, c + 752: CS = 3, 164 registers, no spills), so the 1128 checks are one pass of 376 threads with
        // three dependency chains each: +2 % over one check per thread in three passes (128 registers), which
        // LDPC_A24_SPLIT=1 keeps; two checks per thread on 288 threads: -21 %; 576 threads / 96 registers: -15 %
        // (profiles/r02/launch_shape_sweep.txt).  LDPC_A24_640=1: the 96-register build.
        if (getenv("LDPC_A24_640")) return make_choice<T, 47, true, 24, 1, 640, 1, 1128, 2209, 0, 47>();
        if (const char *sp = getenv("LDPC_A24_SPLIT"))
            if (atoi(sp) == 1) return make_choice<T, 47, true, 24, 1, 512, 1, 1128, 2209, 0, 47, 0xffffffffu, 0xffffffffu, 1, 1, 384>();
        return make_choice<T, 47, true, 24, 1, 384, 1, 1128, 2209, 0, 47, 0xffffffffu, 0xffffffffu, 1, 1, 384, 3>();
    }
    if (regular && vregular && c.dc_max == 28 && c.dv_max == 4 && c.m == 316 && c.n == 2212)          // cut79
        return make_choice<T, 28, true, 4, 1, 640, 2, 316, 2212, 0, 0, 0xffffffffu, 0xffffffffu, 2, 2, 640>();
    int full = 0, last = 0;  // words per word set with the checks sorted by descending degree (see ldpc_decoder::e_words)
    for (int d : c.cdeg) { full += (d >= c.dc_max - 1); last += (d == c.dc_max); }
    if (c.dc_max == 8 && c.dv_max <= 12 && c.m == 972 && c.n == 1944 && full == 972 && last == 162) {  // 802.11n 1944 r1/2
        // exact bodies only for the degrees the code has (variables 2, 3, 4, 11; checks 7, 8): a third of the code size
        bool only = true;
        for (int d : c.vdeg) only &= (d == 2 || d == 3 || d == 4 || d == 11);
        if (only && !getenv("LDPC_WIFI_ALL_DEGREES"))
            return make_choice<T, 8, false, 12, 3, 512, 2, 972, 1944, 7 * 972 + 162, 0, (1u << 2) | (1u << 3) | (1u << 4) | (1u << 11), (1u << 7) | (1u << 8), 3, 3, 512>();
        return make_choice<T, 8, false, 12, 3, 512, 2, 972, 1944, 7 * 972 + 162>();  // three word sets per CTA and thread
    }
    // any other code: run-time dimensions
    if (regular && c.dc_max == 47 && c.dv_max <= 5) return make_choice<T, 47, true, 5, 1, 256, 4, 0, 0>();
    if (c.dc_max <= 8 && c.dv_max <= 12) return make_choice<T, 8, false, 12, 2, 512, 2, 0, 0>();
    if (c.dc_max <= 64 && c.dv_max <= 32) return make_choice<T, 64, false, 32, 1, 512, 1, 0, 0>();
    return KernelChoice();
}

struct Plan {
    KernelChoice kernel;
    int W = 0, threads = 0, smem = 0;  // smem: word sets + control block + degree tables
    // irregular codes: which variable / check every thread handles in pass k of the variable / check phase
    // ([k][thread], 0xffff = none), see build_order()
    uint16_t *d_vorder = nullptr, *d_corder = nullptr;
    int vorder_k = 0, corder_k = 0;
};

}  // namespace ldpc

using namespace ldpc;

struct ldpc_decoder {
    ldpc_code code;
    ldpc_decoder_cfg cfg;
    int device = 0, sm_count = 0, max_smem = 0;
    cudaStream_t stream = nullptr;
    // device tables
    uint8_t *d_cdeg = nullptr, *d_vdeg = nullptr;
    uint16_t *d_vedge = nullptr, *d_eorig = nullptr;
    // Checks are renumbered by descending degree inside the engine: with the slot-major layout word(slot, check) =
    // slot*m + check, the checks that have a slot-k edge are then 0..cnt_k-1 and the unused tail of the last slots
    // needs no shared memory (802.11: 6 966 words per word set instead of 8*972).  dev_cdeg = degrees in that order,
    // e_words = words per word set, d_eorig[word] = the reference's EdgeRAM index of the word (parity-mode dump).
    std::vector<int> dev_cdeg;
    int e_words = 0;
    unsigned long long *d_queue = nullptr;  // [2]: packed launch, int32 launch
    Plan plan16, plan32;
    // host-buffer path: two staging sets so the H2D copy of chunk i+1, the decode of chunk i and the D2H copy
    // of chunk i-1 overlap (copy-in / kernel / copy-out streams)
    void *d_llr[2] = {nullptr, nullptr}; int *d_iters[2] = {nullptr, nullptr}; uint32_t *d_bits[2] = {nullptr, nullptr};
    int *d_post[2] = {nullptr, nullptr}; int *d_v2c[2] = {nullptr, nullptr};
    size_t cap_frames = 0; bool cap_post = false, cap_v2c = false;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    // frames flagged by the packed kernel: list, length of the current call, running total
    int *d_fb_index = nullptr, *d_fb_count = nullptr;
    unsigned long long *d_fb_total = nullptr;
    size_t fb_cap = 0;
    // fed launches of the host pipeline: arrival mark, per-chunk completion counters, host-mapped completion flags
    unsigned long long *d_feed_avail = nullptr, *h_feed_marks = nullptr;
    unsigned int *d_done_count = nullptr, *h_done_flag = nullptr, *d_done_flag = nullptr;
    size_t feed_chunks_cap = 0, feed_marks_cap = 0;
    cudaEvent_t ev_feed = nullptr;
    // Monte-Carlo mode resources
    uint32_t *d_mc_pow = nullptr;      // a^(v+1) mod m
    uint32_t mc_jump = 0;              // a^n mod m
    uint32_t *d_mc_cw = nullptr, *d_mc_info = nullptr;  // [nw32] each
    int *d_mc_pin = nullptr; int mc_pin_cap = 0;
    int *d_mc_iters = nullptr; unsigned short *d_mc_ferr = nullptr; size_t mc_cap = 0;
    unsigned long long *d_mc_counters = nullptr;
    int *d_mc_llr = nullptr; size_t mc_llr_cap = 0;
    ldpc_decoder_stats stats;
};

namespace ldpc {

static int make_plan(const ldpc_decoder &d, int lanes, KernelChoice k, int want_threads, int want_slots, Plan &out)
{
    const ldpc_code &c = d.code;
    if (!k.fn) { set_error("check degree > 64 or variable degree > 32: no kernel instantiation"); return LDPC_ERR_UNSUPPORTED; }
    if ((long long)c.dc_max * c.m > 65535) { set_error("dc_max*m exceeds the 16-bit edge address space"); return LDPC_ERR_UNSUPPORTED; }
    const int per_w = (d.e_words + c.n) * 4;  // messages + channel values
    // co-resident CTAs share the SM's shared memory (228 KB minus 1 KB reserved per CTA)
    const int sm_total = d.max_smem + 1024;
    const int tables = (k.cdeg_in_smem ? ((c.m + 15) & ~15) : 0) + (k.vdeg_in_smem ? ((c.n + 15) & ~15) : 0);
    const int limit = k.ctas_per_sm > 1 ? sm_total / k.ctas_per_sm - 1024 : d.max_smem;
    const int budget = limit - (int)sizeof(Ctrl) - tables - 64;
    int W = std::min(budget / per_w, (int)MAX_W);
    if (want_slots > 0) {  // override, rounded up to whole groups of NI word sets
        int w = std::max(1, (want_slots + lanes - 1) / lanes);
        W = std::min(W, ((w + k.ni - 1) / k.ni) * k.ni);
    }
    if (W < k.ni) { set_error("the word sets one thread interleaves do not fit in shared memory"); return LDPC_ERR_UNSUPPORTED; }
    W -= W % k.ni;  // the check phase walks the word sets in groups of NI
    // CTA size: best check-phase lane efficiency, ties to the larger CTA
    const int items = (W / k.ni) * (c.m / k.cs);
    int best_t = 0; double best_e = -1;
    for (int t = 128; t <= k.max_threads; t += 32) {
        int passes = (items + t - 1) / t;
        double e = double(items) / (double(passes) * t);
        int vp = (c.n + t - 1) / t;
        e = 0.8 * e + 0.2 * double(c.n) / (double(vp) * t);
        if (e > best_e) best_e = e;
    }
    for (int t = 128; t <= k.max_threads; t += 32) {
        int passes = (items + t - 1) / t;
        double e = double(items) / (double(passes) * t);
        int vp = (c.n + t - 1) / t;
        e = 0.8 * e + 0.2 * double(c.n) / (double(vp) * t);
        if (e >= best_e - 0.03) best_t = t;  // largest CTA within 3% of the best lane efficiency
    }
    if (want_threads > 0) best_t = std::min(k.max_threads, std::max(32, (want_threads / 32) * 32));
    if (k.fn_ws && W == k.ws && (k.ts == 0 || best_t == k.ts) && !getenv("LDPC_RUNTIME_W")) k.fn = k.fn_ws;
    out.kernel = k; out.W = W; out.threads = best_t;
    out.smem = W * per_w + (int)sizeof(Ctrl) + tables;
    // the attribute belongs to the kernel instantiation, not to this decoder: decoders that share an instantiation
    // with different word-set counts would otherwise lower each other's limit
    cudaError_t e = cudaFuncSetAttribute((const void *)k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, d.max_smem);
    if (e != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return LDPC_ERR_CUDA; }
    return LDPC_OK;
}

// Work distribution for irregular codes.  Nodes are grouped into units of up to 32 nodes of the same degree (a warp
// that runs one unit executes one exact-degree body, no divergence); nodes keep their index order inside a degree
// class, so a unit is mostly a run of consecutive nodes (coalesced table reads, conflict-free shared-memory access).
// Units are dealt to the warps of the CTA longest-first onto the least loaded warp, so that all warps reach the
// barrier that ends the phase at about the same time (a degree-11 variable of the 802.11 code costs four times a
// degree-2 one).  Returns the [passes][threads] table, 0xffff = idle.
static std::vector<uint16_t> build_order(const std::vector<int> &deg, int threads, double base_cost, int &passes)
{
    const int n = (int)deg.size(), warps = threads / 32;
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return deg[a] > deg[b]; });
    struct Unit { int first, count, degree; };
    std::vector<Unit> units;
    for (int i = 0; i < n;) {
        int j = i;
        while (j < n && j - i < 32 && deg[idx[j]] == deg[idx[i]]) ++j;
        units.push_back({i, j - i, deg[idx[i]]});
        i = j;
    }
    passes = ((int)units.size() + warps - 1) / warps;
    std::vector<double> load(warps, 0.0);
    std::vector<int> used(warps, 0);
    std::vector<uint16_t> table((size_t)passes * threads, 0xffffu);
    for (const Unit &u : units) {  // already sorted by degree, largest first
        int best = -1;
        for (int w = 0; w < warps; ++w)
            if (used[w] < passes && (best < 0 || load[w] < load[best])) best = w;
        for (int l = 0; l < u.count; ++l) table[(size_t)used[best] * threads + best * 32 + l] = (uint16_t)idx[u.first + l];
        used[best]++;
        load[best] += base_cost + u.degree;
    }
    return table;
}

static int upload_order(const ldpc_decoder &d, Plan &pl)
{
    const ldpc_code &c = d.code;
    std::vector<uint16_t> v = build_order(c.vdeg, pl.threads, 1.5, pl.vorder_k);
    std::vector<uint16_t> k = build_order(d.dev_cdeg, pl.threads, 0.0, pl.corder_k);
    CUDA_TRY(cudaMalloc(&pl.d_vorder, v.size() * 2));
    CUDA_TRY(cudaMalloc(&pl.d_corder, k.size() * 2));
    CUDA_TRY(cudaMemcpy(pl.d_vorder, v.data(), v.size() * 2, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(pl.d_corder, k.data(), k.size() * 2, cudaMemcpyHostToDevice));
    return LDPC_OK;
}

static int upload_tables(ldpc_decoder &d)
{
    const ldpc_code &c = d.code;
    std::vector<int> order(c.m), inv(c.m);
    for (int i = 0; i < c.m; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return c.cdeg[a] > c.cdeg[b]; });
    for (int i = 0; i < c.m; ++i) inv[order[i]] = i;
    d.dev_cdeg.resize(c.m);
    std::vector<uint8_t> cdeg(c.m), vdeg(c.n);
    for (int i = 0; i < c.m; ++i) { d.dev_cdeg[i] = c.cdeg[order[i]]; cdeg[i] = (uint8_t)d.dev_cdeg[i]; }
    for (int v = 0; v < c.n; ++v) vdeg[v] = (uint8_t)c.vdeg[v];
    std::vector<uint16_t> vedge((size_t)c.dv_max * c.n, 0);
    d.e_words = 0;
    for (int v = 0; v < c.n; ++v)
        for (int j = 0; j < c.vdeg[v]; ++j) {
            int chk = c.vlist[(size_t)v * c.dv_max + j], slot = c.vslot[(size_t)v * c.dv_max + j];
            const int word = slot * c.m + inv[chk];
            vedge[(size_t)j * c.n + v] = (uint16_t)word;
            d.e_words = std::max(d.e_words, word + 1);
        }
    std::vector<uint16_t> eorig((size_t)std::max(d.e_words, 1), 0xffffu);
    for (int v = 0; v < c.n; ++v)
        for (int j = 0; j < c.vdeg[v]; ++j) {
            int chk = c.vlist[(size_t)v * c.dv_max + j], slot = c.vslot[(size_t)v * c.dv_max + j];
            eorig[vedge[(size_t)j * c.n + v]] = (uint16_t)(slot * c.m + chk);
        }
    CUDA_TRY(cudaMalloc(&d.d_eorig, eorig.size() * 2));
    CUDA_TRY(cudaMemcpy(d.d_eorig, eorig.data(), eorig.size() * 2, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&d.d_cdeg, cdeg.size()));
    CUDA_TRY(cudaMalloc(&d.d_vdeg, vdeg.size()));
    CUDA_TRY(cudaMalloc(&d.d_vedge, vedge.size() * 2));
    CUDA_TRY(cudaMalloc(&d.d_queue, 2 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&d.d_fb_count, sizeof(int)));
    CUDA_TRY(cudaMalloc(&d.d_fb_total, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(d.d_fb_total, 0, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemcpy(d.d_cdeg, cdeg.data(), cdeg.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.d_vdeg, vdeg.data(), vdeg.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.d_vedge, vedge.data(), vedge.size() * 2, cudaMemcpyHostToDevice));
    return LDPC_OK;
}

// ---- fallback plumbing: list the frames the packed kernel flagged (iters == -1) -------------
__global__ void collect_flagged(const int *iters, long long frames, int *index, int *count, unsigned long long *total)
{
    long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (f < frames && iters[f] < 0) {
        index[atomicAdd(count, 1)] = (int)f;
        atomicAdd(total, 1ull);
    }
}

// fed launch (host pipeline): the frames arrive while the kernel runs, and finished chunks are announced to the host
struct Feed {
    const unsigned long long *avail;
    unsigned int *done_count;
    volatile unsigned int *done_flag;
    int done_chunk;
};

static int launch(ldpc_decoder &d, const Plan &pl, int which, const void *llr, int llr_bits, long long frames,
                  int *iters, uint32_t *bits, int *post, int *v2c, const int *index, const int *count,
                  cudaStream_t st, const KParams *mc = nullptr, const Feed *fed = nullptr)
{
    if (frames <= 0) return LDPC_OK;
    const ldpc_code &c = d.code;
    KParams p;
    std::memset(&p, 0, sizeof p);
    p.cdeg = d.d_cdeg; p.vdeg = d.d_vdeg; p.vedge = d.d_vedge;
    p.vorder = pl.d_vorder; p.corder = pl.d_corder; p.vorder_k = pl.vorder_k; p.corder_k = pl.corder_k;
    p.eorig = d.d_eorig;
    p.n = c.n; p.m = c.m; p.E = d.e_words; p.dc_max = c.dc_max; p.dv_max = c.dv_max;
    p.W = pl.W; p.max_iter = d.cfg.max_iter; p.precheck = d.cfg.precheck;
    p.inv_m = (uint32_t)((1ull << 32) / (unsigned)c.m) + 1u;
    p.llr = llr; p.llr_bits = llr_bits; p.frames = frames;
    p.iters = iters; p.bits = bits; p.nw32 = (c.n + 31) / 32; p.post = post; p.v2c = v2c;
    p.queue = d.d_queue + which;
    p.index = index; p.count = count;
    p.mc_mode = 0;
    if (mc) {
        p.mc_mode = mc->mc_mode; p.mc_first = mc->mc_first; p.mc_seed = mc->mc_seed; p.mc_gain = mc->mc_gain;
        p.mc_sigma = mc->mc_sigma; p.mc_cw = mc->mc_cw; p.mc_cw_stride = mc->mc_cw_stride; p.mc_info = mc->mc_info; p.mc_pin = mc->mc_pin;
        p.mc_pin_count = mc->mc_pin_count; p.mc_pin_value = mc->mc_pin_value; p.mc_pow = mc->mc_pow;
        p.mc_jump = mc->mc_jump; p.mc_frame_err = mc->mc_frame_err; p.mc_counters = mc->mc_counters;
    }
    const int lanes = which == 0 ? 2 : 1;
    const long long slots = (long long)pl.W * lanes;
    int grid = (int)std::min<long long>((long long)d.sm_count * pl.kernel.ctas_per_sm, (frames + slots - 1) / slots);
    // Slots take their next frame from the queue one frame early (hides the atomic and lets the channel values be
    // prefetched into L2); with a short queue that would starve the CTAs that start last.  The length of an
    // indirect queue (re-decode list) is only known on the device.
    p.claim_ahead = (!count && frames >= 4 * slots * grid) ? 1 : 0;
    if (fed) { p.claim_ahead = 1; p.avail = fed->avail; p.done_count = fed->done_count; p.done_flag = fed->done_flag; p.done_chunk = fed->done_chunk; }
    const int smem = pl.smem;
    CUDA_TRY(cudaMemsetAsync(p.queue, 0, sizeof(unsigned long long), st));
    pl.kernel.fn<<<grid, pl.threads, smem, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    d.stats.kernel_launches++;
    d.stats.grid = grid;
    if (which == 0 || d.cfg.precision == 32) {
        int resident = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, (const void *)pl.kernel.fn, pl.threads, smem) == cudaSuccess)
            d.stats.resident_ctas_per_sm = resident;
        d.stats.launch_smem_bytes = smem;
    }
    return LDPC_OK;
}

// list of the frames the packed kernel flagged: sized before a pipeline starts, never in the middle of one
static int ensure_fallback_list(ldpc_decoder &d, size_t frames)
{
    if (d.fb_cap >= frames) return LDPC_OK;
    CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(d.d_fb_index);
    d.d_fb_index = nullptr; d.fb_cap = 0;
    CUDA_TRY(cudaMalloc(&d.d_fb_index, frames * sizeof(int)));
    d.fb_cap = frames;
    return LDPC_OK;
}

static int decode_device(ldpc_decoder &d, const void *llr, int llr_bits, long long frames, int *iters,
                         uint32_t *bits, int *post, int *v2c, cudaStream_t st, const KParams *mc = nullptr,
                         const Feed *fed = nullptr)
{
    if (!mc && llr_bits != 16 && llr_bits != 32) { set_error("llr_bits must be 16 or 32"); return LDPC_ERR_ARG; }
    if (frames > 0x7fffffffLL) { set_error("more than 2^31-1 frames in one call"); return LDPC_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(d.device));
    d.stats.frames += (uint64_t)frames;
    // parity-mode message dump: the variable phase writes the real edges of every iteration; the unused slots of
    // short rows read as zero like the reference's untouched EdgeRAM words
    if (v2c) CUDA_TRY(cudaMemsetAsync(v2c, 0, (size_t)frames * d.code.dc_max * d.code.m * sizeof(int), st));
    if (d.cfg.precision == 32)
        return launch(d, d.plan32, 1, llr, llr_bits, frames, iters, bits, post, v2c, nullptr, nullptr, st, mc, fed);
    int rc = launch(d, d.plan16, 0, llr, llr_bits, frames, iters, bits, post, v2c, nullptr, nullptr, st, mc, fed);
    if (rc != LDPC_OK || d.cfg.precision == 16) return rc;
    // Exact int32 re-decode of the flagged frames, in place and without a host round trip: the
    // second launch reads the frame list and its length from device memory.
    if (d.fb_cap < (size_t)frames) {  // device API with a larger batch than any before (the host pipeline sizes it up front)
        int rc2 = ensure_fallback_list(d, (size_t)frames);
        if (rc2 != LDPC_OK) return rc2;
    }
    CUDA_TRY(cudaMemsetAsync(d.d_fb_count, 0, sizeof(int), st));
    collect_flagged<<<(unsigned)((frames + 255) / 256), 256, 0, st>>>(iters, frames, d.d_fb_index, d.d_fb_count, d.d_fb_total);
    CUDA_TRY(cudaGetLastError());
    d.stats.kernel_launches++;
    return launch(d, d.plan32, 1, llr, llr_bits, frames, iters, bits, post, v2c, d.d_fb_index, d.d_fb_count, st, mc);
}

// ---- Monte-Carlo mode ------------------------------------------------------------------------
__global__ void channel_kernel(const KParams p, int *out)
{
    for (long long f = blockIdx.x; f < p.frames; f += gridDim.x) {
        const unsigned long long g = p.mc_first + (unsigned long long)f;
        int *row = out + (size_t)f * p.n;
        if (p.mc_mode == 1) {
            for (int q = threadIdx.x; 4 * q < p.n; q += blockDim.x) {
                float z[4];
                philox_normals(p, g, (uint32_t)q, z);
                for (int h = 0; h < 4; ++h)
                    if (4 * q + h < p.n) row[4 * q + h] = quantise_llr(p, (double)z[h], cw_bit(p, f, 4 * q + h));
            }
        } else {
            const uint32_t state = lehmer_mul((uint32_t)p.mc_seed, lehmer_pow(p.mc_jump, g));
            for (int v = threadIdx.x; v < p.n; v += blockDim.x)
                row[v] = quantise_llr(p, lehmer_normal(lehmer_mul(state, p.mc_pow[v])), cw_bit(p, f, v));
        }
        __syncthreads();
        for (int i = threadIdx.x; i < p.mc_pin_count; i += blockDim.x) row[p.mc_pin[i]] = p.mc_pin_value;
    }
}

static int mc_prepare(ldpc_decoder &d, const ldpc_mc_cfg &cfg, KParams &mc, cudaStream_t st)
{
    const ldpc_code &c = d.code;
    const int nw32 = (c.n + 31) / 32;
    if (cfg.stream != LDPC_STREAM_PHILOX && cfg.stream != LDPC_STREAM_REFERENCE) { set_error("unknown noise stream"); return LDPC_ERR_ARG; }
    if (!(cfg.snr > 0.0) || !(cfg.sigma > 0.0)) { set_error("snr and sigma must be positive"); return LDPC_ERR_ARG; }
    if (cfg.stream == LDPC_STREAM_REFERENCE && (cfg.seed == 0 || cfg.seed >= LEHMER_M)) {
        set_error("Lehmer seed must be in 1..2^31-2 (rngs.cpp:45)"); return LDPC_ERR_ARG;
    }
    if (!d.d_mc_pow) {
        std::vector<uint32_t> pw(c.n);
        uint32_t x = 1;
        for (int v = 0; v < c.n; ++v) { x = lehmer_mul(x, 48271u); pw[v] = x; }  // MULTIPLIER, rngs.cpp:41
        d.mc_jump = x;
        CUDA_TRY(cudaMalloc(&d.d_mc_pow, c.n * sizeof(uint32_t)));
        CUDA_TRY(cudaMemcpy(d.d_mc_pow, pw.data(), c.n * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMalloc(&d.d_mc_cw, nw32 * sizeof(uint32_t)));
        CUDA_TRY(cudaMalloc(&d.d_mc_info, nw32 * sizeof(uint32_t)));
        CUDA_TRY(cudaMalloc(&d.d_mc_counters, 4 * sizeof(unsigned long long)));
    }
    std::vector<uint32_t> cw(nw32, 0u), info(nw32, 0u);
    if (cfg.codeword)
        for (int v = 0; v < c.n; ++v) if (cfg.codeword[v] & 1) cw[v >> 5] |= 1u << (v & 31);
    if (cfg.info_index) {
        for (int i = 0; i < cfg.info_count; ++i) {
            const int v = cfg.info_index[i];
            if (v < 0 || v >= c.n) { set_error("info_index out of range"); return LDPC_ERR_ARG; }
            info[v >> 5] |= 1u << (v & 31);
        }
    } else {
        for (int v = 0; v < c.n; ++v) info[v >> 5] |= 1u << (v & 31);
    }
    CUDA_TRY(cudaMemcpyAsync(d.d_mc_cw, cw.data(), nw32 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d.d_mc_info, info.data(), nw32 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    if (cfg.pin_count > 0) {
        if (!cfg.pin_index) { set_error("pin_index is NULL"); return LDPC_ERR_ARG; }
        for (int i = 0; i < cfg.pin_count; ++i)
            if (cfg.pin_index[i] < 0 || cfg.pin_index[i] >= c.n) { set_error("pin_index out of range"); return LDPC_ERR_ARG; }
        if (d.mc_pin_cap < cfg.pin_count) {
            cudaFree(d.d_mc_pin); d.d_mc_pin = nullptr; d.mc_pin_cap = 0;
            CUDA_TRY(cudaMalloc(&d.d_mc_pin, cfg.pin_count * sizeof(int)));
            d.mc_pin_cap = cfg.pin_count;
        }
        CUDA_TRY(cudaMemcpyAsync(d.d_mc_pin, cfg.pin_index, cfg.pin_count * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));  // the staging vectors above go out of scope
    mc.mc_mode = cfg.stream; mc.mc_first = cfg.first_frame; mc.mc_seed = cfg.seed;
    mc.mc_gain = 2 * cfg.snr; mc.mc_sigma = cfg.sigma;
    mc.mc_cw = cfg.codeword ? d.d_mc_cw : nullptr; mc.mc_cw_stride = 0; mc.mc_info = d.d_mc_info;
    if (cfg.d_codewords) { mc.mc_cw = cfg.d_codewords; mc.mc_cw_stride = nw32; }
    mc.mc_pin = d.d_mc_pin; mc.mc_pin_count = cfg.pin_count > 0 ? cfg.pin_count : 0; mc.mc_pin_value = cfg.pin_value;
    mc.mc_pow = d.d_mc_pow; mc.mc_jump = d.mc_jump;
    mc.mc_frame_err = nullptr; mc.mc_counters = d.d_mc_counters;
    mc.n = c.n;
    return LDPC_OK;
}

static int mc_scratch(ldpc_decoder &d, size_t frames)
{
    if (d.mc_cap >= frames) return LDPC_OK;
    cudaFree(d.d_mc_iters); cudaFree(d.d_mc_ferr);
    d.d_mc_iters = nullptr; d.d_mc_ferr = nullptr; d.mc_cap = 0;
    CUDA_TRY(cudaMalloc(&d.d_mc_iters, frames * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d.d_mc_ferr, frames * sizeof(unsigned short)));
    d.mc_cap = frames;
    return LDPC_OK;
}


static void free_staging(ldpc_decoder &d)
{
    for (int b = 0; b < 2; ++b) {
        cudaFree(d.d_llr[b]); cudaFree(d.d_iters[b]); cudaFree(d.d_bits[b]); cudaFree(d.d_post[b]); cudaFree(d.d_v2c[b]);
        d.d_llr[b] = nullptr; d.d_iters[b] = nullptr; d.d_bits[b] = nullptr; d.d_post[b] = nullptr; d.d_v2c[b] = nullptr;
    }
    d.cap_frames = 0; d.cap_post = d.cap_v2c = false;
}

static int ensure_staging(ldpc_decoder &d, size_t frames, bool post, bool v2c)
{
    if (!d.s_in) {
        CUDA_TRY(cudaStreamCreateWithFlags(&d.s_in, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&d.s_out, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            CUDA_TRY(cudaEventCreateWithFlags(&d.ev_in[b], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&d.ev_k[b], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&d.ev_out[b], cudaEventDisableTiming));
        }
    }
    if (d.cap_frames >= frames && (!post || d.cap_post) && (!v2c || d.cap_v2c)) return LDPC_OK;
    const ldpc_code &c = d.code;
    size_t cap = std::max(frames, d.cap_frames);
    post |= d.cap_post; v2c |= d.cap_v2c;
    CUDA_TRY(cudaDeviceSynchronize());
    free_staging(d);
    for (int b = 0; b < 2; ++b) {
        CUDA_TRY(cudaMalloc(&d.d_llr[b], cap * c.n * sizeof(int)));
        CUDA_TRY(cudaMalloc(&d.d_iters[b], cap * sizeof(int)));
        CUDA_TRY(cudaMalloc(&d.d_bits[b], cap * ((c.n + 31) / 32) * sizeof(uint32_t)));
        if (post) CUDA_TRY(cudaMalloc(&d.d_post[b], cap * c.n * sizeof(int)));
        if (v2c) CUDA_TRY(cudaMalloc(&d.d_v2c[b], cap * (size_t)c.dc_max * c.m * sizeof(int)));
    }
    d.cap_frames = cap; d.cap_post = post; d.cap_v2c = v2c;
    return ensure_fallback_list(d, cap);
}

}  // namespace ldpc

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char *ldpc_strerror(int status)
{
    switch (status) {
    case LDPC_OK: return "ok";
    case LDPC_ERR_IO: return "cannot open or read file";
    case LDPC_ERR_FORMAT: return "malformed parity-check description";
    case LDPC_ERR_ARG: return "invalid argument";
    case LDPC_ERR_CUDA: return "CUDA runtime error";
    case LDPC_ERR_UNSUPPORTED: return "code does not fit the decode kernels";
    case LDPC_ERR_NOMEM: return "out of memory";
    case LDPC_ERR_NO_DEVICE: return "no CUDA device (this engine has no CPU path)";
    default: return "unknown status";
    }
}

const char *ldpc_last_error(void) { return ldpc::last_error(); }

static ldpc_code *finish_code(int st, ldpc_code *c, int *err)
{
    if (err) *err = st;
    if (st != LDPC_OK) { delete c; return nullptr; }
    return c;
}

ldpc_code *ldpc_code_load(const char *path, int format, int *err)
{
    ldpc_code *c = new ldpc_code;
    return finish_code(ldpc::load_file(path, format, *c), c, err);
}

ldpc_code *ldpc_code_from_checks(int n, int m, const int *cdeg, const int *clist, int cstride, int *err)
{
    ldpc_code *c = new ldpc_code;
    return finish_code(ldpc::build_from_checks(n, m, cdeg, clist, cstride, *c), c, err);
}

ldpc_code *ldpc_code_array(int p, int nrows, const int *row_mult, int ncols, const int *col_sel, int backward, int *err)
{
    ldpc_code *c = new ldpc_code;
    return finish_code(ldpc::build_array(p, nrows, row_mult, ncols, col_sel, backward, *c), c, err);
}

void ldpc_code_free(ldpc_code *code) { delete code; }

ldpc_gen *ldpc_gen_load(const char *path, int *err)
{
    ldpc_gen *g = new ldpc_gen;
    int st = ldpc::load_generator(path, *g);
    if (err) *err = st;
    if (st != LDPC_OK) { delete g; return nullptr; }
    return g;
}

ldpc_gen *ldpc_gen_from_code(const ldpc_code *code, const int32_t *parity_cols, int nparity, int *err)
{
    if (!code) { if (err) *err = LDPC_ERR_ARG; return nullptr; }
    ldpc_gen *g = new ldpc_gen;
    int st = ldpc::derive_generator(*code, parity_cols, nparity, *g);
    if (err) *err = st;
    if (st != LDPC_OK) { delete g; return nullptr; }
    return g;
}

int ldpc_gen_save(const ldpc_gen *g, const char *path)
{
    if (!g || !path) return LDPC_ERR_ARG;
    return ldpc::save_generator(*g, path);
}

void ldpc_gen_release_device(ldpc_gen *gen);
void ldpc_gen_free(ldpc_gen *gen)
{
    ldpc_gen_release_device(gen);
    delete gen;
}

int ldpc_gen_dims(const ldpc_gen *g, int *n, int *rows, int *k)
{
    if (!g) return LDPC_ERR_ARG;
    if (n) *n = g->n;
    if (rows) *rows = g->rows;
    if (k) *k = g->n - g->rows;
    return LDPC_OK;
}

int ldpc_gen_indices(const ldpc_gen *g, int32_t *info_index, int32_t *parity_index)
{
    if (!g) return LDPC_ERR_ARG;
    if (info_index) std::copy(g->info_index.begin(), g->info_index.end(), info_index);
    if (parity_index) std::copy(g->parity_index.begin(), g->parity_index.end(), parity_index);
    return LDPC_OK;
}

int ldpc_gen_encode(const ldpc_gen *g, const char *info, int info_len, uint8_t *codeword)
{
    if (!g) return LDPC_ERR_ARG;
    return ldpc::encode(*g, info, info_len, codeword);
}

int ldpc_code_dims(const ldpc_code *c, int *n, int *m, int *edges, int *dc_max, int *dv_max)
{
    if (!c) return LDPC_ERR_ARG;
    if (n) *n = c->n;
    if (m) *m = c->m;
    if (edges) *edges = c->edges;
    if (dc_max) *dc_max = c->dc_max;
    if (dv_max) *dv_max = c->dv_max;
    return LDPC_OK;
}

int ldpc_code_tables(const ldpc_code *c, int *vdeg, int *cdeg, int *vlist, int *clist)
{
    if (!c) return LDPC_ERR_ARG;
    if (vdeg) std::copy(c->vdeg.begin(), c->vdeg.end(), vdeg);
    if (cdeg) std::copy(c->cdeg.begin(), c->cdeg.end(), cdeg);
    if (vlist) std::copy(c->vlist.begin(), c->vlist.end(), vlist);
    if (clist) std::copy(c->clist.begin(), c->clist.end(), clist);
    return LDPC_OK;
}

double ldpc_code_rate(const ldpc_code *c) { return c ? c->rate : 0.0; }

int ldpc_code_save(const ldpc_code *c, const char *path)
{
    if (!c || !path) return LDPC_ERR_ARG;
    return ldpc::save_format_a(*c, path);
}

void ldpc_decoder_cfg_default(ldpc_decoder_cfg *cfg)
{
    if (!cfg) return;
    cfg->max_iter = 30;  // MAX_ITER, ArrayLDPCMacro.h:17
    cfg->precheck = 0;
    cfg->device = 0;
    cfg->precision = 0;
    cfg->threads = 0;
    cfg->frames_per_cta = 0;
}

int ldpc_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { ldpc::set_error(cudaGetErrorString(e)); cudaGetLastError(); return LDPC_ERR_NO_DEVICE; }
    return n;
}

ldpc_decoder *ldpc_decoder_create(const ldpc_code *code, const ldpc_decoder_cfg *cfg_in, int *err)
{
    int st = LDPC_OK;
    ldpc_decoder *d = nullptr;
    ldpc_decoder_cfg cfg;
    ldpc_decoder_cfg_default(&cfg);
    if (cfg_in) cfg = *cfg_in;
    auto fail = [&](int s) { if (err) *err = s; if (d) ldpc_decoder_destroy(d); return (ldpc_decoder *)nullptr; };
    if (!code) { ldpc::set_error("NULL code"); return fail(LDPC_ERR_ARG); }
    if (cfg.max_iter < 1 || (cfg.precision != 0 && cfg.precision != 16 && cfg.precision != 32)) {
        ldpc::set_error("max_iter < 1 or unknown precision"); return fail(LDPC_ERR_ARG);
    }
    int ndev = ldpc_device_count();
    if (ndev <= 0) { if (ndev == 0) ldpc::set_error("no CUDA device visible"); return fail(LDPC_ERR_NO_DEVICE); }
    if (cfg.device < 0 || cfg.device >= ndev) { ldpc::set_error("device ordinal out of range"); return fail(LDPC_ERR_ARG); }
    d = new ldpc_decoder;
    d->code = *code; d->cfg = cfg; d->device = cfg.device;
    std::memset(&d->stats, 0, sizeof d->stats);
    cudaError_t e = cudaSetDevice(cfg.device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, cfg.device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&d->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg.device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { ldpc::set_error(cudaGetErrorString(e)); return fail(LDPC_ERR_CUDA); }
    if ((st = ldpc::upload_tables(*d)) != LDPC_OK) return fail(st);
    if ((st = ldpc::make_plan(*d, 2, ldpc::pick_kernel<ldpc::Packed16>(d->code), cfg.threads, cfg.frames_per_cta, d->plan16)) != LDPC_OK) return fail(st);
    if ((st = ldpc::make_plan(*d, 1, ldpc::pick_kernel<ldpc::Scalar32>(d->code), cfg.threads, cfg.frames_per_cta, d->plan32)) != LDPC_OK) return fail(st);
    if ((st = ldpc::upload_order(*d, d->plan16)) != LDPC_OK) return fail(st);
    if ((st = ldpc::upload_order(*d, d->plan32)) != LDPC_OK) return fail(st);
    d->stats.threads = d->plan16.threads; d->stats.threads32 = d->plan32.threads;
    d->stats.frames_per_cta = d->plan16.W * 2; d->stats.frames_per_cta32 = d->plan32.W;
    d->stats.smem_bytes = d->plan16.smem; d->stats.smem_bytes32 = d->plan32.smem;
    if (err) *err = LDPC_OK;
    return d;
}

void ldpc_decoder_destroy(ldpc_decoder *d)
{
    if (!d) return;
    cudaSetDevice(d->device);
    if (d->stream) cudaStreamSynchronize(d->stream);
    ldpc::free_staging(*d);
    for (int b = 0; b < 2; ++b) {
        if (d->ev_in[b]) cudaEventDestroy(d->ev_in[b]);
        if (d->ev_k[b]) cudaEventDestroy(d->ev_k[b]);
        if (d->ev_out[b]) cudaEventDestroy(d->ev_out[b]);
    }
    if (d->s_in) cudaStreamDestroy(d->s_in);
    if (d->s_out) cudaStreamDestroy(d->s_out);
    cudaFree(d->d_cdeg); cudaFree(d->d_vdeg); cudaFree(d->d_vedge); cudaFree(d->d_eorig);
    cudaFree(d->plan16.d_vorder); cudaFree(d->plan16.d_corder); cudaFree(d->plan32.d_vorder); cudaFree(d->plan32.d_corder); cudaFree(d->d_queue);
    cudaFree(d->d_fb_index); cudaFree(d->d_fb_count); cudaFree(d->d_fb_total);
    cudaFree(d->d_feed_avail); cudaFree(d->d_done_count);
    if (d->h_feed_marks) cudaFreeHost(d->h_feed_marks);
    if (d->h_done_flag) cudaFreeHost(d->h_done_flag);
    if (d->ev_feed) cudaEventDestroy(d->ev_feed);
    cudaFree(d->d_mc_pow); cudaFree(d->d_mc_cw); cudaFree(d->d_mc_info); cudaFree(d->d_mc_pin);
    cudaFree(d->d_mc_iters); cudaFree(d->d_mc_ferr); cudaFree(d->d_mc_counters); cudaFree(d->d_mc_llr);
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d;
}

int ldpc_decode_batch_device(ldpc_decoder *d, const void *d_llr, int llr_bits, size_t frames, int32_t *d_iters,
                             uint32_t *d_bits, int32_t *d_post, int32_t *d_v2c, void *stream)
{
    if (!d || !d_llr || !d_iters) { ldpc::set_error("NULL decoder / llr / iters"); return LDPC_ERR_ARG; }
    cudaStream_t st = stream ? (cudaStream_t)stream : d->stream;
    return ldpc::decode_device(*d, d_llr, llr_bits, (long long)frames, d_iters, d_bits, d_post, d_v2c, st);
}

// Host pipeline for long batches: ONE persistent decode launch per (up to 2^18-frame) batch.  The kernel starts at
// once; the channel values follow in growing chunks on the copy-in stream, each chunk followed by an 8-byte copy
// that moves the arrival mark the kernel polls (ldpc_kernels.cuh: KParams::avail); finished frames are counted per
// result chunk on the device, the thread that completes a chunk raises a flag in host-mapped memory, and this
// thread then copies that chunk's results out on the copy-out stream while the kernel decodes on.  No launch tails
// between chunks, no idle slots while a chunk drains.
static int decode_host_fed(ldpc_decoder *d, const char *llr, int llr_bits, size_t frames, int32_t *iters, uint32_t *bits)
{
    const ldpc_code &c = d->code;
    const size_t esz = (size_t)llr_bits / 8, nw32 = (c.n + 31) / 32;
    const size_t slots = (size_t)d->sm_count * d->plan16.kernel.ctas_per_sm * d->plan16.W * 2;
    const size_t done_chunk = 8192;
    int rc = ldpc::ensure_staging(*d, frames, false, false);
    if (rc != LDPC_OK) return rc;
    const size_t nchunks = (frames + done_chunk - 1) / done_chunk;
    // copy-in chunks of at most 4096 frames.  The
    // arrival mark trails the copy by up to a chunk, so chunks stay small: on a box whose copy-in rate is only a little
    // above the decode rate (eight GPUs copying at once: 23 GB/s per GPU against 55 GB/s alone) a kernel that has caught
    // up with the mark idles for the rest of the chunk in flight (doubling chunks up to 32 768 frames cost 15 % there).
    std::vector<size_t> marks;
    // The first chunks are short (256 frames, doubling up to 4096) so that the first CTAs start decoding some tens
    // of microseconds after the launch instead of after one frame per slot has arrived.
    const bool ramp = !getenv("LDPC_FEED_NO_RAMP");
    for (size_t at = 0, sz = ramp ? 256 : std::max<size_t>(slots, 1024); at < frames; sz = ramp ? std::min<size_t>(4096, sz * 2) : 4096) {
        at = std::min(frames, at + sz);
        marks.push_back(at);
    }
    if (!d->d_feed_avail) {
        CUDA_TRY(cudaMalloc(&d->d_feed_avail, sizeof(unsigned long long)));
        CUDA_TRY(cudaEventCreateWithFlags(&d->ev_feed, cudaEventDisableTiming));
    }
    if (d->feed_chunks_cap < nchunks) {
        CUDA_TRY(cudaDeviceSynchronize());
        cudaFree(d->d_done_count); d->d_done_count = nullptr;
        if (d->h_done_flag) cudaFreeHost(d->h_done_flag);
        d->h_done_flag = nullptr; d->feed_chunks_cap = 0;
        CUDA_TRY(cudaMalloc(&d->d_done_count, nchunks * sizeof(unsigned int)));
        CUDA_TRY(cudaHostAlloc(&d->h_done_flag, nchunks * sizeof(unsigned int), cudaHostAllocMapped));
        CUDA_TRY(cudaHostGetDevicePointer(&d->d_done_flag, d->h_done_flag, 0));
        d->feed_chunks_cap = nchunks;
    }
    if (d->feed_marks_cap < marks.size()) {
        CUDA_TRY(cudaDeviceSynchronize());
        if (d->h_feed_marks) cudaFreeHost(d->h_feed_marks);
        d->h_feed_marks = nullptr; d->feed_marks_cap = 0;
        CUDA_TRY(cudaHostAlloc(&d->h_feed_marks, marks.size() * sizeof(unsigned long long), cudaHostAllocDefault));
        d->feed_marks_cap = marks.size();
    }
    for (size_t i = 0; i < marks.size(); ++i) d->h_feed_marks[i] = marks[i];
    volatile unsigned int *flags = d->h_done_flag;
    for (size_t i = 0; i < nchunks; ++i) flags[i] = 0u;
    cudaStream_t sk = d->stream;
    CUDA_TRY(cudaMemsetAsync(d->d_feed_avail, 0, sizeof(unsigned long long), sk));
    CUDA_TRY(cudaMemsetAsync(d->d_done_count, 0, nchunks * sizeof(unsigned int), sk));
    CUDA_TRY(cudaEventRecord(d->ev_feed, sk));
    ldpc::Feed fed{d->d_feed_avail, d->d_done_count, d->d_done_flag, (int)done_chunk};
    const int saved_precision = d->cfg.precision;
    if (d->cfg.precision == 0) d->cfg.precision = 16;  // the re-decode of flagged frames follows below, after the fed launch
    rc = ldpc::decode_device(*d, d->d_llr[0], llr_bits, (long long)frames, d->d_iters[0], bits ? d->d_bits[0] : nullptr,
                             nullptr, nullptr, sk, nullptr, &fed);
    d->cfg.precision = saved_precision;
    if (rc != LDPC_OK) return rc;
    // the channel values follow the launch
    CUDA_TRY(cudaStreamWaitEvent(d->s_in, d->ev_feed, 0));
    size_t at = 0;
    cudaError_t e = cudaSuccess;
    for (size_t i = 0; i < marks.size() && e == cudaSuccess; ++i) {
        e = cudaMemcpyAsync((char *)d->d_llr[0] + at * c.n * esz, llr + at * c.n * esz, (marks[i] - at) * c.n * esz,
                            cudaMemcpyHostToDevice, d->s_in);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(d->d_feed_avail, &d->h_feed_marks[i], sizeof(unsigned long long), cudaMemcpyHostToDevice, d->s_in);
        at = marks[i];
    }
    if (e != cudaSuccess) {
        // let the kernel run out instead of waiting for frames that will not come
        unsigned long long all = frames;
        cudaMemcpy(d->d_feed_avail, &all, sizeof all, cudaMemcpyHostToDevice);
        cudaStreamSynchronize(sk);
        ldpc::set_error(std::string("host pipeline copy-in: ") + cudaGetErrorString(e));
        return LDPC_ERR_CUDA;
    }
    // results leave chunk by chunk as the kernel announces them
    bool kernel_done = false;
    for (size_t ch = 0; ch < nchunks; ++ch) {
        for (unsigned spins = 0; !flags[ch] && !kernel_done; ++spins) {
            if ((spins & 1023u) == 1023u) {
                cudaError_t q = cudaStreamQuery(sk);
                if (q == cudaSuccess) kernel_done = true;
                else if (q != cudaErrorNotReady) { ldpc::set_error(std::string("decode kernel: ") + cudaGetErrorString(q)); return LDPC_ERR_CUDA; }
            }
        }
        const size_t lo = ch * done_chunk, cnt = std::min(done_chunk, frames - lo);
        CUDA_TRY(cudaMemcpyAsync(iters + lo, d->d_iters[0] + lo, cnt * sizeof(int), cudaMemcpyDeviceToHost, d->s_out));
        if (bits) CUDA_TRY(cudaMemcpyAsync(bits + lo * nw32, d->d_bits[0] + lo * nw32, cnt * nw32 * sizeof(uint32_t), cudaMemcpyDeviceToHost, d->s_out));
    }
    CUDA_TRY(cudaStreamSynchronize(sk));
    if (saved_precision == 0) {
        // exact int32 re-decode of the frames the packed kernel flagged (none on channel-generated frames): an
        // ordinary launch over the device copy of the batch; its results replace the flagged ones
        CUDA_TRY(cudaMemsetAsync(d->d_fb_count, 0, sizeof(int), sk));
        ldpc::collect_flagged<<<(unsigned)((frames + 255) / 256), 256, 0, sk>>>(d->d_iters[0], (long long)frames, d->d_fb_index,
                                                                                  d->d_fb_count, d->d_fb_total);
        CUDA_TRY(cudaGetLastError());
        d->stats.kernel_launches++;
        int flagged = 0;
        CUDA_TRY(cudaMemcpyAsync(&flagged, d->d_fb_count, sizeof(int), cudaMemcpyDeviceToHost, sk));
        CUDA_TRY(cudaStreamSynchronize(sk));
        if (flagged > 0) {
            rc = ldpc::launch(*d, d->plan32, 1, d->d_llr[0], llr_bits, (long long)frames, d->d_iters[0], bits ? d->d_bits[0] : nullptr,
                              nullptr, nullptr, d->d_fb_index, d->d_fb_count, sk);
            if (rc != LDPC_OK) return rc;
            CUDA_TRY(cudaStreamSynchronize(d->s_out));
            CUDA_TRY(cudaMemcpyAsync(iters, d->d_iters[0], frames * sizeof(int), cudaMemcpyDeviceToHost, sk));
            if (bits) CUDA_TRY(cudaMemcpyAsync(bits, d->d_bits[0], frames * nw32 * sizeof(uint32_t), cudaMemcpyDeviceToHost, sk));
            CUDA_TRY(cudaStreamSynchronize(sk));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(d->s_out));
    return LDPC_OK;
}

static int decode_host(ldpc_decoder *d, const void *llr_v, int llr_bits, size_t frames, int32_t *iters, uint32_t *bits,
                       int32_t *post, int32_t *v2c)
{
    if (!d || !llr_v || !iters) { ldpc::set_error("NULL decoder / llr / iters"); return LDPC_ERR_ARG; }
    const char *llr = static_cast<const char *>(llr_v);
    const size_t esz = (size_t)llr_bits / 8;
    if (frames == 0) return LDPC_OK;
    CUDA_TRY(cudaSetDevice(d->device));
    const ldpc_code &c = d->code;
    const size_t nw32 = (c.n + 31) / 32, vsz = (size_t)c.dc_max * c.m;
    {
        // long batches without parity-mode outputs: the fed pipeline, 2^18 frames per launch
        const ldpc::Plan &pl = d->cfg.precision == 32 ? d->plan32 : d->plan16;
        const size_t per_grid = (size_t)d->sm_count * pl.kernel.ctas_per_sm * pl.W * (d->cfg.precision == 32 ? 1 : 2);
        if (!post && !v2c && d->cfg.max_iter > 0 && frames >= 8 * per_grid && !getenv("LDPC_NO_FEED")) {
            size_t batch = (size_t)1 << 18;
            if (const char *env = getenv("LDPC_FEED_BATCH"))  // (tests exercise the batch loop with small batches)
                if (atol(env) > 0) batch = std::max<size_t>((size_t)atol(env), 8 * per_grid);
            for (size_t base = 0; base < frames; base += batch) {
                const size_t cnt = std::min(batch, frames - base);
                int rc = decode_host_fed(d, llr + base * c.n * esz, llr_bits, cnt, iters + base, bits ? bits + base * nw32 : nullptr);
                if (rc != LDPC_OK) return rc;
            }
            return LDPC_OK;
        }
    }
    // Chunked pipeline: the H2D copy of chunk i+1, the decode of chunk i and the D2H copy of chunk i-1 overlap
    // (pinned host memory makes the copies asynchronous).  The first chunk is one frame per slot so that the
    // decode starts after a short copy; chunks then grow by half (the copy engine outruns the decoder by more than
    // that) up to a few frames per slot, which keeps the idle slots at the end of every launch a small share.
    const size_t slots = (size_t)d->sm_count * d->plan16.kernel.ctas_per_sm * d->plan16.W * 2;
    size_t max_chunk = std::max<size_t>(slots * 4, 4096), first_chunk = std::max<size_t>(slots, 1024);
    if (post || v2c) max_chunk = first_chunk = std::max<size_t>(slots, 1024);  // parity-mode outputs are large
    max_chunk = std::min(max_chunk, frames);
    first_chunk = std::min(first_chunk, max_chunk);
    int rc = ldpc::ensure_staging(*d, max_chunk, post != nullptr, v2c != nullptr);
    if (rc != LDPC_OK) return rc;
    cudaStream_t sk = d->stream;
    size_t index = 0;
    size_t chunk = first_chunk;
    for (size_t base = 0; base < frames; base += chunk, chunk = std::min(max_chunk, chunk + chunk / 2), ++index) {
        const size_t cnt = std::min(chunk, frames - base);
        const int b = (int)(index & 1);
        if (index >= 2) CUDA_TRY(cudaStreamWaitEvent(d->s_in, d->ev_k[b], 0));   // decode of chunk i-2 consumed this buffer
        CUDA_TRY(cudaMemcpyAsync(d->d_llr[b], llr + base * c.n * esz, cnt * c.n * esz, cudaMemcpyHostToDevice, d->s_in));
        CUDA_TRY(cudaEventRecord(d->ev_in[b], d->s_in));
        CUDA_TRY(cudaStreamWaitEvent(sk, d->ev_in[b], 0));
        if (index >= 2) CUDA_TRY(cudaStreamWaitEvent(sk, d->ev_out[b], 0));      // results of chunk i-2 have left
        rc = ldpc::decode_device(*d, d->d_llr[b], llr_bits, (long long)cnt, d->d_iters[b], bits ? d->d_bits[b] : nullptr,
                                 post ? d->d_post[b] : nullptr, v2c ? d->d_v2c[b] : nullptr, sk);
        if (rc != LDPC_OK) return rc;
        CUDA_TRY(cudaEventRecord(d->ev_k[b], sk));
        CUDA_TRY(cudaStreamWaitEvent(d->s_out, d->ev_k[b], 0));
        CUDA_TRY(cudaMemcpyAsync(iters + base, d->d_iters[b], cnt * sizeof(int), cudaMemcpyDeviceToHost, d->s_out));
        if (bits) CUDA_TRY(cudaMemcpyAsync(bits + base * nw32, d->d_bits[b], cnt * nw32 * sizeof(uint32_t), cudaMemcpyDeviceToHost, d->s_out));
        if (post) CUDA_TRY(cudaMemcpyAsync(post + base * c.n, d->d_post[b], cnt * c.n * sizeof(int), cudaMemcpyDeviceToHost, d->s_out));
        if (v2c) CUDA_TRY(cudaMemcpyAsync(v2c + base * vsz, d->d_v2c[b], cnt * vsz * sizeof(int), cudaMemcpyDeviceToHost, d->s_out));
        CUDA_TRY(cudaEventRecord(d->ev_out[b], d->s_out));
    }
    CUDA_TRY(cudaStreamSynchronize(d->s_out));
    CUDA_TRY(cudaStreamSynchronize(sk));
    return LDPC_OK;
}

int ldpc_decode_batch(ldpc_decoder *d, const int32_t *llr, size_t frames, int32_t *iters, uint32_t *bits,
                      int32_t *post, int32_t *v2c)
{
    return decode_host(d, llr, 32, frames, iters, bits, post, v2c);
}

int ldpc_decode_batch_i16(ldpc_decoder *d, const int16_t *llr, size_t frames, int32_t *iters, uint32_t *bits,
                          int32_t *post, int32_t *v2c)
{
    return decode_host(d, llr, 16, frames, iters, bits, post, v2c);
}

int ldpc_mc_run_device(ldpc_decoder *d, const ldpc_mc_cfg *cfg, size_t frames, uint16_t *d_frame_err,
                       int32_t *d_iters, uint64_t *d_counters, void *stream)
{
    if (!d || !cfg) { ldpc::set_error("NULL decoder / cfg"); return LDPC_ERR_ARG; }
    if (frames == 0) return LDPC_OK;
    CUDA_TRY(cudaSetDevice(d->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->stream;
    ldpc::KParams mc;
    std::memset(&mc, 0, sizeof mc);
    int rc = ldpc::mc_prepare(*d, *cfg, mc, st);
    if (rc != LDPC_OK) return rc;
    if (!d_iters) {  // the fallback list is built from the iteration counts, so they always exist
        if ((rc = ldpc::mc_scratch(*d, frames)) != LDPC_OK) return rc;
        d_iters = d->d_mc_iters;
    }
    mc.mc_frame_err = d_frame_err;
    if (d_counters) mc.mc_counters = (unsigned long long *)d_counters;
    else CUDA_TRY(cudaMemsetAsync(d->d_mc_counters, 0, 4 * sizeof(unsigned long long), st));
    return ldpc::decode_device(*d, nullptr, 32, (long long)frames, d_iters, nullptr, nullptr, nullptr, st, &mc);
}

int ldpc_mc_run(ldpc_decoder *d, const ldpc_mc_cfg *cfg, size_t frames, uint16_t *frame_err, int32_t *iters,
                ldpc_mc_counters *totals)
{
    if (!d || !cfg) { ldpc::set_error("NULL decoder / cfg"); return LDPC_ERR_ARG; }
    if (totals) std::memset(totals, 0, sizeof *totals);
    if (frames == 0) return LDPC_OK;
    CUDA_TRY(cudaSetDevice(d->device));
    int rc = ldpc::mc_scratch(*d, frames);
    if (rc != LDPC_OK) return rc;
    cudaStream_t st = d->stream;
    rc = ldpc_mc_run_device(d, cfg, frames, d->d_mc_ferr, d->d_mc_iters, nullptr, st);
    if (rc != LDPC_OK) return rc;
    if (frame_err) CUDA_TRY(cudaMemcpyAsync(frame_err, d->d_mc_ferr, frames * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
    if (iters) CUDA_TRY(cudaMemcpyAsync(iters, d->d_mc_iters, frames * sizeof(int), cudaMemcpyDeviceToHost, st));
    unsigned long long c[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(c, d->d_mc_counters, sizeof c, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (totals) { totals->frames = c[0]; totals->frame_errors = c[1]; totals->bit_errors = c[2]; totals->iter_sum = c[3]; }
    return LDPC_OK;
}

int ldpc_mc_channel(ldpc_decoder *d, const ldpc_mc_cfg *cfg, size_t frames, int32_t *llr)
{
    if (!d || !cfg || !llr) { ldpc::set_error("NULL decoder / cfg / llr"); return LDPC_ERR_ARG; }
    if (frames == 0) return LDPC_OK;
    CUDA_TRY(cudaSetDevice(d->device));
    cudaStream_t st = d->stream;
    ldpc::KParams mc;
    std::memset(&mc, 0, sizeof mc);
    int rc = ldpc::mc_prepare(*d, *cfg, mc, st);
    if (rc != LDPC_OK) return rc;
    if (d->mc_llr_cap < frames) {
        cudaFree(d->d_mc_llr); d->d_mc_llr = nullptr; d->mc_llr_cap = 0;
        CUDA_TRY(cudaMalloc(&d->d_mc_llr, frames * d->code.n * sizeof(int)));
        d->mc_llr_cap = frames;
    }
    mc.frames = (long long)frames;
    ldpc::channel_kernel<<<(unsigned)std::min<size_t>(frames, 4096), 256, 0, st>>>(mc, d->d_mc_llr);
    CUDA_TRY(cudaGetLastError());
    d->stats.kernel_launches++;
    CUDA_TRY(cudaMemcpyAsync(llr, d->d_mc_llr, frames * d->code.n * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return LDPC_OK;
}

int ldpc_hard_decision_batch(ldpc_decoder *d, const int32_t *values, size_t frames, int32_t *fail, uint32_t *bits)
{
    if (!d || !values || !fail) { ldpc::set_error("NULL decoder / values / fail"); return LDPC_ERR_ARG; }
    // zero iterations: the kernel initialises the slot from the values, evaluates the syndrome of their hard
    // decisions and stops; the iteration output then carries the syndrome flag (see KParams::max_iter)
    const int saved = d->cfg.max_iter;
    d->cfg.max_iter = 0;
    int rc = ldpc_decode_batch(d, values, frames, fail, bits, nullptr, nullptr);
    d->cfg.max_iter = saved;
    return rc;
}

int ldpc_decoder_device(const ldpc_decoder *d) { return d ? d->device : LDPC_ERR_ARG; }
const ldpc_code *ldpc_decoder_code(const ldpc_decoder *d) { return d ? &d->code : nullptr; }
int ldpc_decoder_max_iter(const ldpc_decoder *d) { return d ? d->cfg.max_iter : LDPC_ERR_ARG; }

int ldpc_decoder_sync(ldpc_decoder *d)
{
    if (!d) return LDPC_ERR_ARG;
    CUDA_TRY(cudaSetDevice(d->device));
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    return LDPC_OK;
}

int ldpc_decoder_get_stats(const ldpc_decoder *d, ldpc_decoder_stats *out)
{
    if (!d || !out) return LDPC_ERR_ARG;
    *out = d->stats;
    if (d->d_fb_total) {  // blocking read of the device-side counter
        unsigned long long total = 0;
        CUDA_TRY(cudaSetDevice(d->device));
        CUDA_TRY(cudaMemcpy(&total, d->d_fb_total, sizeof total, cudaMemcpyDeviceToHost));
        out->fallback_frames = total;
    }
    return LDPC_OK;
}

}  // extern "C"
