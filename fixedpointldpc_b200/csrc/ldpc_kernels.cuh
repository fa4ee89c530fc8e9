// ldpc_kernels.cuh -- sm_100a decode kernels of the fixed-point LDPC engine.
//
// Persistent CTAs (one, two or four per SM) keep W "word sets" resident in shared memory for the whole life
// of the frames decoded in them: E edge messages + n channel values, one 32-bit word each.
// A word holds one frame (Scalar32: exact int32 arithmetic) or two frames (Packed16: int16x2
// lanes, 13-bit magnitude guard).  Every lane of every word set is an independent frame slot
// with its own iteration counter; a slot that finishes is refilled from a global queue, so
// early termination never idles the other slots.
//
// Per trip of the main loop (== one flooding iteration of ArrayLDPC_Decoder.cpp:63-168 for
// every resident frame, rotated so that the syndrome needs no pass of its own):
//   variable phase  (ArrayLDPC_Decoder.cpp:121-156)  thread per variable, gather over the
//                   static edge-address table; the hard decision of the posterior goes
//                   into a spare bit of every outgoing message and of the channel word
//   check phase     (ArrayLDPC_Decoder.cpp:66-118, 296-333)  thread per check: forward chain
//                   in registers (XOR-ing the hard decisions on the way: the syndrome),
//                   backward chain + combine fused with the write-back
//   stop decision   early termination / max_iter / pre-check (:164-167, :443-450); stopping
//                   frames leave (results from the channel words) and the next ones move in
//
// Message words in shared memory
//   v2c (variable -> check): sign | hd | magnitude   (Scalar32: bit31 | bit30 | 30 bits,
//                                                     Packed16 lane: bit15 | bit14 | 14 bits)
//   c2v (check -> variable): two's complement of -c2v (negated so the variable phase only adds)
// The pairwise operator sxor (ArrayLDPC_Decoder.cpp:677-694) is evaluated on magnitudes only;
// the sign of every outgoing message is the XOR of the incoming sign bits, which is exact
// because a zero magnitude absorbs (sxor(0,y) == 0) and sgn(0) only matters when the result is 0.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace ldpc {

struct KParams {
    // code
    const uint8_t *cdeg;    // [m]
    const uint8_t *vdeg;    // [n]
    const uint16_t *vedge;  // [dv_max][n]  word index (slot*m + check) of edge j of variable v; checks are numbered
                            //              by descending degree in here, so the words of a word set are 0..E-1
    const uint16_t *eorig;  // [E]          the same word in the reference's numbering (EdgeRAM order of the v2c dump)
    int n, m, E, dc_max, dv_max;  // E <= dc_max*m words per word set (slot-major)
    // irregular codes: node handled by thread t in pass k of the variable / check phase = order[k*blockDim + t]
    // (0xffff: none).  A warp's 32 nodes in one pass have the same degree and the passes are dealt so that all
    // warps carry about the same work (ldpc_decoder.cu: build_order).
    const uint16_t *vorder, *corder;
    int vorder_k, corder_k;
    // schedule
    int W;         // word sets per CTA
    int max_iter;  // MAX_ITER; 0 = hard decision + syndrome only (iters output: 0 pass, 1 fail)
    uint32_t inv_m;  // floor(2^32 / m) + 1: i / m == umulhi(i, inv_m) for i < 2^16
    int precheck;  // decode_fixpoint's hardDecision pre-check
    int claim_ahead;  // every slot takes its next frame index one frame early (long queues only)
    // fed launches (host pipeline): frames [0, *avail) have arrived in `llr`; NULL: all of them
    const unsigned long long *avail;
    // fed launches: done_count[f / done_chunk] counts finished frames; the thread that completes a chunk sets the
    // host-mapped done_flag[chunk], and the host then copies that chunk's results out while the kernel runs on
    unsigned int *done_count;
    volatile unsigned int *done_flag;
    int done_chunk;
    // io
    const void *llr;  // [frames][n] int32 or int16
    int llr_bits;
    long long frames;
    int *iters;      // [frames]
    uint32_t *bits;  // [frames][nw32] or NULL
    int nw32;
    int *post;  // [frames][n] or NULL
    int *v2c;   // [frames][dc_max][m] or NULL
    unsigned long long *queue;  // next frame index
    // optional indirection (exact re-decode of the frames the packed kernel flagged): queue
    // position q decodes frame index[q], and the number of positions is *count
    const int *index;
    const int *count;
    // Monte-Carlo mode: the channel values are generated in the refill step instead of read from
    // memory and the finish step counts errors instead of (or besides) storing bits.
    int mc_mode;                   // 0 off, 1 Philox4x32-10 counter stream, 2 the reference's Lehmer stream
    unsigned long long mc_first;   // global index of frame 0 of this launch
    unsigned long long mc_seed;    // Philox key / Lehmer state before global frame 0
    double mc_gain;                // 2*snr   (LLR = 2*snr*(1 - 2c + N(0,sigma)), PerfTest.cpp:112)
    double mc_sigma;
    const uint32_t *mc_cw;         // [nw32] transmitted codeword bits or NULL (all-zero codeword)
    int mc_cw_stride;              // 0: one codeword for all frames; nw32: frame f sends mc_cw + f*stride
    const uint32_t *mc_info;       // [nw32] positions calculateBER counts (ArrayLDPC_Decoder.cpp:707-722)
    const int *mc_pin;             // shortening: LLR_fp[pin[i]] = pin_value (PerfTest.cpp:410-414)
    int mc_pin_count, mc_pin_value;
    const uint32_t *mc_pow;        // [n] a^(v+1) mod (2^31-1)  (mode 2)
    uint32_t mc_jump;              // a^n mod (2^31-1): one frame consumes n uniforms (mode 2)
    unsigned short *mc_frame_err;  // [frames] info-bit errors per frame (saturating) or NULL
    unsigned long long *mc_counters;  // frames, frame errors, bit errors, iteration sum
};

// prmt with sign-replicating selectors (0xbb99): 0xffff in every 16-bit lane whose sign bit is
// set.  __byte_perm() masks the replicate bit off the selector, so this has to be PTX.
__device__ __forceinline__ uint32_t lane_sign_mask(uint32_t x)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xbb99;" : "=r"(r) : "r"(x));
    return r;
}

__device__ __forceinline__ uint32_t or3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xfe;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// ------------------------------------------------------------------------------------------
// lane traits
// ------------------------------------------------------------------------------------------

// One frame per word, exact while every value stays below 2^30 in magnitude.
struct Scalar32 {
    static constexpr int LANES = 1;
    static constexpr uint32_t MAG = 0x3fffffffu, SIGN = 0x80000000u, HD = 0x40000000u;
    static constexpr uint32_t ALL = 0xffffffffu;

    // magnitude part of sxor: min + max(0,10-(sum&255)>>2) - max(0,10-(diff&255)>>2)
    //                       = min + min(10,(diff&255)>>2) - min(10,(sum&255)>>2)
    __device__ static __forceinline__ uint32_t g(uint32_t a, uint32_t b)
    {
        uint32_t mn = min(a, b);
        uint32_t s = a + b;
        uint32_t d = s - 2u * mn;
        uint32_t pk = __byte_perm(s, d, 0x5410);          // [d.lo16 | s.lo16]
        uint32_t q = (pk >> 2) & 0x003f003fu;             // ((x & 255) >> 2) per half
        uint32_t u = __vminu2(q, 0x000a000au);
        return (uint32_t)__dp2a_lo((int)u, 0x000001ff, (int)mn);  // mn - u(sum) + u(diff)
    }
    // word written by the check phase: -c2v, where c2v = +o if the sign bit of `inv` is set
    // (inv = ~(XOR of the other incoming words)) and -o otherwise
    __device__ static __forceinline__ uint32_t neg_c2v(uint32_t o, uint32_t inv)
    {
        uint32_t mneg = (uint32_t)((int)inv >> 31);  // all ones where c2v > 0 -> store -o
        return (o ^ mneg) - mneg;
    }
    __device__ static __forceinline__ uint32_t fail_bits(uint32_t acc) { return (acc >> 30) & 1u; }

    struct Acc { int v; };
    __device__ static __forceinline__ Acc acc_init(uint32_t llr) { return Acc{(int)llr}; }
    __device__ static __forceinline__ void acc_sub(Acc &a, uint32_t nc) { a.v -= (int)nc; }
    // posterior word + its hard decision in message position (bit = post <= 0, quirk Q3)
    template <int D> __device__ static __forceinline__ uint32_t post_word(const Acc &a, uint32_t &hd)
    {
        hd = (uint32_t)(a.v - 1) >> 1;  // sign of post-1, moved to bit 30 (the users mask with HD)
        return (uint32_t)a.v;
    }
    template <int D> __device__ static __forceinline__ uint32_t posterior(uint32_t lx, const uint32_t *x, uint32_t &hd)
    {
        Acc acc = acc_init(lx);
#pragma unroll
        for (int j = 0; j < D; ++j) acc_sub(acc, x[j]);
        return post_word<D>(acc, hd);
    }
    // sign | magnitude of post + (-c2v); the caller ORs the hard-decision bit in
    __device__ static __forceinline__ uint32_t v2c_signmag(uint32_t post, uint32_t nc)
    {
        int v = (int)post + (int)nc;
        return (uint32_t)abs(v) | ((uint32_t)v & SIGN);
    }
    __device__ static __forceinline__ bool guard_hit(uint32_t) { return false; }
    __device__ static __forceinline__ uint32_t guard_lanes(uint32_t) { return 0u; }
    __device__ static __forceinline__ int lane_value(uint32_t w, int) { return (int)w; }
    __device__ static __forceinline__ int v2c_value(uint32_t w, int)
    {
        int mag = (int)(w & MAG);
        return (w & SIGN) ? -mag : mag;
    }
    __device__ static __forceinline__ uint32_t hd_bit(uint32_t w, int) { return (w >> 30) & 1u; }
    // A channel word also carries the hard decision of the variable's last posterior in bit 30, which is a copy of
    // the sign bit as long as |LLR| < 2^30: llr_restore() gives the channel value back, llr_with_hd() stores the bit.
    __device__ static __forceinline__ uint32_t llr_restore(uint32_t x) { return (x & ~HD) | ((x >> 1) & HD); }
    __device__ static __forceinline__ uint32_t llr_with_hd(uint32_t x, uint32_t hd) { return (x & ~HD) | (hd & HD); }
    // lane content of a channel word / of a fresh frame's message word: v2c(0) = channel value
    // (ArrayLDPC_Decoder.cpp:45-61), carrying that value's hard decision
    __device__ static __forceinline__ uint32_t llr_lane(int val, bool &bad)
    {
        bad = false;
        return (uint32_t)val;
    }
    __device__ static __forceinline__ uint32_t init_lane(int val)
    {
        return (uint32_t)abs(val) | ((uint32_t)val & SIGN) | (val <= 0 ? HD : 0u);
    }
    // refill pass: the lane's hard-decision bit in a channel word, the stored form of a channel value, its range check
    __device__ static __forceinline__ uint32_t hd_mask(int) { return HD; }
    __device__ static __forceinline__ uint32_t llr_word(int val) { return (uint32_t)val; }
    __device__ static __forceinline__ uint32_t range_key(int) { return 0u; }
    __device__ static __forceinline__ bool range_bad(uint32_t) { return false; }
    __device__ static __forceinline__ void store_lane(uint32_t *word, int, uint32_t x) { *word = x; }
};

// Two frames per word (int16x2).  Exact as long as every |LLR|, |posterior| and |message| stays
// below 2^13; a lane that leaves that range is flagged and its frame is re-decoded by the
// Scalar32 kernel (there is no saturation in the reference).
struct Packed16 {
    static constexpr int LANES = 2;
    static constexpr uint32_t MAG = 0x1fff1fffu, SIGN = 0x80008000u, HD = 0x40004000u;
    static constexpr uint32_t GUARD = 0x60006000u;  // magnitude bits 13,14 of the un-flagged word
    static constexpr uint32_t ALL = 0xffffffffu;
    static constexpr int LIMIT = 1 << 13;

    __device__ static __forceinline__ uint32_t g(uint32_t a, uint32_t b)
    {
        uint32_t mn = __vminu2(a, b);
        uint32_t s = a + b;                      // lanes < 2^15: no carry across
        uint32_t d;                              // |a-b| = s - 2*mn per lane, no borrow; on the FMA pipe
        asm("mad.lo.u32 %0, %1, 0xfffffffe, %2;" : "=r"(d) : "r"(mn), "r"(s));
        uint32_t us = __vminu2(s & 0x00fc00fcu, 0x00280028u);  // 4*min(10,(s&255)>>2)
        uint32_t ud = __vminu2(d & 0x00fc00fcu, 0x00280028u);
        uint32_t x = 4u * mn + ud;               // 4*(mn + u(diff)), lanes < 2^16
        // x - us on the FMA pipe (the ALU pipe is the bound resource); every lane stays >= 0
        asm("mad.lo.u32 %0, %1, 0xffffffff, %0;" : "+r"(x) : "r"(us));
        return x >> 2;  // low two bits of every lane are zero (IMAD.HI here, 2 FMA slots for 1 ALU slot: 4 - 7 % slower on all codes)
    }
    __device__ static __forceinline__ uint32_t neg_c2v(uint32_t o, uint32_t inv)
    {
        uint32_t mneg = lane_sign_mask(inv);     // 0xffff in lanes where c2v > 0
        return __vadd2(o, mneg) ^ mneg;          // ~(o-1) == -o in those lanes
    }
    __device__ static __forceinline__ uint32_t fail_bits(uint32_t acc)
    {
        return ((acc >> 14) & 1u) | ((acc >> 29) & 2u);
    }

    struct Acc { int lo, hi; };
    __device__ static __forceinline__ Acc acc_init(uint32_t llr)
    {
        return Acc{__dp2a_lo((int)llr, 0x00000001, 0), __dp2a_lo((int)llr, 0x00000100, 0)};
    }
    __device__ static __forceinline__ void acc_sub(Acc &a, uint32_t nc)
    {
        a.lo = __dp2a_lo((int)nc, 0x000000ff, a.lo);  // -= (int16) low lane
        a.hi = __dp2a_lo((int)nc, 0x0000ff00, a.hi);  // -= (int16) high lane
    }
    // Posteriors are summed in 32 bits and clamped to +-CLAMP before packing: a clamped lane still
    // yields |post + (-c2v)| >= CLAMP - (2^13+9) > 2^13, so the message guard below catches it,
    // and CLAMP + 2^13 + 9 < 2^15 keeps the packed add from wrapping.
    static constexpr int CLAMP = 24000;
    static constexpr uint32_t CLAMP2 = 0x5dc05dc0u, NCLAMP2 = 0xa240a240u;  // +CLAMP / -CLAMP in both lanes
    template <int D> __device__ static __forceinline__ uint32_t post_word(const Acc &a, uint32_t &hd)
    {
        int lo = a.lo, hi = a.hi;
        // D <= 2: |post| <= LLR_LIMIT-1 + 2*(2^13+19) and |post + (-c2v)| stay below 2^15 (see LLR_LIMIT)
        uint32_t pw;
        if (D > 2) {
            // saturating pack (I2IP), then +-CLAMP on both lanes at once: three instructions for five
            asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(pw) : "r"(hi), "r"(lo));
            pw = __vmaxs2(__vmins2(pw, CLAMP2), NCLAMP2);
        } else {
            pw = __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410);
        }
        hd = __vadd2(pw, 0xffffffffu) >> 1;  // the users mask with HD  // sign of post-1 per lane, moved to bit 14
        return pw;
    }
    // posterior word of a degree-D variable from its channel word and the D stored words (-c2v)
    template <int D> __device__ static __forceinline__ uint32_t posterior(uint32_t lx, const uint32_t *x, uint32_t &hd)
    {
        // The 32-bit sums cost two IDP per word; words are first added in the lanes as far as a lane holds them
        // (|LLR| < LLR_LIMIT, |c2v| <= C2V_MAX = 2^13 + 19, see LLR_LIMIT).
        if (D <= 3) {
            // |LLR| + 3 * C2V_MAX < 2^15: the whole sum stays in the lanes
            static_assert(LLR_LIMIT - 1 + 3 * (LIMIT + 19) < (1 << 15), "degree-3 posterior leaves its lane");
            uint32_t s = x[0];
#pragma unroll
            for (int j = 1; j < D; ++j) s = __vadd2(s, x[j]);
            const uint32_t pw = __vadd2(__vadd2(lx, ~s), 0x00010001u);  // lx - s per lane
            hd = __vadd2(pw, 0xffffffffu) >> 1;  // the users mask with HD
            return pw;
        }
        Acc acc = acc_init(lx);
#pragma unroll
        for (int j = 0; j < D; j += 3) {  // three words at a time: 3 * C2V_MAX < 2^15
            uint32_t s = x[j];
            if (j + 1 < D) s = __vadd2(s, x[j + 1]);
            if (j + 2 < D) s = __vadd2(s, x[j + 2]);
            acc_sub(acc, s);
        }
        return post_word<D>(acc, hd);
    }
    __device__ static __forceinline__ uint32_t v2c_signmag(uint32_t post, uint32_t nc)
    {
        uint32_t v = __vadd2(post, nc);
        uint32_t m = lane_sign_mask(v);                  // 0xffff in negative lanes
        uint32_t t = __vadd2(v, m);                      // v-1 there
        return t ^ (m & 0x7fff7fffu);                    // sign | |v|; magnitude bits 13,14 feed the guard
    }
    __device__ static __forceinline__ bool guard_hit(uint32_t guard) { return (guard & GUARD) != 0u; }
    __device__ static __forceinline__ uint32_t guard_lanes(uint32_t guard)
    {
        return ((guard & 0x00006000u) ? 1u : 0u) | ((guard & 0x60000000u) ? 2u : 0u);
    }
    __device__ static __forceinline__ int lane_value(uint32_t w, int lane)
    {
        return (int)(int16_t)(lane ? (w >> 16) : (w & 0xffffu));
    }
    __device__ static __forceinline__ int v2c_value(uint32_t w, int lane)
    {
        uint32_t h = lane ? (w >> 16) : (w & 0xffffu);
        int mag = (int)(h & 0x3fffu);
        return (h & 0x8000u) ? -mag : mag;
    }
    __device__ static __forceinline__ uint32_t hd_bit(uint32_t w, int lane) { return (w >> (lane ? 30 : 14)) & 1u; }
    // channel lanes stay below LLR_LIMIT < 2^13 in magnitude, so bit 14 of a lane is a copy of its sign bit and
    // can carry the hard decision of the variable's last posterior (see Scalar32)
    __device__ static __forceinline__ uint32_t llr_restore(uint32_t x) { return (x & ~HD) | ((x >> 1) & HD); }
    __device__ static __forceinline__ uint32_t llr_with_hd(uint32_t x, uint32_t hd) { return (x & ~HD) | (hd & HD); }
    // Channel values are admitted below LLR_LIMIT, slightly under 2^13, so that the unclamped posterior of a
    // degree <= 2 variable cannot wrap its lane: LLR_LIMIT-1 + 2*C2V_MAX + C2V_MAX < 2^15, where
    // C2V_MAX = 2^13-1 + 20 bounds every check output (each sxor adds at most 10 to the smaller magnitude).
    static constexpr int LLR_LIMIT = 8100;
    __device__ static __forceinline__ uint32_t llr_lane(int val, bool &bad)
    {
        bad = (uint32_t)(val + LLR_LIMIT - 1) >= 2u * LLR_LIMIT - 1u;  // |val| >= LLR_LIMIT
        return (uint32_t)val & 0xffffu;
    }
    // |val| >= LLR_LIMIT  <=>  (uint32_t)(val + LLR_LIMIT - 1) >= 2 * LLR_LIMIT - 1: the refill keeps the maximum of the keys
    __device__ static __forceinline__ uint32_t hd_mask(int lane) { return lane ? 0x40000000u : 0x00004000u; }
    __device__ static __forceinline__ uint32_t llr_word(int val) { return (uint32_t)val & 0xffffu; }
    __device__ static __forceinline__ uint32_t range_key(int val) { return (uint32_t)(val + LLR_LIMIT - 1); }
    __device__ static __forceinline__ bool range_bad(uint32_t key) { return key >= 2u * LLR_LIMIT - 1u; }
    __device__ static __forceinline__ uint32_t init_lane(int val)
    {
        // a value outside the guard range was flagged when it was loaded; only its low 14 magnitude bits are kept here
        return ((uint32_t)abs(val) & 0x3fffu) | (val < 0 ? 0x8000u : 0u) | (val <= 0 ? 0x4000u : 0u);
    }
    // a lane is a 16-bit half of the word: no read-modify-write
    __device__ static __forceinline__ void store_lane(uint32_t *word, int lane, uint32_t x)
    {
        reinterpret_cast<uint16_t *>(word)[lane] = (uint16_t)x;
    }
};

// ------------------------------------------------------------------------------------------
// channel: BPSK + AWGN + the reference's quantiser LLR_fp = int(2*snr*(1 - 2c + N(0,sigma)) * 2^4)
// (PerfTest.cpp:108-120, 164-170, 287-297, 494-504), truncation toward zero, no clipping (Q11)
// ------------------------------------------------------------------------------------------
constexpr uint32_t LEHMER_M = 2147483647u;  // rngs.cpp:40

__host__ __device__ __forceinline__ uint32_t lehmer_mul(uint32_t a, uint32_t b)
{
    unsigned long long p = (unsigned long long)a * b;
    p = (p & LEHMER_M) + (p >> 31);
    p = (p & LEHMER_M) + (p >> 31);
    return (uint32_t)(p >= LEHMER_M ? p - LEHMER_M : p);
}

__host__ __device__ __forceinline__ uint32_t lehmer_pow(uint32_t base, unsigned long long e)
{
    uint32_t r = 1u;
    while (e) {
        if (e & 1ull) r = lehmer_mul(r, base);
        base = lehmer_mul(base, base);
        e >>= 1;
    }
    return r;
}

// rvgs.cpp:152-181 (Odeh-Evans inverse normal) on the uniform rngs.cpp:52-69 returns for `state`.
// Written with explicit round-to-nearest multiplies/adds so nvcc cannot contract them into FMAs:
// the reference build evaluates every product and sum separately.
__device__ __forceinline__ double lehmer_normal(uint32_t state)
{
    const double u = (double)state / 2147483647.0;
    const double t = u < 0.5 ? sqrt(__dmul_rn(-2.0, log(u))) : sqrt(__dmul_rn(-2.0, log(__dsub_rn(1.0, u))));
    double pn = __dadd_rn(0.204231210245e-1, __dmul_rn(t, 0.453642210148e-4));
    pn = __dadd_rn(0.342242088547, __dmul_rn(t, pn));
    pn = __dadd_rn(1.0, __dmul_rn(t, pn));
    pn = __dadd_rn(0.322232431088, __dmul_rn(t, pn));
    double qn = __dadd_rn(0.103537752850, __dmul_rn(t, 0.385607006340e-2));
    qn = __dadd_rn(0.531103462366, __dmul_rn(t, qn));
    qn = __dadd_rn(0.588581570495, __dmul_rn(t, qn));
    qn = __dadd_rn(0.099348462606, __dmul_rn(t, qn));
    const double ratio = __ddiv_rn(pn, qn);
    return u < 0.5 ? __dsub_rn(ratio, t) : __dsub_rn(t, ratio);
}

__device__ __forceinline__ int quantise_llr(const KParams &p, double z, uint32_t bit)
{
    // 2*snr*(1 - 2*c + (0 + sigma*z)) * 16, in the reference's evaluation order
    const double noise = __dadd_rn(0.0, __dmul_rn(p.mc_sigma, z));
    const double x = __dadd_rn(bit ? -1.0 : 1.0, noise);
    return (int)__dmul_rn(__dmul_rn(p.mc_gain, x), 16.0);
}

struct Philox4 { uint32_t x[4]; };
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox4{{c0, c1, c2, c3}};
}

// four standard normals for variables 4*q..4*q+3 of global frame g: two Box-Muller pairs
__device__ __forceinline__ void philox_normals(const KParams &p, unsigned long long g, uint32_t q, float z[4])
{
    const Philox4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), q, 0x4c445043u, (uint32_t)p.mc_seed,
                                    (uint32_t)(p.mc_seed >> 32));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float u1 = ((float)(r.x[2 * h] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1), 24 bits
        const float u2 = ((float)(r.x[2 * h + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float rad = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        z[2 * h] = rad * cs;
        z[2 * h + 1] = rad * sn;
    }
}

__device__ __forceinline__ uint32_t cw_bit(const KParams &p, long long f, int v)
{
    return p.mc_cw ? (p.mc_cw[(size_t)f * p.mc_cw_stride + (v >> 5)] >> (v & 31)) & 1u : 0u;
}

// ------------------------------------------------------------------------------------------
// control block in shared memory (after the message words, channel words and hard-decision bits)
// ------------------------------------------------------------------------------------------
constexpr int MAX_W = 16;         // word sets per CTA
constexpr bool PREFETCH_VEDGE = true;   // software-prefetch the next variable's edge addresses (costs registers)
constexpr int MAX_SLOTS = 2 * MAX_W;

struct Ctrl {
    // Votes of one trip, read by that trip's stop decision.  Every thread takes the decision for itself, without a
    // barrier, so these words rotate through three buffers (trip % 3): the buffer a trip writes was cleared two
    // decisions ago and is next touched two barriers later.
    uint32_t fail[3][MAX_W];   // per word set: lanes with at least one unsatisfied check (check phase)
    uint32_t gflag[3][MAX_W];  // per word set: lanes that left the guard range (variable phase / load; Packed16)
    int fid[MAX_SLOTS];     // frame index decoded in the slot, -1 = idle
    int next[MAX_SLOTS];    // frame the slot decodes after this one (taken from the queue one frame early), -1 = none
    int newfid[MAX_SLOTS];  // frame moving into the slot during a refill
    unsigned int start[MAX_SLOTS];  // trip in which the slot's frame moved in: iterations completed = trip - start
    uint32_t lehmer[MAX_SLOTS];  // Lehmer state before the slot's frame (MC mode 2)
    unsigned int errs[MAX_SLOTS];  // info-bit errors of the frames being finished (MC mode)
};

// ------------------------------------------------------------------------------------------
// check phase (ArrayLDPC_Decoder.cpp:66-118 + the syndrome of :296-333): NI independent check nodes of exact
// degree D per thread, interleaved instruction by instruction so every thread carries NI dependency chains
// (the chains are serial by construction -- sxor is not associative -- and one chain per thread leaves the
// ALU pipe idle).  The NI nodes are the same check c in NI consecutive word sets: e[i] = e0 + i*wstride.
//
// The forward chain F[k] = g(F[k-1], m[k]) stays in registers; on the way it XORs the incoming words: bit 30/14
// of the XOR is the parity of the senders' hard decisions -- this check's syndrome bit for the state the last
// variable phase left (returned in fb) -- and bit 31/15 the parity of the signs, so bit 31/15 of
// (~xor ^ word_k) is the inverted sign of the outgoing message of slot k.  The backward chain and the combine
// c2v[k] = g(F[k-1], B[k+1]) walk back down, re-reading each word once and overwriting it in place.
// ------------------------------------------------------------------------------------------
template <class T, int D, int NI>
__device__ __forceinline__ void check_nodes(uint32_t *e0, int m, int wstride, uint32_t (&fb)[NI])
{
    uint32_t fwd[NI][D - 1], w0[NI], nacc[NI], bwd[NI];
    uint32_t *e[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        e[i] = e0 + i * wstride;
        w0[i] = e[i][0];
        nacc[i] = w0[i];
        fwd[i][0] = w0[i] & T::MAG;
    }
#pragma unroll
    for (int k = 1; k < D - 1; ++k)
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const uint32_t w = e[i][k * m];
            nacc[i] ^= w;
            fwd[i][k] = T::g(fwd[i][k - 1], w & T::MAG);
        }
    // The backward pass reads every word a second time (volatile: otherwise the compiler keeps the D words of the
    // forward pass alive and, out of registers, parks them in local memory -- which is L2 here).
#pragma unroll
    for (int k = D - 1; k >= 1; --k)
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const uint32_t w = k == D - 1 ? e[i][k * m] : *reinterpret_cast<volatile uint32_t *>(&e[i][k * m]);
            const uint32_t mag = w & T::MAG;
            uint32_t o;
            if (k == D - 1) {
                nacc[i] ^= w;
                fb[i] = T::fail_bits(nacc[i]);
                nacc[i] = ~nacc[i];
                o = fwd[i][D - 2];  // c2v[d-1] = Forward[d-2]
                bwd[i] = mag;
            } else {
                o = T::g(fwd[i][k - 1], bwd[i]);  // c2v[k] = sxor(Forward[k-1], Backward[k+1])
                bwd[i] = T::g(bwd[i], mag);       // Backward[k]
            }
            e[i][k * m] = T::neg_c2v(o, nacc[i] ^ w);
        }
#pragma unroll
    for (int i = 0; i < NI; ++i) e[i][0] = T::neg_c2v(bwd[i], nacc[i] ^ w0[i]);  // c2v[0] = Backward[1]
}

// one check node of run-time degree 2 <= d <= DC (predicated; only for degrees without an exact body)
template <class T, int DC>
__device__ __forceinline__ uint32_t check_node_any(uint32_t *e, int m, int d)
{
    uint32_t fwd[DC - 1];
    const uint32_t w0 = e[0];
    uint32_t nacc = w0;
    fwd[0] = w0 & T::MAG;
    uint32_t last = fwd[0];  // ends as Forward[d-2] without a runtime-indexed read of fwd[]
#pragma unroll
    for (int k = 1; k < DC - 1; ++k) {
        if (k < d - 1) {
            const uint32_t w = e[k * m];
            nacc ^= w;
            fwd[k] = T::g(fwd[k - 1], w & T::MAG);
            last = fwd[k];
        }
    }
    nacc ^= e[(d - 1) * m];
    const uint32_t fb = T::fail_bits(nacc);
    nacc = ~nacc;
    uint32_t bwd = 0;
#pragma unroll
    for (int k = DC - 1; k >= 1; --k) {
        if (k < d) {
            uint32_t w = e[k * m];
            uint32_t mag = w & T::MAG;
            uint32_t o;
            if (k == d - 1) {
                o = last;
                bwd = mag;
            } else {
                o = T::g(fwd[k - 1], bwd);
                bwd = T::g(bwd, mag);
            }
            e[k * m] = T::neg_c2v(o, nacc ^ w);
        }
    }
    e[0] = T::neg_c2v(bwd, nacc ^ w0);
    return fb;
}

// thread per (word-set group, check); the syndrome bits are OR-ed into fail[word set], one shared atomic per warp
// CMASK: check degrees that get an exact body (bit d); a named irregular code lists its own, which keeps the hot
// code small (the 802.11 code has degrees 7 and 8 only -- fifteen exact bodies are 5 000 instructions)
// CS: (regular codes, NI == 1) checks of one word set per thread, interleaved like the NI word sets are: check c,
// c + m/CS, ... -- instruction-level parallelism for a CTA that holds a single word set
template <class T, int DC, bool REG, int NI, unsigned CMASK, int CS>
__device__ __forceinline__ void check_phase(const KParams &p, uint32_t *fail, uint32_t *edge, const uint8_t *cdeg_s,
                                            int items, int m, int E, int W, int nthreads)
{
    const int tid = threadIdx.x, lane_id = tid & 31;
    if (!REG) {
        // irregular code: the checks of a warp's pass have one degree (p.corder), one group of NI word sets at a time
        for (int wg = 0; wg < W / NI; ++wg) {
            uint32_t mine[NI];
#pragma unroll
            for (int j = 0; j < NI; ++j) mine[j] = 0u;
            uint32_t cnext = p.corder[tid];
            for (int k = 0; k < p.corder_k; ++k) {
                const int c = (int)cnext;
                if (k + 1 < p.corder_k) cnext = p.corder[(k + 1) * nthreads + tid];
                if (c == 0xffff) continue;
                uint32_t *e0 = edge + (size_t)(wg * NI) * E + c;
                const int d = cdeg_s[c];
                uint32_t fb[NI];
#pragma unroll
                for (int j = 0; j < NI; ++j) fb[j] = 0u;
                bool done = false;
                if (DC <= 16) {
                    // exact-degree bodies: no per-edge predicates or branches inside
                    switch (d) {
#define LDPC_CCASE(D) case D: if (D <= DC && ((CMASK >> D) & 1u)) { check_nodes<T, (D <= DC ? D : 2), NI>(e0, m, E, fb); done = true; } break;
                        LDPC_CCASE(2) LDPC_CCASE(3) LDPC_CCASE(4) LDPC_CCASE(5) LDPC_CCASE(6) LDPC_CCASE(7) LDPC_CCASE(8)
                        LDPC_CCASE(9) LDPC_CCASE(10) LDPC_CCASE(11) LDPC_CCASE(12) LDPC_CCASE(13) LDPC_CCASE(14)
                        LDPC_CCASE(15) LDPC_CCASE(16)
#undef LDPC_CCASE
                    default: break;
                    }
                }
                if (!done)  // the loader rejects checks of degree < 2
                    for (int j = 0; j < NI; ++j) fb[j] = check_node_any<T, DC>(e0 + (size_t)j * E, m, d);
#pragma unroll
                for (int j = 0; j < NI; ++j) mine[j] |= fb[j];
            }
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                const uint32_t r = __reduce_or_sync(0xffffffffu, mine[j]);
                if (lane_id == 0 && r) atomicOr(&fail[wg * NI + j], r);
            }
        }
        return;
    }
    // When the CTA holds a single word-set group (W == NI) every item votes for the same NI word sets: the
    // thread ORs its items' verdicts in registers and the warp reduces once, after the loop.
    const bool one_group = (W == NI);
    uint32_t mine[NI];
#pragma unroll
    for (int j = 0; j < NI; ++j) mine[j] = 0u;
    for (int i0 = tid - lane_id; i0 < items; i0 += nthreads) {
        const int i = i0 + lane_id;
        const bool valid = i < items;
        int wg = 0;
        uint32_t fb[NI];
#pragma unroll
        for (int j = 0; j < NI; ++j) fb[j] = 0u;
        if (valid && CS == 1) {
            wg = (int)__umulhi((uint32_t)i, p.inv_m);
            const int c = i - wg * m;
            check_nodes<T, DC, NI>(edge + (size_t)(wg * NI) * E + c, m, E, fb);
        }
        if (valid && CS > 1) {
            static_assert(CS == 1 || NI == 1, "check split needs one word set per thread");
            const int mh = m / CS;
            wg = i / mh;
            const int c = i - wg * mh;
            uint32_t fs[CS];
            check_nodes<T, DC, CS>(edge + (size_t)wg * E + c, m, mh, fs);
#pragma unroll
            for (int j = 0; j < CS; ++j) fb[0] |= fs[j];
        }
        if (one_group) {
#pragma unroll
            for (int j = 0; j < NI; ++j) mine[j] |= fb[j];
            continue;
        }
        const int g0 = __shfl_sync(0xffffffffu, wg, 0);
        // a warp covers at most two word-set groups when m >= 32; anything else goes the slow way
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const uint32_t r0 = __reduce_or_sync(0xffffffffu, (valid && wg == g0) ? fb[j] : 0u);
            const uint32_t r1 = __reduce_or_sync(0xffffffffu, (valid && wg == g0 + 1) ? fb[j] : 0u);
            if (lane_id == 0) {
                if (r0) atomicOr(&fail[g0 * NI + j], r0);
                if (r1) atomicOr(&fail[(g0 + 1) * NI + j], r1);
            }
            if (valid && wg > g0 + 1 && fb[j]) atomicOr(&fail[wg * NI + j], fb[j]);
        }
    }
    if (one_group && tid - lane_id < items) {
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const uint32_t r = __reduce_or_sync(0xffffffffu, mine[j]);
            if (lane_id == 0 && r) atomicOr(&fail[j], r);
        }
    }
}

// ------------------------------------------------------------------------------------------
// variable phase for one variable node of exact degree D (ArrayLDPC_Decoder.cpp:121-156:
// post = LLR + sum c2v, v2c_j = post - c2v_j), NW word sets at a time for instruction-level
// parallelism.  The hard decision of the posterior rides in a spare bit of every outgoing message (the check
// phase XORs them into the syndrome) and is also kept in a spare bit of the variable's channel word, which
// is what a stopping frame's result is read from after the check phase has overwritten the messages.
// PARITY: also store the posteriors and the messages of every iteration (parity-mode outputs; the last
// iteration's survive).
// ------------------------------------------------------------------------------------------

template <class T, int D, int NW, bool PARITY, bool ACC>
__device__ __forceinline__ void variable_words(const KParams &p, Ctrl *ctrl, uint32_t *gflag, uint32_t *gacc, char *base,
                                               uint32_t stride, const uint32_t *off, uint32_t *llr, int v, int w, int n)
{
    // ACC: gacc[] is the thread's running OR of the message words of word sets w, w+1, tested once per phase
    uint32_t x[NW][D > 0 ? D : 1], pw[NW], hd[NW], guard[NW], lx[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        lx[i] = T::llr_restore(llr[(w + i) * n + v]);
#pragma unroll
        for (int j = 0; j < D; ++j) x[i][j] = *reinterpret_cast<uint32_t *>(base + i * stride + off[j]);
        pw[i] = T::template posterior<D>(lx[i], x[i], hd[i]);
        guard[i] = ACC ? gacc[i] : 0u;
    }
#pragma unroll
    for (int j = 0; j < D; j += 2)
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            const uint32_t a = T::v2c_signmag(pw[i], x[i][j]);
            *reinterpret_cast<uint32_t *>(base + i * stride + off[j]) = a | (hd[i] & T::HD);
            x[i][j] = a;
            if (j + 1 < D) {
                const uint32_t b = T::v2c_signmag(pw[i], x[i][j + 1]);
                *reinterpret_cast<uint32_t *>(base + i * stride + off[j + 1]) = b | (hd[i] & T::HD);
                x[i][j + 1] = b;
                guard[i] = or3(guard[i], a, b);
            } else {
                guard[i] |= a;
            }
        }
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        if (ACC) gacc[i] = guard[i];
        else if (T::guard_hit(guard[i])) atomicOr(&gflag[w + i], T::guard_lanes(guard[i]));
        llr[(w + i) * n + v] = T::llr_with_hd(lx[i], hd[i]);
        if (PARITY) {
#pragma unroll
            for (int lane = 0; lane < T::LANES; ++lane) {
                const int f = ctrl->fid[(w + i) * T::LANES + lane];
                if (f < 0) continue;
                if (p.post) p.post[(size_t)f * n + v] = T::lane_value(pw[i], lane);
                if (p.v2c) {
                    int *out = p.v2c + (size_t)f * p.dc_max * p.m;  // EdgeRAM order: [slot][check]
#pragma unroll
                    for (int j = 0; j < D; ++j) out[p.eorig[off[j] >> 2]] = T::v2c_value(x[i][j], lane);
                }
            }
        }
    }
}

template <class T, int D, bool PARITY>
__device__ __forceinline__ void variable_node(const KParams &p, Ctrl *ctrl, uint32_t *gflag, uint32_t *edge, uint32_t *llr,
                                              int v, int W, int n, int E, const uint32_t *off)
{
    char *base = reinterpret_cast<char *>(edge);
    const uint32_t stride = (uint32_t)E * 4u;
    int w = 0;
    if (D <= 12 && W == 3) {  // all three word sets of the CTA at once
        variable_words<T, D, 3, PARITY, false>(p, ctrl, gflag, nullptr, base, stride, off, llr, v, 0, n);
        return;
    }
    for (; w + 1 < W; w += 2, base += 2 * stride)
        variable_words<T, D, 2, PARITY, false>(p, ctrl, gflag, nullptr, base, stride, off, llr, v, w, n);
    if (w < W) variable_words<T, D, 1, PARITY, false>(p, ctrl, gflag, nullptr, base, stride, off, llr, v, w, n);
}

// the same with the guard bits of the first two word sets (all there are for the named codes) OR-ed into the
// thread's accumulators instead of tested per variable; pays off for the short regular variable nodes
template <class T, int D, bool PARITY>
__device__ __forceinline__ void variable_node_acc(const KParams &p, Ctrl *ctrl, uint32_t *gflag, uint32_t (&gacc)[2], uint32_t *edge,
                                                  uint32_t *llr, int v, int W, int n, int E, const uint32_t *off)
{
    char *base = reinterpret_cast<char *>(edge);
    const uint32_t stride = (uint32_t)E * 4u;
    if (W == 1) {
        variable_words<T, D, 1, PARITY, true>(p, ctrl, gflag, gacc, base, stride, off, llr, v, 0, n);
        return;
    }
    variable_words<T, D, 2, PARITY, true>(p, ctrl, gflag, gacc, base, stride, off, llr, v, 0, n);
    int w = 2;
    base += 2 * stride;
    for (; w + 1 < W; w += 2, base += 2 * stride)
        variable_words<T, D, 2, PARITY, false>(p, ctrl, gflag, nullptr, base, stride, off, llr, v, w, n);
    if (w < W) variable_words<T, D, 1, PARITY, false>(p, ctrl, gflag, nullptr, base, stride, off, llr, v, w, n);
}

template <class T, int D, bool PARITY>
__device__ __forceinline__ void variable_node(const KParams &p, Ctrl *ctrl, uint32_t *gflag, uint32_t *edge, uint32_t *llr,
                                              int v, int W, int n, int E)
{
    uint32_t off[D > 0 ? D : 1];  // byte offsets of the D edge words inside a word set
#pragma unroll
    for (int j = 0; j < D; ++j) off[j] = (uint32_t)p.vedge[j * n + v] * 4u;
    variable_node<T, D, PARITY>(p, ctrl, gflag, edge, llr, v, W, n, E, off);
}

// any degree (slow path for degrees without an exact instantiation): two passes over the words
template <class T>
__device__ __forceinline__ void variable_node_any(const KParams &p, Ctrl *ctrl, uint32_t *gflag, uint32_t *edge, uint32_t *llr,
                                                  int v, int W, int dv, int n, int E)
{
    for (int w = 0; w < W; ++w) {
        uint32_t *ew = edge + (size_t)w * E;
        const uint32_t lx = T::llr_restore(llr[(size_t)w * n + v]);
        typename T::Acc acc = T::acc_init(lx);
        for (int j = 0; j < dv; ++j) T::acc_sub(acc, ew[p.vedge[(size_t)j * n + v]]);
        uint32_t hd, guard = 0;
        const uint32_t pw = T::template post_word<64>(acc, hd);
        for (int j = 0; j < dv; ++j) {
            uint32_t *q = &ew[p.vedge[(size_t)j * n + v]];
            const uint32_t a = T::v2c_signmag(pw, *q);
            guard |= a;
            *q = a | (hd & T::HD);
        }
        if (T::guard_hit(guard)) atomicOr(&gflag[w], T::guard_lanes(guard));
        llr[(size_t)w * n + v] = T::llr_with_hd(lx, hd);
        for (int lane = 0; lane < T::LANES; ++lane) {
            const int f = ctrl->fid[w * T::LANES + lane];
            if (f < 0) continue;
            if (p.post) p.post[(size_t)f * n + v] = T::lane_value(pw, lane);
            if (p.v2c)
                for (int j = 0; j < dv; ++j) {
                    const uint32_t a = p.vedge[(size_t)j * n + v];
                    p.v2c[(size_t)f * p.dc_max * p.m + p.eorig[a]] = T::v2c_value(ew[a], lane);
                }
        }
    }
}

// REGV: every variable has degree DV -- the edge addresses of the thread's next variable are fetched while the
// current one is processed (the table lives in global memory / L2)
// VMASK: variable degrees that get an exact body (bit d), see CMASK
template <class T, int DV, bool PARITY, bool REGV, unsigned VMASK>
__device__ __forceinline__ void variable_phase(const KParams &p, Ctrl *ctrl, uint32_t *gflag, uint32_t *edge, uint32_t *llr, int W,
                                               int n, int E, const uint8_t *vdeg, int nthreads)
{
    if (REGV) {
        uint32_t gacc[2] = {0u, 0u};  // guard bits of the thread's variables in word sets 0 and 1
        int v = threadIdx.x;
        uint32_t next[DV];
#pragma unroll
        for (int j = 0; j < DV; ++j) next[j] = v < n ? (uint32_t)p.vedge[j * n + v] * 4u : 0u;
        for (; v < n; v += nthreads) {
            uint32_t off[DV];
#pragma unroll
            for (int j = 0; j < DV; ++j) off[j] = next[j];
            const int vn = v + nthreads;
#pragma unroll
            for (int j = 0; j < DV; ++j) next[j] = vn < n ? (uint32_t)p.vedge[j * n + vn] * 4u : 0u;
            variable_node_acc<T, DV, PARITY>(p, ctrl, gflag, gacc, edge, llr, v, W, n, E, off);
        }
        if (T::guard_hit(gacc[0])) atomicOr(&gflag[0], T::guard_lanes(gacc[0]));
        if (T::guard_hit(gacc[1])) atomicOr(&gflag[1], T::guard_lanes(gacc[1]));
    } else {
        uint32_t vnext = p.vorder[threadIdx.x];
        for (int k = 0; k < p.vorder_k; ++k) {
            const int v = (int)vnext;
            if (k + 1 < p.vorder_k) vnext = p.vorder[(k + 1) * nthreads + threadIdx.x];
            if (v == 0xffff) continue;
            const int dv = vdeg[v];
            bool done = false;
            if (DV <= 12) {
                // exact-degree bodies: no per-edge predicates or branches inside
                switch (dv) {
#define LDPC_VCASE(D) case D: if (D <= DV && ((VMASK >> D) & 1u)) { variable_node<T, (D <= DV ? D : 1), PARITY>(p, ctrl, gflag, edge, llr, v, W, n, E); done = true; } break;
                    LDPC_VCASE(0) LDPC_VCASE(1) LDPC_VCASE(2) LDPC_VCASE(3) LDPC_VCASE(4) LDPC_VCASE(5) LDPC_VCASE(6)
                    LDPC_VCASE(7) LDPC_VCASE(8) LDPC_VCASE(9) LDPC_VCASE(10) LDPC_VCASE(11) LDPC_VCASE(12)
#undef LDPC_VCASE
                default: break;
                }
            } else if (dv == DV) {
                variable_node<T, DV, PARITY>(p, ctrl, gflag, edge, llr, v, W, n, E);
                done = true;
            }
            if (!done) variable_node_any<T>(p, ctrl, gflag, edge, llr, v, W, dv, n, E);
        }
    }
}

// Forward square array codes in closed form (class ROM, ArrayLDPCMacro.h:42-82; the variable-phase address of
// decode_fixpoint, ArrayLDPC_Decoder.cpp:583-589): variable v = b*P + j of column group b meets, in row group a,
// check a*P + ((j - a*b) mod P), and sits in that check's slot b.  The word index b*m + a*P + t_a follows from
// t_0 = j, t_(a+1) = t_a - b (mod P): two adds and an unsigned min per edge instead of a dependent table load from
// L2 in a phase that is latency-bound (the host verifies the code against this formula before choosing the kernel).
template <class T, int DV, int P, bool PARITY>
__device__ __forceinline__ void variable_phase_array(const KParams &p, Ctrl *ctrl, uint32_t *gflag, uint32_t *edge, uint32_t *llr,
                                                     int W, int n, int m, int E, int nthreads)
{
    uint32_t gacc[2] = {0u, 0u};
    constexpr uint32_t INV_P = (uint32_t)((1ull << 32) / P) + 1u;  // v / P == umulhi(v, INV_P) for v < 2^16
    for (int v = threadIdx.x; v < n; v += nthreads) {
        const uint32_t b = __umulhi((uint32_t)v, INV_P);
        uint32_t u = ((uint32_t)v - b * P) * 4u;       // 4 * t_a
        const uint32_t row = b * (uint32_t)m * 4u;      // byte offset of slot b
        const uint32_t step = b * 4u;
        uint32_t off[DV];
#pragma unroll
        for (int a = 0; a < DV; ++a) {
            off[a] = row + (uint32_t)(a * P * 4) + u;
            const uint32_t x = u - step;                // wraps to a huge value when t_a < b
            u = min(x, x + (uint32_t)(P * 4));
        }
        variable_node_acc<T, DV, PARITY>(p, ctrl, gflag, gacc, edge, llr, v, W, n, E, off);
    }
    if (T::guard_hit(gacc[0])) atomicOr(&gflag[0], T::guard_lanes(gacc[0]));
    if (W > 1 && T::guard_hit(gacc[1])) atomicOr(&gflag[1], T::guard_lanes(gacc[1]));
}

// parity-mode variant (posteriors and messages written through every iteration): kept out of line so that its
// register needs do not shape the allocation of the throughput path
template <class T, int DV, bool REGV, int ARRP, unsigned VMASK>
__device__ __noinline__ void variable_phase_parity(const KParams &p, Ctrl *ctrl, uint32_t *gflag, uint32_t *edge, uint32_t *llr, int W,
                                                   int n, int m, int E, const uint8_t *vdeg, int nthreads)
{
    if (ARRP) variable_phase_array<T, DV, (ARRP ? ARRP : 1), true>(p, ctrl, gflag, edge, llr, W, n, m, E, nthreads);
    else variable_phase<T, DV, true, REGV, VMASK>(p, ctrl, gflag, edge, llr, W, n, E, vdeg, nthreads);
}

// ------------------------------------------------------------------------------------------
// results of stopping frames, refill
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void *ptr)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}

// frame index behind queue position q (-1 past the end)
__device__ __forceinline__ int queue_frame(const KParams &p, unsigned long long q, long long frames)
{
    return q < (unsigned long long)frames ? (p.index ? p.index[q] : (int)q) : -1;
}

// fed launches: frames [0, result) have arrived in global memory (the host's copy-in stream moves the mark)
__device__ __forceinline__ long long frames_arrived(const KParams &p, long long frames)
{
    if (!p.avail) return frames;
    return (long long)*reinterpret_cast<const volatile unsigned long long *>(p.avail);
}

// hard decisions of 32 consecutive variables of the frame `fo` that leaves slot s: packed bits and / or the
// error count of Monte-Carlo mode (calculateBER, ArrayLDPC_Decoder.cpp:707-722, on packed words)
__device__ __forceinline__ void emit_word(const KParams &p, Ctrl *ctrl, int s, int fo, int v, int n, uint32_t bit)
{
    const uint32_t word = __ballot_sync(0xffffffffu, bit);
    if ((threadIdx.x & 31) == 0 && v < n) {
        if (p.bits) p.bits[(size_t)fo * p.nw32 + (v >> 5)] = word;
        if (p.mc_mode != 0) {
            const uint32_t sent = p.mc_cw ? p.mc_cw[(size_t)fo * p.mc_cw_stride + (v >> 5)] : 0u;
            const uint32_t mask = p.mc_info ? p.mc_info[v >> 5] : 0xffffffffu;
            const unsigned int e = __popc((word ^ sent) & mask);
            if (e) atomicAdd(&ctrl->errs[s], e);
        }
    }
}

// One slot of a decode-from-memory launch: the decoded bits of the frame that leaves (the hard decisions its last
// variable phase left in the channel words, 32 consecutive variables per ballot) and the channel values of the
// frame that moves in -- same thread, same word, so no barrier in between; four loads in flight per thread.
// `bits_out` / `src`: rows of the two frames (NULL: nothing to write / no frame).  Pointers advance by the CTA
// size instead of being rebuilt per element.  Returns true if a value of this thread left the packed range.
// BITS / LOAD select the two halves at compile time, the hard-decision bit of the lane is one mask for the whole pass,
// and the range check of the packed kernel is a running maximum tested once (T::range_key / T::range_bad) -- the pass
// is paid in issue slots (profiles/r02/launch_shape_sweep.txt).
template <class T, class SRC, bool BITS, bool LOAD>
__device__ __forceinline__ bool swap_frame_pass(uint32_t *lw, int lane, uint32_t *bits_out, const SRC *src, int n, int nthreads)
{
    constexpr int UNR = 4;
    const int tid = threadIdx.x;
    const bool leader = (tid & 31) == 0;
    const uint32_t hd = T::hd_mask(lane);
    uint32_t *word = lw + tid;
    uint32_t *bout = bits_out + (tid >> 5);
    const SRC *in = src + tid;
    uint32_t key = 0u;
    for (int v0 = tid; v0 - tid < n; v0 += UNR * nthreads) {
        int val[UNR];
        if (LOAD) {
#pragma unroll
            for (int u = 0; u < UNR; ++u) val[u] = v0 + u * nthreads < n ? (int)in[u * nthreads] : 0;
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int v = v0 + u * nthreads;
            if (v - tid < n) {  // the same for every thread of the CTA
                const bool mine = v < n;
                if (BITS) {
                    const uint32_t w = mine ? word[u * nthreads] : 0u;
                    const uint32_t bits = __ballot_sync(0xffffffffu, (w & hd) != 0u);
                    if (leader && mine) bout[u * (nthreads >> 5)] = bits;
                }
                if (LOAD && mine) {
                    T::store_lane(&word[u * nthreads], lane, T::llr_word(val[u]));
                    key = max(key, T::range_key(val[u]));
                }
            }
        }
        word += UNR * nthreads;
        if (BITS) bout += UNR * (nthreads >> 5);
        if (LOAD) in += UNR * nthreads;
    }
    return T::range_bad(key);
}

template <class T, class SRC>
__device__ __forceinline__ bool swap_frame(uint32_t *lw, int lane, uint32_t *bits_out, const SRC *src, int n, int nthreads)
{
    if (bits_out && src) return swap_frame_pass<T, SRC, true, true>(lw, lane, bits_out, src, n, nthreads);
    if (src) return swap_frame_pass<T, SRC, false, true>(lw, lane, bits_out, src, n, nthreads);
    if (bits_out) return swap_frame_pass<T, SRC, true, false>(lw, lane, bits_out, src, n, nthreads);
    return false;
}

// Results of the slots in `fin` (unless `first`), then the next frames move in: the lane's messages are cleared
// and its channel values loaded or generated.  Zero messages make the next variable phase produce
// post = LLR, v2c = LLR -- the reference's initialisation (ArrayLDPC_Decoder.cpp:45-61) -- for the new lane while
// it runs an ordinary iteration for the lane's neighbour.  Returns the number of active slots.  Called by every
// thread of the CTA (contains barriers).
template <class T>
__device__ __forceinline__ int finish_and_refill(const KParams &p, Ctrl *ctrl, uint32_t *edge, uint32_t *llr,
                                                 uint32_t fin, bool first, int n, int E, int W, long long frames,
                                                 unsigned int trip, int buf, int nthreads)
{
    const int tid = threadIdx.x, lane_id = tid & 31;
    const int nslots = W * T::LANES;
    const bool mine = tid < nslots && ((fin >> tid) & 1u);  // one thread per stopping slot (all in warp 0)
    const bool emit = !first && (p.bits || p.mc_mode != 0);
    unsigned long long claim = 0;
    bool was_over = false;
    int old_it = 0;
    if (mine) {
        const int s = tid, w = s / T::LANES, lane = s % T::LANES;
        was_over = (ctrl->gflag[buf][w] >> lane) & 1u;
        old_it = (int)(trip - ctrl->start[s]);
        if (!first)
            p.iters[ctrl->fid[s]] = was_over ? -1 : (p.max_iter == 0 ? (int)((ctrl->fail[buf][w] >> lane) & 1u) : old_it);
        // With claim_ahead the slot's next frame was taken from the queue (and its channel values pulled into L2)
        // when the previous one moved in; the atomic issued here is for the frame after, and its round trip
        // overlaps the passes below.
        claim = atomicAdd(p.queue, 1ull);
        const int f = p.claim_ahead ? ctrl->next[s] : queue_frame(p, claim, frames);
        if (p.avail && f >= 0)  // fed launch: the frame's channel values may still be on their way (start of the batch only)
            while (frames_arrived(p, frames) <= (long long)f) __nanosleep(200);
        ctrl->newfid[s] = f;
        if (p.mc_mode == 2 && f >= 0)
            ctrl->lehmer[s] = lehmer_mul((uint32_t)p.mc_seed, lehmer_pow(p.mc_jump, p.mc_first + (unsigned long long)f));
        ctrl->errs[s] = 0u;
    }
    // Decode-from-memory launches that claim ahead need no barrier here: the frame that moves in is ctrl->next[s],
    // written one refill ago, and nothing the owner stores above is read by the other threads before the next barrier.
    // (Monte-Carlo launches read lehmer[] / errs[], sync claims read newfid[], fed launches wait for the owner's
    // arrival check.)
    const bool quick = p.claim_ahead && p.mc_mode == 0 && !p.avail;
    if (!quick) __syncthreads();
    if (p.mc_mode != 0 && emit) {
        // Monte-Carlo mode generates the channel values with another thread-to-variable mapping, so the decoded
        // bits (kept in the channel words) are collected first
        for (uint32_t left = fin; left; left &= left - 1u) {
            const int s = __ffs(left) - 1;
            const int fo = ctrl->fid[s];
            const uint32_t *lw = llr + (size_t)(s / T::LANES) * n;
            if (fo >= 0)
                for (int base = 0; base < n; base += nthreads) {
                    const int v = base + tid;
                    emit_word(p, ctrl, s, fo, v, n, v < n ? T::hd_bit(lw[v], s % T::LANES) : 0u);
                }
        }
        __syncthreads();
    }
    uint32_t bad_slots = 0;
    for (uint32_t left = fin; left; left &= left - 1u) {
        const int s = __ffs(left) - 1;
        const int w = s / T::LANES, lane = s % T::LANES;
        const int fo = emit ? ctrl->fid[s] : -1, fn = quick ? ctrl->next[s] : ctrl->newfid[s];
        uint32_t *ew = edge + (size_t)w * E, *lw = llr + (size_t)w * n;
        if (p.mc_mode == 0) {
            uint32_t *bits_out = (fo >= 0 && p.bits) ? p.bits + (size_t)fo * p.nw32 : nullptr;
            bool bad;
            if (p.llr_bits == 16)
                bad = swap_frame<T>(lw, lane, bits_out, fn >= 0 ? reinterpret_cast<const int16_t *>(p.llr) + (size_t)fn * n : nullptr, n, nthreads);
            else
                bad = swap_frame<T>(lw, lane, bits_out, fn >= 0 ? reinterpret_cast<const int *>(p.llr) + (size_t)fn * n : nullptr, n, nthreads);
            if (bad) bad_slots |= 1u << s;
        } else if (p.mc_mode == 1) {
            for (int q = tid; 4 * q < n; q += nthreads) {
                float z[4] = {0.f, 0.f, 0.f, 0.f};
                if (fn >= 0) philox_normals(p, p.mc_first + (unsigned long long)fn, (uint32_t)q, z);
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int v = 4 * q + h;
                    if (v < n) {
                        bool bad;
                        const int val = fn >= 0 ? quantise_llr(p, (double)z[h], cw_bit(p, fn, v)) : 0;
                        T::store_lane(&lw[v], lane, T::llr_lane(val, bad));
                        if (bad) bad_slots |= 1u << s;
                    }
                }
            }
        } else {
            for (int v = tid; v < n; v += nthreads) {
                bool bad;
                const int val = fn >= 0 ? quantise_llr(p, lehmer_normal(lehmer_mul(ctrl->lehmer[s], p.mc_pow[v])), cw_bit(p, fn, v)) : 0;
                T::store_lane(&lw[v], lane, T::llr_lane(val, bad));
                if (bad) bad_slots |= 1u << s;
            }
        }
        // its messages start from zero
        if (!first) {
#pragma unroll 4
            for (int i = tid; i < E; i += nthreads) T::store_lane(&ew[i], lane, 0u);
        }
    }
    for (int s = 0; bad_slots; ++s, bad_slots >>= 1)  // judged by the coming trip's stop decision
        if (bad_slots & 1u) atomicOr(&ctrl->gflag[buf == 2 ? 0 : buf + 1][s / T::LANES], 1u << (s % T::LANES));
    if (p.mc_mode != 0 && p.mc_pin_count > 0) {  // shortening: known positions pinned (PerfTest.cpp:410-414)
        __syncthreads();
        for (int i = tid; i < p.mc_pin_count; i += nthreads) {
            const int v = p.mc_pin[i];
            for (uint32_t left = fin; left; left &= left - 1u) {
                const int s = __ffs(left) - 1;
                if (ctrl->newfid[s] < 0) continue;
                bool bad;
                T::store_lane(&llr[(size_t)(s / T::LANES) * n + v], s % T::LANES, T::llr_lane(p.mc_pin_value, bad));
                if (bad) atomicOr(&ctrl->gflag[buf == 2 ? 0 : buf + 1][s / T::LANES], 1u << (s % T::LANES));
            }
        }
    }
    __syncthreads();
    int ahead = -1;
    if (mine) {
        const int s = tid;
        if (!first && p.mc_mode != 0 && !was_over) {  // flagged frames are counted by their exact re-decode
            const unsigned int e = ctrl->errs[s];
            if (p.mc_frame_err) p.mc_frame_err[ctrl->fid[s]] = (unsigned short)min(e, 65535u);
            atomicAdd(&p.mc_counters[0], 1ull);
            if (e) atomicAdd(&p.mc_counters[1], 1ull);
            if (e) atomicAdd(&p.mc_counters[2], (unsigned long long)e);
            atomicAdd(&p.mc_counters[3], (unsigned long long)old_it);
        }
        if (!first && p.done_count && ctrl->fid[s] >= 0) {
            // results of the frame that left (iteration count by this thread, bits by the ballot leaders before the
            // barrier behind us) are complete: count it, and tell the host when its chunk is
            const int fo = ctrl->fid[s], chunk = fo / p.done_chunk;
            const long long lo = (long long)chunk * p.done_chunk;
            const unsigned int size = (unsigned int)min((long long)p.done_chunk, frames - lo);
            __threadfence();
            if (atomicAdd(&p.done_count[chunk], 1u) + 1u == size) {
                __threadfence_system();
                p.done_flag[chunk] = 1u;
            }
        }
        ctrl->fid[s] = ctrl->newfid[s];
        ctrl->start[s] = trip + 1u;  // its first (initialising) variable phase runs in the coming trip
        if (p.claim_ahead) {
            ahead = queue_frame(p, claim, frames);
            ctrl->next[s] = ahead;
        }
    }
    if (p.claim_ahead && p.mc_mode == 0 && tid < 32 && !p.avail) {
        // pull the channel values of the frames just claimed into L2: they are read one frame time from now
        const size_t bytes = (size_t)n * (p.llr_bits >> 3);
        for (uint32_t left = fin; left; left &= left - 1u) {
            const int f = __shfl_sync(0xffffffffu, ahead, __ffs(left) - 1);
            if (f >= 0) {
                const char *first_byte = reinterpret_cast<const char *>(p.llr) + (size_t)f * bytes;
                const char *line = first_byte - (reinterpret_cast<uintptr_t>(first_byte) & 127u) + (size_t)lane_id * 128u;
                for (; line < first_byte + bytes; line += 32 * 128) prefetch_l2(line);
            }
        }
    }
    __syncthreads();
    int active = 0;
    for (int s = 0; s < nslots; ++s) active += ctrl->fid[s] >= 0;
    return active;
}

// Stop / continue decision of every slot, taken after the check phase (its votes are the syndrome of the state the
// variable phase before it left): early termination after every iteration (ArrayLDPC_Decoder.cpp:164-167),
// MAX_ITER (:63), decode_fixpoint's pre-check on the channel hard decisions (:443-450, iteration count 0), the
// hard-decision-only mode (max_iter == 0) and lanes that left the packed range.  Every warp evaluates all slots
// (lane = slot) from words nobody writes before the next barrier, so the mask of stopping slots is the same in
// every thread and no barrier is needed to agree on it.
template <class T>
__device__ __forceinline__ uint32_t stop_decision(const KParams &p, Ctrl *ctrl, int W, unsigned int trip, int buf)
{
    const int s = threadIdx.x & 31;
    bool stop = false;
    if (s < W * T::LANES && ctrl->fid[s] >= 0) {
        const int w = s / T::LANES, lane = s % T::LANES;
        const int it = (int)(trip - ctrl->start[s]);  // 0: the frame has only been initialised
        const bool pass = !((ctrl->fail[buf][w] >> lane) & 1u);
        const bool over = (ctrl->gflag[buf][w] >> lane) & 1u;
        stop = it >= p.max_iter || (pass && (it >= 1 || p.precheck)) || over;
    }
    const uint32_t fin = __ballot_sync(0xffffffffu, stop);
    // the buffer the trip after next votes into (last read one decision ago, before two barriers)
    if (threadIdx.x < MAX_W) {
        const int clr = buf == 0 ? 2 : buf - 1;
        ctrl->fail[clr][threadIdx.x] = 0u;
        ctrl->gflag[clr][threadIdx.x] = 0u;
    }
    return fin;
}

// ------------------------------------------------------------------------------------------
// the kernel
//   DC   largest check degree, REG: every check has degree DC
//   DV   largest variable degree with an exact body
//   NI   word sets per thread in the check phase (W is a multiple of NI)
//   M, N compile-time m and n of the named codes (0 = read them from the parameters); EA: words per word set if
//        that is less than DC*M (irregular named code)
//   ARRP forward square array code with this circulant size: edge addresses in closed form (variable_phase_array)
//   VMASK, CMASK  irregular codes: the variable / check degrees that get exact bodies (all ones: every degree up to DV / DC)
//   WS, TS  word sets per CTA and CTA size as compile-time constants (0: p.W / blockDim.x): the named codes run with the
//        shape their plan computes; the loops over word sets and over the CTA's passes then have one shape and constant
//        trip counts (measured +3 to +6 % on all four codes, profiles/r02/launch_shape_sweep.txt)
//   CS   regular codes with NI == 1: checks of one word set interleaved per thread (m is a multiple of CS)
//   MAXT, NCTA  launch bounds: CTA size and co-resident CTAs per SM.  Co-resident CTAs drift out of phase, so
//        one's latency-bound variable phase overlaps the other's ALU-bound check phase.
//
// One trip of the main loop is one iteration of ArrayLDPC_Decoder.cpp:63-168 for every resident frame, rotated so
// that the syndrome needs no pass of its own:
//     variable phase i   (for a frame that just moved in: the initialisation, :45-61)
//     check phase i+1    which first XORs the hard decisions it reads -> syndrome after iteration i
//     stop decision      frames whose syndrome is zero (or that reached MAX_ITER) leave with the hard decisions
//                        the variable phase kept in the channel words; the messages check phase i+1 wrote for
//                        them are never used
// so a frame that stops costs one check phase more than the reference executes, and a converged frame is
// found by the same two barriers per trip that the phases need anyway.
// ------------------------------------------------------------------------------------------
template <class T, int DC, bool REG, int DV, int NI, int MAXT, int NCTA, int M, int N, int EA, int ARRP, unsigned VMASK, unsigned CMASK, int WS, int TS, int CS = 1>
__global__ void __launch_bounds__(MAXT, NCTA) decode_kernel(const __grid_constant__ KParams p)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int tid = threadIdx.x, nthreads = TS ? TS : (int)blockDim.x;
    // the named codes get their dimensions as compile-time constants: every k*m word offset of the check
    // phase then folds into the load/store immediate
    const int n = N ? N : p.n, m = M ? M : p.m, E = EA ? EA : (M ? DC * M : p.E), W = WS ? WS : p.W;
    uint32_t *edge = smem;                 // [W][E]
    uint32_t *llr = edge + (size_t)W * E;  // [W][n] channel values (+ the hard decision of the last posterior)
    Ctrl *ctrl = reinterpret_cast<Ctrl *>(llr + (size_t)W * n);
    // degree tables next to the control block (irregular codes dispatch on them once per node and phase)
    constexpr bool REGV = PREFETCH_VEDGE && M != 0 && REG && DV <= 8 && ARRP == 0;
    uint8_t *cdeg_s = reinterpret_cast<uint8_t *>(ctrl + 1);
    uint8_t *vdeg_s = cdeg_s + (REG ? 0 : ((m + 15) & ~15));
    if (!REG) for (int i = tid; i < m; i += nthreads) cdeg_s[i] = p.cdeg[i];
    if (!REGV && !ARRP) for (int i = tid; i < n; i += nthreads) vdeg_s[i] = p.vdeg[i];
    const int nslots = W * T::LANES;
    const long long frames = p.count ? (long long)*p.count : p.frames;

    for (int i = tid; i < W * (E + n); i += nthreads) smem[i] = 0u;
    if (tid < MAX_W)
        for (int b = 0; b < 3; ++b) { ctrl->fail[b][tid] = 0u; ctrl->gflag[b][tid] = 0u; }
    if (tid < MAX_SLOTS) {
        ctrl->fid[tid] = -1; ctrl->start[tid] = 0u; ctrl->newfid[tid] = -1;
        ctrl->next[tid] = (p.claim_ahead && tid < nslots) ? queue_frame(p, atomicAdd(p.queue, 1ull), frames) : -1;
    }
    __syncthreads();

    const int items = (W / NI) * (m / CS);  // (word-set group, check)
#ifdef LDPC_PHASE_TIMING
    long long t_phase[4] = {0, 0, 0, 0}, t_mark = clock64();
#define LDPC_MARK(k) do { long long t_now = clock64(); t_phase[k] += t_now - t_mark; t_mark = t_now; } while (0)
#else
#define LDPC_MARK(k) do { } while (0)
#endif

    uint32_t fin = (nslots >= 32) ? 0xffffffffu : ((1u << nslots) - 1u);  // slots to (re)fill
    bool first = true;
    // trip t votes into buffer t % 3; `buf` is the buffer of the trip that ended last (what a refill reports from)
    unsigned int trip = 0;
    int buf = 2;
    for (;;) {
        if (fin) {
            // trip - 1 is the trip whose stop decision released the slots
            const int active = finish_and_refill<T>(p, ctrl, edge, llr, fin, first, n, E, W, frames, trip - 1u, buf, nthreads);
            if (active == 0) break;
            first = false;
        }
        buf = buf == 2 ? 0 : buf + 1;
        LDPC_MARK(0);
        if (p.post || p.v2c) variable_phase_parity<T, DV, REGV, ARRP, VMASK>(p, ctrl, ctrl->gflag[buf], edge, llr, W, n, m, E, vdeg_s, nthreads);
        else if (ARRP) variable_phase_array<T, DV, (ARRP ? ARRP : 1), false>(p, ctrl, ctrl->gflag[buf], edge, llr, W, n, m, E, nthreads);
        else variable_phase<T, DV, false, REGV, VMASK>(p, ctrl, ctrl->gflag[buf], edge, llr, W, n, E, vdeg_s, nthreads);
        __syncthreads();
        LDPC_MARK(1);
        check_phase<T, DC, REG, NI, CMASK, CS>(p, ctrl->fail[buf], edge, cdeg_s, items, m, E, W, nthreads);
        __syncthreads();
        LDPC_MARK(2);
        fin = stop_decision<T>(p, ctrl, W, trip, buf);
        ++trip;
        LDPC_MARK(3);
    }
#ifdef LDPC_PHASE_TIMING
    if (tid == 0 && blockIdx.x == 0 && trip > 0)
        printf("phase cycles (CTA 0): results+refill %lld variable %lld check %lld stop decision %lld | trips %u\n", t_phase[0],
               t_phase[1], t_phase[2], t_phase[3], trip);
#endif
}

}  // namespace ldpc
