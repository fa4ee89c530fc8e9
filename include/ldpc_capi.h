/*
 * ldpc_capi.h -- C ABI of the B200 fixed-point LDPC decode engine (libldpc_b200.so).
 *
 * This is the drop-in boundary for the reference's decode path: plain pointers and sizes,
 * no C++/torch types, no exceptions.  Every entry point names the reference interface it
 * replaces (file:line in tyc85/FixedPointLDPC).  The C++ facade in ArrayLDPCMacro.h /
 * PerfTest.h (same directory) re-creates the reference's class and driver names on top of
 * these calls; INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - all functions returning int return LDPC_OK (0) or a negative ldpc_status;
 *     constructors return NULL and store the status in *err (err may be NULL);
 *   - a decoder handle is bound to one CUDA device and one stream and must be used from one
 *     host thread at a time (the reference is not re-entrant at all:
 *     ArrayLDPC_Decoder.cpp:21-37, 430-440);
 *   - there is NO CPU fallback: without a CUDA device every decoder call fails with
 *     LDPC_ERR_NO_DEVICE.
 */
#ifndef LDPC_CAPI_H
#define LDPC_CAPI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ldpc_code ldpc_code;       /* parity-check tables (host)           */
typedef struct ldpc_gen ldpc_gen;         /* generator equations, Format B (host) */
typedef struct ldpc_decoder ldpc_decoder; /* device-resident decode engine        */

typedef enum {
    LDPC_OK = 0,
    LDPC_ERR_IO = -1,          /* file cannot be opened / short read                      */
    LDPC_ERR_FORMAT = -2,      /* file parsed but inconsistent (vlist <-> clist, ranges)  */
    LDPC_ERR_ARG = -3,         /* NULL / out-of-range argument                            */
    LDPC_ERR_CUDA = -4,        /* CUDA runtime error, see ldpc_last_error()               */
    LDPC_ERR_UNSUPPORTED = -5, /* code does not fit the kernels (degree / smem limits)    */
    LDPC_ERR_NOMEM = -6,
    LDPC_ERR_NO_DEVICE = -7    /* no CUDA device: the engine has no CPU path              */
} ldpc_status;

/* On-disk H formats. */
enum {
    LDPC_FMT_AUTO = 0,
    LDPC_FMT_A = 1, /* zero-based alist-like, what FP_Decoder::ReadH parses              */
    LDPC_FMT_C = 3  /* legacy one-based check lists (H2212_316_array_cut79.txt)          */
};

const char *ldpc_strerror(int status);
/* Text of the most recent failure on this thread (CUDA error string, offending token, ...). */
const char *ldpc_last_error(void);

/* ------------------------------------------------------------------ code tables (host) */

/* Replaces FP_Decoder::ReadH (ArrayLDPC_Decoder.cpp:642-674), whose file name is hard
 * coded (:646) and whose reads are unchecked.  Format C has no reader in the reference. */
ldpc_code *ldpc_code_load(const char *path, int format, int *err);

/* Build from check lists: clist[c*cstride + k], k < cdeg[c], zero-based.  Rows are sorted
 * ascending (the reference's addr_count slot lookup needs that, ArrayLDPC_Decoder.cpp:121-154). */
ldpc_code *ldpc_code_from_checks(int n, int m, const int *cdeg, const int *clist, int cstride, int *err);

/* Array-code structure == class ROM (ArrayLDPCMacro.h:42-82): check (i, t) of row group i
 * touches variable b*p + ((t + sign*row_mult[i]*col_sel[b]) mod p) of every selected column
 * group b; sign = +1 ("forward", what decode_fixpoint addresses, ArrayLDPC_Decoder.cpp:474-476)
 * or -1 when backward != 0.  row_mult/col_sel may be NULL for 0..nrows-1 / 0..ncols-1. */
ldpc_code *ldpc_code_array(int p, int nrows, const int *row_mult, int ncols, const int *col_sel,
                           int backward, int *err);

void ldpc_code_free(ldpc_code *code);

/* n, m, edges, dc_max, dv_max  (NUM_VAR, NUM_CHK, -, CHK_DEG, VAR_DEG of ArrayLDPCMacro.h:18-24). */
int ldpc_code_dims(const ldpc_code *code, int *n, int *m, int *edges, int *dc_max, int *dv_max);
/* Dense copies of the four ReadH tables; short rows are padded with -1.  Any pointer may be NULL. */
int ldpc_code_tables(const ldpc_code *code, int *vdeg, int *cdeg, int *vlist, int *clist);
/* Rate reported by ROM::getRate (ArrayLDPCMacro.h:60) for array codes built by ldpc_code_array,
 * (n - m)/n otherwise. */
double ldpc_code_rate(const ldpc_code *code);
/* Writes Format A exactly as codes/alist_from_arraycode.m:28-60 does. */
int ldpc_code_save(const ldpc_code *code, const char *path);

/* ------------------------------------------------------------------ decoder (device)  */

typedef struct {
    int max_iter;   /* MAX_ITER, ArrayLDPCMacro.h:17; default 30                                  */
    int precheck;   /* 1: decode_fixpoint semantics (return 0 when the channel hard decision
                       already satisfies H, ArrayLDPC_Decoder.cpp:443-450); 0: decode_general_fp */
    int device;     /* CUDA device ordinal                                                        */
    int precision;  /* 0 = auto (packed int16x2 kernel + exact int32 re-decode of frames whose
                       values leave the 13-bit guard range), 32 = int32 kernel only, 16 = packed
                       only (frames that leave the range report iters = -1)                       */
    int threads;    /* CTA size override, 0 = auto                                                */
    int frames_per_cta; /* resident frame slots per CTA override, 0 = auto                        */
} ldpc_decoder_cfg;

void ldpc_decoder_cfg_default(ldpc_decoder_cfg *cfg);

ldpc_decoder *ldpc_decoder_create(const ldpc_code *code, const ldpc_decoder_cfg *cfg, int *err);
void ldpc_decoder_destroy(ldpc_decoder *dec);

/* Batched FP_Decoder::decode_general_fp / decode_fixpoint (ArrayLDPC_Decoder.cpp:18-171,
 * 422-639) on HOST buffers; copies in, decodes on the GPU, copies out, returns when done.
 *   llr    [frames][n]            int32, the reference's `const int *LLR`
 *   iters  [frames]               return value of the reference call (0..max_iter)
 *   bits   [frames][ceil(n/32)]   DecodedCodeword, bit v%32 of word v/32            (or NULL)
 *   post   [frames][n]            Posteriori_fp                                     (or NULL)
 *   v2c    [frames][dc_max][m]    EdgeRAM image, slot-major (ArrayLDPCMacro.h:101,162);
 *                                 slots >= cdeg[c] are written as 0                 (or NULL)
 * For frames with iters == 0 (pre-check hit) post/v2c hold the channel values; the reference
 * leaves them stale (quirk Q6). */
int ldpc_decode_batch(ldpc_decoder *dec, const int32_t *llr, size_t frames, int32_t *iters,
                      uint32_t *bits, int32_t *post, int32_t *v2c);

/* The same call for callers that already hold 16-bit quantised LLRs (the reference's channel values stay far
 * below 2^15: SURVEY.md 0.4): half the host->device bytes of the `const int *LLR` layout.  Results are identical
 * to ldpc_decode_batch on the widened values. */
int ldpc_decode_batch_i16(ldpc_decoder *dec, const int16_t *llr, size_t frames, int32_t *iters,
                          uint32_t *bits, int32_t *post, int32_t *v2c);

/* Batched FP_Decoder::decode_general(const double *) (ArrayLDPC_Decoder.cpp:735-933): the reference's floating-point
 * decoder with the exact box-plus sxor(double, double) (:724-732) and checkPost() (:335-372), in FP64 on the GPU.
 *   llr [frames][n] double; iters / bits as above; post [frames][n] double (Posteriori); v2c [frames][dc_max][m]
 *   double (EdgeRAM[slot].BRAM[check]).
 * Same schedule and the same evaluation order of every sum as the reference; libm's log/exp differ from the host's
 * by an ulp at most, so results agree with the reference to a relative 1e-9 on posteriors and messages and -- on
 * all but marginal frames -- exactly in iteration counts and bits (tests/test_gpu_f64.py). */
int ldpc_decode_batch_f64(ldpc_decoder *dec, const double *llr, size_t frames, int32_t *iters, uint32_t *bits,
                          double *post, double *v2c);

/* Same on DEVICE buffers, asynchronous on `stream` (a cudaStream_t, NULL = the decoder's own
 * stream).  llr_bits = 32 (int32) or 16 (int16) selects the input element type. */
int ldpc_decode_batch_device(ldpc_decoder *dec, const void *d_llr, int llr_bits, size_t frames,
                             int32_t *d_iters, uint32_t *d_bits, int32_t *d_post, int32_t *d_v2c,
                             void *stream);

/* Batched FP_Decoder::hardDecision / checkPost_fp / checkPost_fp_general (ArrayLDPC_Decoder.cpp:270-333,
 * 375-420) on HOST buffers: bits = (value > 0 ? 0 : 1) and fail[f] = 1 if any check is unsatisfied, else 0.
 * `values` are channel LLRs (hardDecision) or posteriors (checkPost*). */
int ldpc_hard_decision_batch(ldpc_decoder *dec, const int32_t *values, size_t frames, int32_t *fail, uint32_t *bits);

/* CUDA device ordinal the decoder is bound to, its parity-check tables, its MAX_ITER. */
int ldpc_decoder_device(const ldpc_decoder *dec);
const ldpc_code *ldpc_decoder_code(const ldpc_decoder *dec);
int ldpc_decoder_max_iter(const ldpc_decoder *dec);

/* Blocks until everything queued on the decoder's streams has finished. */
int ldpc_decoder_sync(ldpc_decoder *dec);

typedef struct {
    uint64_t kernel_launches;  /* decode kernels launched since creation                     */
    uint64_t frames;           /* frames decoded                                              */
    uint64_t fallback_frames;  /* frames re-decoded by the int32 kernel (guard range left)    */
    int threads;               /* CTA size in use (packed / int32)                            */
    int threads32;
    int frames_per_cta;        /* resident frame slots per CTA (packed / int32)               */
    int frames_per_cta32;
    int grid;                  /* CTAs per launch                                             */
    int smem_bytes;            /* dynamic shared memory per CTA (packed kernel)               */
    int smem_bytes32;
    int resident_ctas_per_sm;  /* cudaOccupancyMaxActiveBlocksPerMultiprocessor of the last decode launch     */
    int launch_smem_bytes;     /* dynamic shared memory of the last decode launch                             */
} ldpc_decoder_stats;

int ldpc_decoder_get_stats(const ldpc_decoder *dec, ldpc_decoder_stats *out);

/* ------------------------------------------------------------------ generator / encoder (host) */

/* Replaces FP_Encoder::FP_Encoder(char*, int) (ArrayLDPC_Encoder.cpp:34-157): parses Format B
 * ("n rows / dv dc / ColumnFlag[n] / ChkDeg[rows] / rows lists"), flag 1 = parity column. */
ldpc_gen *ldpc_gen_load(const char *path, int *err);
/* Derive generator equations from H by GF(2) elimination (replaces the offline MATLAB tools
 * codes/simplfy_generator_alist.m, codes/alist_to_binary.m).  parity_cols[nparity] fixes which columns
 * are solved for (e.g. the flagged columns of an existing Format B file); NULL picks pivots greedily from
 * the LAST column backwards, so the information bits occupy the leading columns.  rank(H) equations result. */
ldpc_gen *ldpc_gen_from_code(const ldpc_code *code, const int32_t *parity_cols, int nparity, int *err);
/* Writes Format B as FP_Encoder reads it. */
int ldpc_gen_save(const ldpc_gen *gen, const char *path);
void ldpc_gen_free(ldpc_gen *gen);
/* n, rows (= parity equations), k = n - rows */
int ldpc_gen_dims(const ldpc_gen *gen, int *n, int *rows, int *k);
/* InfoIndex[k] / ParityIndex[rows] (ArrayLDPC_Encoder.cpp:56-69); either may be NULL */
int ldpc_gen_indices(const ldpc_gen *gen, int32_t *info_index, int32_t *parity_index);
/* FP_Encoder::encode(char*, int) (ArrayLDPC_Encoder.cpp:160-225): info bytes, LSB first, the last byte
 * supplies k % 8 bits; codeword[n] receives 0/1. */
int ldpc_gen_encode(const ldpc_gen *gen, const char *info, int info_len, uint8_t *codeword);

/* Batched GPU encoder: parity_i = popcount(G_i & info) & 1 on bit-packed rows (the reference encodes one
 * message at a time on the host, ArrayLDPC_Encoder.cpp:199-210).
 *   info       [frames][ceil(k/8)] message bytes, LSB first like FP_Encoder::encode (:171-183)
 *   codewords  [frames][ceil(n/32)] packed codeword bits, bit v%32 of word v/32 */
int ldpc_encode_batch(ldpc_gen *gen, int device, const uint8_t *info, size_t frames, uint32_t *codewords);
/* Device buffers: d_info [frames][ceil(k/32)] packed message words, d_codewords as above; asynchronous. */
int ldpc_encode_batch_device(ldpc_gen *gen, int device, const uint32_t *d_info, size_t frames, uint32_t *d_codewords,
                             void *stream);

/* ------------------------------------------------------------------ Monte-Carlo mode (device) */

/* Noise streams. */
enum {
    LDPC_STREAM_PHILOX = 1,   /* counter-based Philox4x32-10, counter = (global frame, variable/4) */
    LDPC_STREAM_REFERENCE = 2 /* the reference's own stream: Lehmer LCG rngs.cpp:52-69 + Odeh-Evans
                                 Normal rvgs.cpp:152-181, one uniform per bit, regenerated in parallel by
                                 O(log) skip-ahead; reproduces the reference's frames bit for bit */
};

/* One simulation point == the loop body of ArrayLDPC_Debug_Wifi / ArrayLDPC_Debug / ArrayLDPC_PerfTest /
 * ArrayLDPC_Debug_Shorten (PerfTest.cpp:97-135, 276-311, 385-426, 491-512): BPSK + AWGN + quantiser
 * (LLR_fp = int(2*snr*(1 - 2c + N(0,sigma)) * 2^4), truncation, no clipping), decode, calculateBER. */
typedef struct {
    double snr;                /* the drivers' `snr` / `EbN0` variable: 2*10^(dB/10)*R (PerfTest.cpp:62,159,253,487) */
    double sigma;              /* sqrt(1/snr) */
    int stream;                /* LDPC_STREAM_*                                                          */
    uint64_t seed;             /* Philox key, or the Lehmer state before global frame 0 (rngs.cpp:45: 123456789) */
    uint64_t first_frame;      /* global index of the first frame of this call (frames are independent)  */
    const uint8_t *codeword;   /* [n] transmitted bits, NULL = all-zero codeword                         */
    const uint32_t *d_codewords; /* DEVICE pointer, [frames][ceil(n/32)] packed: frame f sends row f (e.g. the output of
                                  ldpc_encode_batch_device for random messages); overrides `codeword` when set   */
    const int32_t *info_index; /* [info_count] positions calculateBER compares (ArrayLDPC_Decoder.cpp:707-722),
                                  NULL = all n positions                                                  */
    int info_count;
    const int32_t *pin_index;  /* shortening: LLR_fp[pin_index[i]] = pin_value after the channel
                                  (ArrayLDPC_Debug_Shorten, PerfTest.cpp:410-414)                         */
    int pin_count;
    int pin_value;
} ldpc_mc_cfg;

typedef struct {
    uint64_t frames;       /* Counter                                                */
    uint64_t frame_errors; /* pckerror: frames with at least one info-bit error      */
    uint64_t bit_errors;   /* biterror                                               */
    uint64_t iter_sum;     /* sum of the decoder's return values                     */
} ldpc_mc_counters;

/* Simulates `frames` frames starting at cfg->first_frame.  Host outputs (any may be NULL):
 *   frame_err [frames] info-bit errors of every frame (saturating at 65535), in frame order, so the caller
 *             can apply the reference's sequential stopping rule `while(pckerror < 100)` exactly;
 *   iters     [frames] the decoder's return value per frame;
 *   totals    sums over the call. */
int ldpc_mc_run(ldpc_decoder *dec, const ldpc_mc_cfg *cfg, size_t frames, uint16_t *frame_err, int32_t *iters,
                ldpc_mc_counters *totals);

/* Same with DEVICE outputs, asynchronous on `stream`: d_counters is uint64[4] in ldpc_mc_counters order and is
 * accumulated into (the caller zeroes it). */
int ldpc_mc_run_device(ldpc_decoder *dec, const ldpc_mc_cfg *cfg, size_t frames, uint16_t *d_frame_err,
                       int32_t *d_iters, uint64_t *d_counters, void *stream);

/* ------------------------------------------------------------------ Monte-Carlo point on all GPUs of the box */

/* Stopping rule of the reference's frame loops, applied to the frames in global index order, so that the counters are
 * the ones the sequential loop produces whatever the number of GPUs and the batch size:
 *   target_block_errors  `while(pckerror < 100)`   (PerfTest.cpp:97, 276, 385, 491); 0 = none
 *   max_frames           `while(Counter < MaxPckNum)` (PerfTest.cpp:580); 0 = none
 * count_iterations: ArrayLDPC_PerfTest / ArrayLDPC_TimeTrial take the decoder's return value as the block's error
 * count (quirk Q9, PerfTest.cpp:507-511, 596-600) instead of calculateBER's. */
typedef struct {
    uint64_t target_block_errors;
    uint64_t max_frames;
    size_t frames_per_round;   /* frames per GPU and round (0 = 131072)                                            */
    int count_iterations;
    int32_t *iters_out;        /* optional host buffer [iters_cap]: the decoder's return value of every counted    */
    size_t iters_cap;          /*   frame, in frame order (ArrayLDPC_Debug_Shorten prints them, PerfTest.cpp:419)   */
} ldpc_mc_stop;

typedef struct {
    uint64_t frames;           /* Counter                                                                           */
    uint64_t block_errors;     /* pckerror                                                                          */
    uint64_t errors;           /* biterror (or the iteration sum under count_iterations)                            */
    uint64_t iter_sum;         /* sum of the decoder's return values over the counted frames                        */
    uint64_t iter_hist[32];    /* frames per return value 0..31                                                     */
    uint64_t rounds;           /* rounds (= counter all-reduces) executed                                           */
    int reached;               /* 1: stopped on target_block_errors                                                 */
    int devices;
    double seconds;            /* wall time of the run                                                              */
} ldpc_mc_result;

typedef struct ldpc_mc_group ldpc_mc_group;

/* One decoder handle per GPU (all for the same code and configuration, each on its own device).  Creates the NCCL
 * communicator over those devices (none for a single one) and one host thread per device per run.  Device r of R
 * simulates the frames [(round*R + r)*B, +B) of every round; per round the counters travel through ONE ncclAllReduce
 * (320 bytes over NVLink), per-frame results only in the final round. */
ldpc_mc_group *ldpc_mc_group_create(ldpc_decoder *const *decoders, int ndev, int *err);
int ldpc_mc_group_size(const ldpc_mc_group *group);
int ldpc_mc_group_run(ldpc_mc_group *group, const ldpc_mc_cfg *cfg, const ldpc_mc_stop *stop, ldpc_mc_result *result);
void ldpc_mc_group_destroy(ldpc_mc_group *group);
/* create + run + destroy */
int ldpc_mc_run_multi(ldpc_decoder *const *decoders, int ndev, const ldpc_mc_cfg *cfg, const ldpc_mc_stop *stop,
                      ldpc_mc_result *result);

/* Channel only: the quantised LLRs [frames][n] the simulation above would decode (host buffer). */
int ldpc_mc_channel(ldpc_decoder *dec, const ldpc_mc_cfg *cfg, size_t frames, int32_t *llr);

/* Number of CUDA devices visible, or a negative status. */
int ldpc_device_count(void);

#ifdef __cplusplus
}
#endif
#endif
