/*
 * ArrayLDPCMacro.h -- source-compatible facade of the reference's codec header for the B200 engine.
 *
 * Same names as tyc85/FixedPointLDPC's ArrayLDPCMacro.h (enums :17-40, ROM :42-82, ControlFSM :109-118,
 * FP_Decoder :121-176, FP_Encoder :179-214), but every decode runs on the GPU through the C ABI in
 * ldpc_capi.h.  The classes here are header-only glue: they own no algorithm.  A driver written against the
 * reference header (PerfTest.cpp, Wrapper.cpp) compiles against this one and links libldpc_b200.so.
 *
 * Code selection.  The reference picks its code with a compile-time enum (:18-24, shipped = 802.11).  Define
 * LDPC_CODE_VARIANT before including this header to do the same:
 *     0 (default)  802.11n n=1944, tables from ReadH()                       (enum CodeWifi, :18-20)
 *     1            array p=47 r=5 n=2209, implicit ROM structure             (commented enum, :22-24)
 *     2            array p=47 r=24 n=2209
 *     3            shortened p=79 array code n=2212 (H2212_316_array_cut79)
 * The engine itself is not compiled per code: FP_Decoder(path) and FP_Decoder(ldpc_code*) take any code at
 * run time.
 *
 * Differences from the reference that a caller can observe (each one fenced on purpose):
 *   - decode_fixpoint without a preceding setState(PCV) while the FSM sits in C2V (the reference would keep
 *     iterating on the previous frame's messages, quirk Q7) is rejected: it prints to cerr and returns -1;
 *   - (kept, not fenced) after a pre-check hit decode_fixpoint returns 0 and Posteriori_fp / EdgeRAM keep the
 *     previous frame's values exactly like the reference (quirk Q6); only the batch C ABI reports channel values;
 *   - check_fp(int*) evaluates the intended parity check (the reference indexes out of range, :224);
 *   - decode_general(const double*) (dead double-precision path) is not provided by the engine: returns -1;
 *   - I/O errors throw std::runtime_error instead of system("pause"); exit(0).
 */
#ifndef ARRAY_MACRO_H
#define ARRAY_MACRO_H

#include <math.h>
#include <stdint.h>

#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "ldpc_capi.h"

using namespace std;

#ifndef LDPC_CODE_VARIANT
#define LDPC_CODE_VARIANT 0
#endif

enum Simulation { MAX_ITER = 30, NUM_PEEK = 1000000, SEED = 100 };
#if LDPC_CODE_VARIANT == 0
enum CodeWifi { NUM_VAR = 1944, NUM_CHK = 972, NUM_CGRP = 12, NUM_VGRP = 24, CHK_DEG = 8, VAR_DEG = 11,
                P = 81, CIR_SIZE = 81, INFO_LENGTH = 972, CWD_LENGTH = 1944 };
#elif LDPC_CODE_VARIANT == 1
enum Code { NUM_VAR = 2209, NUM_CHK = 235, NUM_CGRP = 5, VAR_DEG = 5, NUM_VGRP = 47, CHK_DEG = 47,
            P = 47, CIR_SIZE = 47, INFO_LENGTH = 1978, CWD_LENGTH = 2209 };
#elif LDPC_CODE_VARIANT == 2
enum Code { NUM_VAR = 2209, NUM_CHK = 1128, NUM_CGRP = 24, VAR_DEG = 24, NUM_VGRP = 47, CHK_DEG = 47,
            P = 47, CIR_SIZE = 47, INFO_LENGTH = 1104, CWD_LENGTH = 2209 };
#elif LDPC_CODE_VARIANT == 3
enum Code { NUM_VAR = 2212, NUM_CHK = 316, NUM_CGRP = 4, VAR_DEG = 4, NUM_VGRP = 28, CHK_DEG = 28,
            P = 79, CIR_SIZE = 79, INFO_LENGTH = 1899, CWD_LENGTH = 2212 };
#else
#error "LDPC_CODE_VARIANT must be 0..3"
#endif
enum RAM_Const { RAM_WIDTH = 32, RAM_SLICE = 8, RAM_DEPTH = NUM_CGRP * CIR_SIZE };
enum Precision { WIDTH_MASK = 0x000000ff, SIGN_MASK = 0x00000080, INT_WIDTH = 4, FRAC_WIDTH = 4,
                 INT_WIDTH_NOISE = 4, FRAC_WIDTH_NOISE = 6 };
enum StateFSM { IDLE, PCV, V2C, SXOR, C2V, SIMEND };

/* Circulant shift table + rate of the array structure (reference :42-82). */
class ROM {
public:
    ROM()
    {
        for (int i = 0; i < NUM_CGRP; i++)
            for (int j = 0; j < NUM_VGRP; j++) CirShift[i][j] = (i * j) % P;
        CodeRate = 1 - double(NUM_CGRP * P - NUM_CGRP + 1) / (double(P) * P);
    }
    double getRate() { return CodeRate; }
    int getCirShift(int Chk, int Var) { return CirShift[Chk][Var]; }

private:
    int CirShift[NUM_CGRP][NUM_VGRP];
    double CodeRate;
};

/* One edge-slot bank of the message memory (reference :85-106): bank = slot of the edge inside its check,
 * address = check index.  The engine keeps the messages in GPU shared memory in exactly this [slot][check]
 * order; this class is the host-side view of the image the last decode left (FP_Decoder::getEdgeRAM). */
class Memory {
public:
    Memory() : Address(0), BRAM_fp(RAM_DEPTH, 0) {}
    int rdData() { return BRAM_fp[Address]; }
    void wrtData(int in) { BRAM_fp[Address] = in; }
    void setAddress(int Addr) { Address = Addr; }

private:
    int Address;
    std::vector<int> BRAM_fp;
};

class ControlFSM {
public:
    ControlFSM() { CurState = IDLE; }
    void setState(int in) { CurState = in; }
    int getState() { return CurState; }

private:
    int CurState;
};

namespace ldpc_facade {
inline void fail(const char *what, int status)
{
    throw std::runtime_error(std::string(what) + ": " + ldpc_strerror(status) + " (" + ldpc_last_error() + ")");
}
#if LDPC_CODE_VARIANT == 3
static const int cut79_rows[4] = {0, 1, 3, 4};
static const int cut79_cols[28] = {2, 6, 7, 14, 17, 18, 22, 26, 27, 30, 36, 37, 38, 46, 47, 49, 55, 56, 57, 58,
                                   61, 62, 65, 66, 67, 76, 77, 78};
#endif
}  // namespace ldpc_facade

class FP_Decoder {
public:
    /* Compile-time code like the reference: the array variants carry their structure (class ROM), the 802.11
     * variant gets its tables from ReadH(). */
    FP_Decoder() { init_members(); select_variant(); }
    /* Run-time code: any Format A / Format C file. */
    explicit FP_Decoder(const char *h_path, int format = LDPC_FMT_AUTO) { init_members(); load(h_path, format); }
    ~FP_Decoder() { release(); }

    /* --- the decode entry points (ArrayLDPC_Decoder.cpp:18-171, 422-639) --- */
    int decode_general_fp(const int *LLR) { return run(LLR, 0); }
    int decode_fixpoint(const int *LLR)
    {
        /* :443-450 pre-check, then the FSM gate :462/:488.  In state PCV both are one engine call: the precheck
         * decoder returns 0 for a channel word that already satisfies H and leaves Posteriori_fp / EdgeRAM alone. */
        if (FSM.getState() == PCV) {
            int it = run(LLR, 1);
            if (it == 0) return 0;                         /* :449, state stays PCV */
            FSM.setState(last_syndrome_fail ? C2V : IDLE); /* :619-630 */
            return it;
        }
        if (!hardDecision(LLR)) return 0;
        if (FSM.getState() == C2V) {
            cerr << "FP_Decoder::decode_fixpoint: continuing a non-converged frame without setState(PCV) is not "
                    "supported by the GPU engine" << endl;
            return -1;
        }
        return 0; /* IDLE etc.: the reference's loop body never runs and Iteration stays 0 */
    }
    /* ArrayLDPC_Decoder.cpp:735-933: the floating-point decoder (exact box-plus in FP64 on the GPU); fills Posteriori
     * (getPost), DecodedCodeword and the double image of the edge memory (getEdgeDouble) */
    int decode_general(const double *LLR)
    {
        ensure_decoder();
        std::vector<uint32_t> bits((n_ + 31) / 32);
        Posteriori.resize(n_); EdgeRAM.resize((size_t)dc_ * m_);
        int iters = 0;
        int rc = ldpc_decode_batch_f64(dec_gen, LLR, 1, &iters, &bits[0], &Posteriori[0], &EdgeRAM[0]);
        if (rc != LDPC_OK) ldpc_facade::fail("FP_Decoder: decode_general failed", rc);
        unpack(bits);
        return iters;
    }
    double getPost(int Addr) { return Posteriori[Addr]; }
    void wrtPost(int Addr, double in) { Posteriori[Addr] = in; }
    double getEdgeDouble(int slot, int chk) { return EdgeRAM[(size_t)slot * m_ + chk]; } /* EdgeRAM[slot].BRAM[chk] */
    /* ArrayLDPC_Decoder.cpp:724-732 */
    double sxor(double x, double y)
    {
        double v1 = fabs(x), v2 = fabs(y), sum_abs = v1 + v2, diff_abs = fabs(v1 - v2);
        return sgn(x) * sgn(y) * ((v2 < v1 ? v2 : v1) + log(1 + exp(-sum_abs)) - log(1 + exp(-diff_abs)));
    }

    /* Batched extension: frames x CWD_LENGTH LLRs in, iteration counts out (bits optional, packed). */
    int decode_batch(const int *LLR, size_t frames, int *iters, uint32_t *bits, bool fixpoint)
    {
        ensure_decoder();
        ldpc_decoder *d = fixpoint ? dec_pre : dec_gen;
        return ldpc_decode_batch(d, LLR, frames, iters, bits, NULL, NULL);
    }

    int sgn(double x) { return (x > 0) ? 1 : -1; }
    int sgn(int x) { return (x > 0) ? 1 : -1; }
    int fmin(int x, int y) { return x <= y ? x : y; }
    int fmax(int x, int y) { return x >= y ? x : y; }
    double fmin(double x, double y) { return x <= y ? x : y; }
    /* The pairwise check operator as a scalar helper (ArrayLDPC_Decoder.cpp:677-694); the decode kernels
     * carry their own device version, this one exists because the reference exposes it publicly. */
    int sxor(int x, int y)
    {
        int v1 = abs(x), v2 = abs(y);
        int part1 = Constant - (((v1 + v2) & WIDTH_MASK) >> 2), part2 = Constant - ((abs(v1 - v2) & WIDTH_MASK) >> 2);
        return sgn(x) * sgn(y) * (fmin(v1, v2) + fmax(part1, 0) - fmax(part2, 0));
    }

    /* --- syndrome helpers (ArrayLDPC_Decoder.cpp:210-420): evaluated by the engine, not on the host --- */
    int hardDecision(const int *in) { return syndrome_of(in); }
    int checkPost_fp() { return syndrome_of(&Posteriori_fp[0]); }
    int checkPost_fp_general() { return syndrome_of(&Posteriori_fp[0]); }
    int checkPost() { return checkPost_fp(); }
    int check() { return check_fp(&TrueCodeword[0]); }
    int check_fp(int *word)
    {
        std::vector<int> llr(n_);
        for (int i = 0; i < n_; i++) llr[i] = word[i] ? -1 : 1;
        std::vector<int> keep(DecodedCodeword);
        int r = syndrome_of(&llr[0]);
        DecodedCodeword.swap(keep);
        return r;
    }

    int getPost_fp(int Addr) { return Posteriori_fp[Addr]; }
    void wrtPost(int Addr, int in) { Posteriori_fp[Addr] = in; }
    int getState() { return FSM.getState(); }
    void setState(int in) { FSM.setState(in); }
    double getRate() { return CodeROM.getRate(); } /* "don't use get rate for wifi code" (reference :150) */

    /* --- BER bookkeeping (ArrayLDPC_Decoder.cpp:178-206, 698-722) --- */
    void setInfoBit(char *in, int in_len)
    {
        int counter = 0;
        for (int i = 0; i < in_len - 1; i++)
            for (int j = 0; j < 8; j++) TrueInfoBit[counter++] = (in[i] >> j) & 1;
        for (int j = 0; j < INFO_LENGTH % 8; j++) TrueInfoBit[counter++] = (in[in_len - 1] >> j) & 1;
    }
    void setInfoIndex(int *in) { for (int i = 0; i < INFO_LENGTH; i++) InfoIndex[i] = in[i]; }
    void setCodeword(int *in) { for (int i = 0; i < n_; i++) TrueCodeword[i] = in[i]; }
    int calculateBER()
    {
        for (int i = 0; i < INFO_LENGTH; i++)
            if (DecodedCodeword[InfoIndex[i]] != TrueInfoBit[i]) BitError++;
        return BitError;
    }
    void resetBER() { BitError = 0; }

    /* --- tables --- */
    void ReadH() { load("H_802.11_IndZero.txt", LDPC_FMT_A); } /* hard-coded name, reference :646 */
    void ReadH(const char *path) { load(path, LDPC_FMT_AUTO); }

    /* --- extensions used by the facade drivers and the tests --- */
    int getDecodedBit(int Addr) { return DecodedCodeword[Addr]; }
    int getEdge(int slot, int chk) { return EdgeRAM_fp[(size_t)slot * m_ + chk]; } /* EdgeRAM[slot].BRAM_fp[chk] */
    Memory getEdgeRAM(int slot)                                                    /* copy of bank `slot` */
    {
        Memory bank;
        for (int c = 0; c < m_ && c < RAM_DEPTH; c++) { bank.setAddress(c); bank.wrtData(getEdge(slot, c)); }
        bank.setAddress(0);
        return bank;
    }
    int getInfoIndexAt(int i) { return InfoIndex[i]; }
    int getTrueInfoBit(int i) { return TrueInfoBit[i]; }
    ldpc_decoder *engine(bool fixpoint) { ensure_decoder(); return fixpoint ? dec_pre : dec_gen; }
    /* one more decoder of the same code per additional GPU (multi-GPU frame loops of the drivers) */
    ldpc_decoder *engine_on(bool fixpoint, int device)
    {
        if (device == 0) return engine(fixpoint);
        ensure_decoder();
        const size_t slot = (size_t)device * 2 + (fixpoint ? 1 : 0);
        if (more_.size() <= slot) more_.resize(slot + 1, NULL);
        if (!more_[slot]) {
            ldpc_decoder_cfg cfg;
            ldpc_decoder_cfg_default(&cfg);
            cfg.max_iter = MAX_ITER; cfg.precheck = fixpoint ? 1 : 0; cfg.device = device;
            int err = LDPC_OK;
            more_[slot] = ldpc_decoder_create(code_, &cfg, &err);
            if (!more_[slot]) ldpc_facade::fail("FP_Decoder: cannot create the GPU decoder", err);
        }
        return more_[slot];
    }
    const ldpc_code *code() { return code_; }

private:
    FP_Decoder(const FP_Decoder &);
    FP_Decoder &operator=(const FP_Decoder &);

    void init_members()
    {
        code_ = NULL; dec_gen = NULL; dec_pre = NULL; n_ = NUM_VAR; m_ = NUM_CHK; dc_ = CHK_DEG;
        BitError = 0; last_syndrome_fail = 0;
        DecodedCodeword.assign(CWD_LENGTH, 0); TrueCodeword.assign(CWD_LENGTH, 0); Posteriori_fp.assign(CWD_LENGTH, 0);
        TrueInfoBit.assign(INFO_LENGTH + 8, 0); InfoIndex.assign(INFO_LENGTH, 0);
        EdgeRAM_fp.assign((size_t)CHK_DEG * NUM_CHK, 0);
    }
    void select_variant()
    {
        int err = LDPC_OK;
#if LDPC_CODE_VARIANT == 1 || LDPC_CODE_VARIANT == 2
        adopt(ldpc_code_array(P, NUM_CGRP, NULL, NUM_VGRP, NULL, 0, &err), err);
#elif LDPC_CODE_VARIANT == 3
        adopt(ldpc_code_array(P, NUM_CGRP, ldpc_facade::cut79_rows, NUM_VGRP, ldpc_facade::cut79_cols, 1, &err), err);
#else
        (void)err; /* 802.11: tables arrive with ReadH() */
#endif
    }
    void load(const char *path, int format)
    {
        int err = LDPC_OK;
        adopt(ldpc_code_load(path, format, &err), err);
    }
    void adopt(ldpc_code *c, int err)
    {
        if (!c) ldpc_facade::fail("FP_Decoder: cannot set up the code", err);
        release();
        code_ = c;
        int e = 0, dv = 0;
        ldpc_code_dims(code_, &n_, &m_, &e, &dc_, &dv);
        DecodedCodeword.assign(n_, 0); TrueCodeword.assign(n_, 0); Posteriori_fp.assign(n_, 0);
        EdgeRAM_fp.assign((size_t)dc_ * m_, 0);
    }
    void release()
    {
        if (dec_gen) ldpc_decoder_destroy(dec_gen);
        if (dec_pre) ldpc_decoder_destroy(dec_pre);
        for (size_t i = 0; i < more_.size(); i++) if (more_[i]) ldpc_decoder_destroy(more_[i]);
        more_.clear();
        if (code_) ldpc_code_free(code_);
        dec_gen = dec_pre = NULL; code_ = NULL;
    }
    void ensure_decoder()
    {
        if (dec_gen) return;
        if (!code_) throw std::runtime_error("FP_Decoder: no parity-check tables (call ReadH first)");
        ldpc_decoder_cfg cfg;
        ldpc_decoder_cfg_default(&cfg);
        cfg.max_iter = MAX_ITER;
        int err = LDPC_OK;
        cfg.precheck = 0;
        dec_gen = ldpc_decoder_create(code_, &cfg, &err);
        if (!dec_gen) ldpc_facade::fail("FP_Decoder: cannot create the GPU decoder", err);
        cfg.precheck = 1;
        dec_pre = ldpc_decoder_create(code_, &cfg, &err);
        if (!dec_pre) ldpc_facade::fail("FP_Decoder: cannot create the GPU decoder", err);
    }
    void unpack(const std::vector<uint32_t> &bits)
    {
        for (int v = 0; v < n_; v++) DecodedCodeword[v] = (bits[v >> 5] >> (v & 31)) & 1;
    }
    int run(const int *LLR, int fixpoint)
    {
        ensure_decoder();
        std::vector<uint32_t> bits((n_ + 31) / 32);
        int iters = 0;
        if (fixpoint) { post_tmp.resize(n_); edge_tmp.resize(EdgeRAM_fp.size()); }
        int *post = fixpoint ? &post_tmp[0] : &Posteriori_fp[0], *edge = fixpoint ? &edge_tmp[0] : &EdgeRAM_fp[0];
        int rc = ldpc_decode_batch(fixpoint ? dec_pre : dec_gen, LLR, 1, &iters, &bits[0], post, edge);
        if (rc != LDPC_OK) ldpc_facade::fail("FP_Decoder: decode failed", rc);
        unpack(bits); /* a pre-check hit leaves the channel hard decisions here, like hardDecision (:276) */
        if (fixpoint) {
            if (iters == 0) return 0; /* Posteriori_fp / EdgeRAM keep the previous frame's values (quirk Q6) */
            Posteriori_fp.swap(post_tmp);
            EdgeRAM_fp.swap(edge_tmp);
        }
        /* the kernel stops on the first passing syndrome, so only a frame that used every iteration can have
         * failed its last check; ask the engine (checkPost_fp, :619-630) */
        last_syndrome_fail = 0;
        if (iters >= MAX_ITER) {
            std::vector<int> keep(DecodedCodeword);
            last_syndrome_fail = syndrome_of(&Posteriori_fp[0]);
            DecodedCodeword.swap(keep);
        }
        return iters;
    }
    int syndrome_of(const int *values)
    {
        ensure_decoder();
        std::vector<uint32_t> bits((n_ + 31) / 32);
        int fail = 0;
        int rc = ldpc_hard_decision_batch(dec_gen, values, 1, &fail, &bits[0]);
        if (rc != LDPC_OK) ldpc_facade::fail("FP_Decoder: syndrome evaluation failed", rc);
        unpack(bits);
        return fail;
    }

    class ROM CodeROM;
    class ControlFSM FSM;
    ldpc_code *code_;
    ldpc_decoder *dec_gen, *dec_pre;
    std::vector<ldpc_decoder *> more_;
    int n_, m_, dc_;
    std::vector<int> DecodedCodeword, TrueCodeword, TrueInfoBit, InfoIndex, Posteriori_fp, EdgeRAM_fp;
    std::vector<int> post_tmp, edge_tmp;
    std::vector<double> Posteriori, EdgeRAM;
    int BitError;
    int last_syndrome_fail;
    static const int Constant = int((5.0 / 8.0) * (1 << FRAC_WIDTH));
};

/* Generator-equation encoder (reference :179-214, ArrayLDPC_Encoder.cpp:34-225). */
class FP_Encoder {
public:
    FP_Encoder(const char *Filename, int flag)
    {
        if (flag) cout << "reading file " << Filename << endl;
        int err = LDPC_OK;
        gen_ = ldpc_gen_load(Filename, &err);
        if (!gen_) {
            cout << "Exception opening/reading file " << Filename << endl; /* reference :145-149 then exits */
            ldpc_facade::fail("FP_Encoder", err);
        }
        int n = 0, rows = 0, k = 0;
        ldpc_gen_dims(gen_, &n, &rows, &k);
        Codeword.assign(n, 0);
        InfoIndex.assign(k, 0);
        ldpc_gen_indices(gen_, &InfoIndex[0], NULL);
        if (flag) cout << "encoder initialized" << endl;
    }
    ~FP_Encoder() { ldpc_gen_free(gen_); }
    int encode(char *in, int in_len)
    {
        std::vector<uint8_t> cw(Codeword.size());
        int rc = ldpc_gen_encode(gen_, in, in_len, &cw[0]);
        if (rc != LDPC_OK) ldpc_facade::fail("FP_Encoder::encode", rc);
        for (size_t i = 0; i < cw.size(); i++) Codeword[i] = cw[i];
        return 2209; /* "hard coded for now", reference :163,:224 */
    }
    /* The debug overload also packs the codeword into out (reference :228-324) without the console dump. */
    int encode(char *in, char *out, int in_len)
    {
        encode(in, in_len);
        for (size_t i = 0; i < Codeword.size(); i++) out[i / 8] = (char)(out[i / 8] ^ (Codeword[i] << (i % 8)));
        return (int)((Codeword.size() + 7) / 8);
    }
    int getCodeword(int addr) { return Codeword[addr]; }
    int getInfoIndex(int addr) { return InfoIndex[addr]; }
    const ldpc_gen *generator() { return gen_; }

private:
    FP_Encoder(const FP_Encoder &);
    FP_Encoder &operator=(const FP_Encoder &);
    ldpc_gen *gen_;
    std::vector<int> Codeword;
    std::vector<int32_t> InfoIndex;
};

#endif
