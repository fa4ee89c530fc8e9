// facade_check.cpp -- drives the reference-compatible FP_Decoder class (include/ArrayLDPCMacro.h) exactly like
// the reference's drivers do, one frame at a time, and compares every observable with vectors dumped from the
// reference.  Built and run by tests/test_gpu_facade.py.   usage: facade_check <dir> <general|fixpoint>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ArrayLDPCMacro.h"

template <class T> static std::vector<T> slurp(const std::string &path)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<T> v(sz / sizeof(T));
    if (fread(v.data(), sizeof(T), v.size(), f) != v.size()) exit(2);
    fclose(f);
    return v;
}

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    const std::string dir = argv[1];
    const bool fixpoint = !strcmp(argv[2], "fixpoint");
    if (!strcmp(argv[2], "latency")) {
        // one frame per call, like the reference's drivers: wall time per call of the facade's decode entry points
        std::vector<int> llr = slurp<int>(dir + "/llr.bin"), it0 = slurp<int>(dir + "/iters.bin");
        FP_Decoder D((dir + "/H.txt").c_str());
        const int frames = (int)it0.size(), reps = 50;
        D.decode_general_fp(&llr[0]);  // device set-up
        for (int mode = 0; mode < 2; mode++) {
            long total_it = 0;
            auto t0 = std::chrono::steady_clock::now();
            for (int r = 0; r < reps; r++)
                for (int f = 0; f < frames; f++) {
                    if (mode) { D.setState(PCV); total_it += D.decode_fixpoint(&llr[(size_t)f * CWD_LENGTH]); }
                    else total_it += D.decode_general_fp(&llr[(size_t)f * CWD_LENGTH]);
                }
            double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / (reps * frames);
            printf("%s: %.1f us per call (%d calls, %.2f iterations on average)\n", mode ? "decode_fixpoint" : "decode_general_fp", us,
                   reps * frames, double(total_it) / (reps * frames));
        }
        return 0;
    }
    if (!strcmp(argv[2], "double")) {
        // FP_Decoder::decode_general(const double *) (ArrayLDPC_Decoder.cpp:735-933) through the facade, one frame at a
        // time: return value and DecodedCodeword equal to the reference's, posteriors within the stated tolerance
        std::vector<double> llr64 = slurp<double>(dir + "/llr64.bin"), post64 = slurp<double>(dir + "/post64.bin");
        std::vector<int> it64 = slurp<int>(dir + "/iters.bin");
        std::vector<unsigned char> b64 = slurp<unsigned char>(dir + "/bits.bin");
        FP_Decoder D((dir + "/H.txt").c_str());
        int wrong = 0;
        for (size_t f = 0; f < it64.size(); f++) {
            int it = D.decode_general(&llr64[f * CWD_LENGTH]);
            if (it != it64[f]) { printf("frame %zu: iters %d want %d\n", f, it, it64[f]); wrong++; continue; }
            const double rel = it < MAX_ITER ? 1e-9 : 1e-6;
            for (int v = 0; v < CWD_LENGTH; v++) {
                if (D.getDecodedBit(v) != b64[f * CWD_LENGTH + v]) { printf("frame %zu: bit %d\n", f, v); wrong++; break; }
                const double want = post64[f * CWD_LENGTH + v], tol = rel * (fabs(want) > 1 ? fabs(want) : 1);
                if (fabs(D.getPost(v) - want) > tol) { printf("frame %zu: post %d\n", f, v); wrong++; break; }
            }
        }
        if (D.sxor(0.3, -2.0) >= 0 || fabs(D.sxor(5.0, 5.0) - (5.0 + log(1 + exp(-10.0)) - log(2.0))) > 1e-15) { printf("sxor(double)\n"); wrong++; }
        printf("frames %zu zero_hits 0 mismatches %d\n", it64.size(), wrong);
        return wrong ? 1 : 0;
    }
    std::vector<int> llr = slurp<int>(dir + "/llr.bin"), iters = slurp<int>(dir + "/iters.bin"),
                     post = slurp<int>(dir + "/post.bin"), edge = slurp<int>(dir + "/edge.bin"),
                     cdeg = slurp<int>(dir + "/cdeg.bin");
    std::vector<unsigned char> bits = slurp<unsigned char>(dir + "/bits.bin");
    const int n = CWD_LENGTH, m = NUM_CHK, frames = (int)iters.size();
    FP_Decoder Decoder((dir + "/H.txt").c_str());
    int bad = 0, zero_hits = 0;
    std::vector<int> prev_post(n, 0);
    for (int f = 0; f < frames; f++) {
        const int *x = &llr[(size_t)f * n];
        int it;
        if (fixpoint) {
            Decoder.setState(PCV);  // PerfTest.cpp:121,180,298
            it = Decoder.decode_fixpoint(x);
        } else {
            it = Decoder.decode_general_fp(x);
        }
        if (it != iters[f]) { printf("frame %d: iters %d want %d\n", f, it, iters[f]); bad++; continue; }
        for (int v = 0; v < n; v++)
            if (Decoder.getDecodedBit(v) != bits[(size_t)f * n + v]) { printf("frame %d: bit %d\n", f, v); bad++; break; }
        if (it == 0) {
            // pre-check hit: Posteriori_fp keeps the previous frame's values (quirk Q6) and the FSM stays in PCV
            zero_hits++;
            for (int v = 0; v < n; v++)
                if (Decoder.getPost_fp(v) != prev_post[v]) { printf("frame %d: stale post %d\n", f, v); bad++; break; }
            if (Decoder.getState() != PCV) { printf("frame %d: state %d after pre-check hit\n", f, Decoder.getState()); bad++; }
            continue;
        }
        for (int v = 0; v < n; v++) {
            if (Decoder.getPost_fp(v) != post[(size_t)f * n + v]) { printf("frame %d: post %d\n", f, v); bad++; break; }
            prev_post[v] = Decoder.getPost_fp(v);
        }
        for (int c = 0; c < m && bad < 10; c++)
            for (int k = 0; k < cdeg[c]; k++)
                if (Decoder.getEdge(k, c) != edge[((size_t)f * CHK_DEG + k) * m + c]) { printf("frame %d: edge %d %d\n", f, k, c); bad++; break; }
        if (fixpoint) {
            const int want_state = (it == MAX_ITER && Decoder.checkPost_fp()) ? C2V : IDLE;
            if (Decoder.getState() != want_state) { printf("frame %d: state %d want %d\n", f, Decoder.getState(), want_state); bad++; }
        }
        // the syndrome helpers on the engine
        if (Decoder.hardDecision(x) != (fixpoint ? 1 : Decoder.hardDecision(x))) bad++;
        if (it < MAX_ITER && Decoder.checkPost_fp_general() != 0) { printf("frame %d: converged frame fails checkPost\n", f); bad++; }
    }
    // scalar helper parity with the pinned spot values
    if (Decoder.sxor(5, -3) != -1 || Decoder.sxor(255, 1) != 11 || Decoder.sxor(200, 60) != 69 || Decoder.sxor(4770, -4000) != -3990 ||
        Decoder.sxor(0, 7) != 0 || Decoder.sxor(-20, 20) != -10)
        { printf("sxor helper\n"); bad++; }
    if (fixpoint) {
        // FSM gate: without setState(PCV) after a converged frame the reference's loop does not run (returns 0)
        Decoder.setState(IDLE);
        int it = Decoder.decode_fixpoint(&llr[0]);
        if (it != 0) { printf("IDLE gate returned %d\n", it); bad++; }
    }
    printf("frames %d zero_hits %d mismatches %d\n", frames, zero_hits, bad);
    return bad ? 1 : 0;
}
