// PerfTest.cpp (facade) -- the reference's Monte-Carlo drivers (PerfTest.cpp:23-607) on the B200 engine.
//
// Every driver of the reference is one loop: draw a frame from the channel, decode it, update counters, stop
// on a rule.  Here the loop body is a batched GPU launch (ldpc_mc_run: channel + decode + calculateBER fused)
// that reports per-frame results in frame order, and the stopping rule is applied to that ordered list, so
// the counters -- and therefore the printed lines -- are the ones the sequential reference produces.  The
// noise is the reference's own stream (LDPC_STREAM_REFERENCE), continued across drivers like the file-static
// state of rngs.cpp.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "ArrayLDPCMacro.h"
#include "ArrayLDPC.h"
#include "PerfTest.h"

using std::cin;
using std::cout;
using std::endl;
using std::ofstream;

namespace {

// rngs.cpp:45-47: one process-wide stream, never re-seeded by any driver
unsigned long long g_stream_state = 123456789ULL;
const unsigned long long LEHMER_M = 2147483647ULL, LEHMER_A = 48271ULL;

unsigned long long mulmod(unsigned long long a, unsigned long long b) { return (a * b) % LEHMER_M; }
unsigned long long powmod(unsigned long long b, unsigned long long e)
{
    unsigned long long r = 1;
    for (b %= LEHMER_M; e; e >>= 1, b = mulmod(b, b))
        if (e & 1) r = mulmod(r, b);
    return r;
}
void consume_uniforms(unsigned long long count) { g_stream_state = mulmod(g_stream_state, powmod(LEHMER_A, count)); }

// PerfTest.cpp:33 and :221-224 (the continuation lines of the literal carry their leading tabs)
const char kWifiMessage[122] = "OMG  how long   dd   should this string be to make it 243";
const char kArrayMessage[248] =
    "OMG how long should this string be to make it 248, just imagine that. "
    "\t\t\t\t\t\t\t\t   I guess it's still not long enough. Let's see. This is a testing string "
    "\t\t\t\t\t\t\t\t\tfor a lot of characters so that we have some random bit stream that's"
    "\t\t\t\t\t\t\t\t\tcorrect";

struct Point {
    double biterror, pckerror;
    long Counter;
    unsigned long long iter_sum;
    unsigned long long iter_hist[32];  // frames per decoder return value
    int devices;
    double seconds;
};

enum Rule { COUNT_INFO_BIT_ERRORS, COUNT_ITERATIONS };

// GPUs the frame loops run on: all visible ones, or the first LDPC_GPUS of them
int gpu_count()
{
    int n = ldpc_device_count();
    if (n < 1) ldpc_facade::fail("no CUDA device", n < 0 ? n : LDPC_ERR_NO_DEVICE);
    const char *env = getenv("LDPC_GPUS");
    if (env && atoi(env) > 0 && atoi(env) < n) n = atoi(env);
    return n;
}

// The frame loop shared by all drivers, on all GPUs of the box (ldpc_mc_group_run: device r of R simulates the frames
// [(round*R + r)*B, +B) of every round, the counters are all-reduced over NCCL once per round, and the run is cut on
// exactly the frame on which the reference's sequential loop stops).  stop_errors: `while(pckerror < stop_errors)`;
// max_frames: `while(Counter < MaxPckNum)`.  Returns the reference's three counters.
Point simulate(FP_Decoder &Decoder, bool fixpoint, double snr, double sigma, const std::vector<uint8_t> *codeword,
               const std::vector<int32_t> *info_index, const std::vector<int32_t> *pins, int pin_value, Rule rule,
               long stop_errors, long max_frames, std::vector<int> *iters_log)
{
    const int R = gpu_count();
    std::vector<ldpc_decoder *> dec(R);
    for (int r = 0; r < R; ++r) dec[r] = Decoder.engine_on(fixpoint, r);
    ldpc_mc_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.snr = snr; cfg.sigma = sigma;
    cfg.stream = LDPC_STREAM_REFERENCE; cfg.seed = g_stream_state;
    // LDPC_STREAM=philox: the counter-based stream (float Box-Muller, several times cheaper per frame than the
    // reference's double-precision inverse normal) for long waterfall points; the counters then differ from the
    // reference's by Monte-Carlo noise only
    const char *stream_env = getenv("LDPC_STREAM");
    if (stream_env && !strcmp(stream_env, "philox")) cfg.stream = LDPC_STREAM_PHILOX;
    cfg.codeword = codeword ? &(*codeword)[0] : NULL;
    cfg.info_index = info_index ? &(*info_index)[0] : NULL;
    cfg.info_count = info_index ? (int)info_index->size() : 0;
    if (pins && !pins->empty()) { cfg.pin_index = &(*pins)[0]; cfg.pin_count = (int)pins->size(); cfg.pin_value = pin_value; }

    ldpc_mc_stop stop;
    memset(&stop, 0, sizeof stop);
    stop.target_block_errors = stop_errors > 0 ? (uint64_t)stop_errors : 0;
    stop.max_frames = max_frames > 0 ? (uint64_t)max_frames : 0;
    stop.count_iterations = rule == COUNT_ITERATIONS;
    const char *env = getenv("LDPC_MC_ROUND");
    stop.frames_per_round = env && atol(env) > 0 ? (size_t)atol(env) : (size_t)1 << 15;
    if (max_frames > 0 && (size_t)max_frames < stop.frames_per_round * R)
        stop.frames_per_round = ((size_t)max_frames + R - 1) / R;
    std::vector<int32_t> log;
    if (iters_log) {
        log.assign((size_t)1 << 22, 0);
        stop.iters_out = &log[0];
        stop.iters_cap = log.size();
    }
    ldpc_mc_result res;
    // NCCL writes its banner / debug lines to stdout unless told otherwise; stdout is the reference's transcript
    if (R > 1) setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
    int rc = ldpc_mc_run_multi(&dec[0], R, &cfg, &stop, &res);
    if (rc != LDPC_OK) ldpc_facade::fail("simulation", rc);
    Point pt;
    memset(&pt, 0, sizeof pt);
    pt.biterror = (double)res.errors; pt.pckerror = (double)res.block_errors; pt.Counter = (long)res.frames;
    pt.iter_sum = res.iter_sum; pt.devices = res.devices; pt.seconds = res.seconds;
    for (int b = 0; b < 32; ++b) pt.iter_hist[b] = res.iter_hist[b];
    if (iters_log) iters_log->assign(log.begin(), log.begin() + std::min<size_t>(log.size(), (size_t)res.frames));
    consume_uniforms((unsigned long long)pt.Counter * CWD_LENGTH);  // one uniform per transmitted bit (rvgs.cpp:169)
    return pt;
}

void print_point(const Point &p)
{
    cout << p.biterror << " " << p.pckerror << " " << p.Counter << endl
         << " FER: " << p.pckerror / p.Counter << " BER: " << p.biterror / p.Counter / CWD_LENGTH << endl;
}

void touch_outputs(const char *Filename)
{
    // the reference opens (truncates) Filename and Filename_log.txt and writes nothing (PerfTest.cpp:461-480)
    ofstream FilePtr(Filename);
    if (!FilePtr) { std::cerr << "failed to open " << Filename << endl; exit(0); }
    FilePtr.close();
    std::string log = std::string(Filename) + "_log.txt";  // char[40] + sprintf in the reference (quirk Q17)
    ofstream LogPtr(log.c_str());
    if (!LogPtr) { std::cerr << "failed to open " << log << endl; exit(0); }
    LogPtr.close();
}

void encode_message(FP_Encoder &Encoder, const char *msg, int len, std::vector<uint8_t> &cw, std::vector<int32_t> &idx)
{
    std::vector<char> buf(msg, msg + len);
    Encoder.encode(&buf[0], len);
    cw.resize(CWD_LENGTH);
    for (int i = 0; i < CWD_LENGTH; i++) cw[i] = (uint8_t)Encoder.getCodeword(i);
    idx.resize(INFO_LENGTH);
    for (int i = 0; i < INFO_LENGTH; i++) idx[i] = Encoder.getInfoIndex(i);
}

}  // namespace

void LDPC_PutSeed(long x) { g_stream_state = (unsigned long long)x; }
long LDPC_GetSeed() { return (long)g_stream_state; }

void noMoreMemory()
{
    std::cerr << "Unable to satisfy request for memory\n";
    abort();
}

// PerfTest.cpp:23-140
int ArrayLDPC_Debug_Wifi()
{
    double EbN0_dB = 3;
    class FP_Decoder Decoder;
    class FP_Encoder Encoder("H_802.11_IndZerog.txt", 0);
    cout << "EbNo in dB? ";
    cin >> EbN0_dB;
    double snr = 2 * pow(10.0, EbN0_dB / 10) * 0.5;
    double sigma = sqrt(1 / snr);
    cout << "SNR is " << 10 * log10(snr) << " dB" << endl;
    std::vector<uint8_t> cw;
    std::vector<int32_t> idx;
    encode_message(Encoder, kWifiMessage, 122, cw, idx);
    Decoder.ReadH();
    print_point(simulate(Decoder, false, snr, sigma, &cw, &idx, NULL, 0, COUNT_INFO_BIT_ERRORS, 100, 0, NULL));
    return 0;
}

// PerfTest.cpp:217-316
int ArrayLDPC_Debug()
{
    class FP_Decoder Decoder;
    class FP_Encoder Encoder("G_array_forward.txt", 0);
    double EbN0_dB = 4.5;
    double snr = 2 * pow(10.0, EbN0_dB / 10) * Decoder.getRate();
    double sigma = sqrt(1 / snr);
    cout << "SNR is " << 10 * log10(snr) << " dB" << endl;
    std::vector<uint8_t> cw;
    std::vector<int32_t> idx;
    encode_message(Encoder, kArrayMessage, 248, cw, idx);
    print_point(simulate(Decoder, true, snr, sigma, &cw, &idx, NULL, 0, COUNT_INFO_BIT_ERRORS, 100, 0, NULL));
    return 0;
}

// PerfTest.cpp:318-431
int ArrayLDPC_Debug_Shorten(int short_len)
{
    class FP_Decoder Decoder;
    class FP_Encoder Encoder("G_array_forward.txt", 0);
    double EbN0_dB = 4.5;
    double snr = 2 * pow(10.0, EbN0_dB / 10) * (1978.0 - 976.0) / 2209.0;  // :355 overrides the getRate() line
    double sigma = sqrt(1 / snr);
    cout << "SNR is " << 10 * log10(snr) << " dB" << endl;
    char InfoStream[248];
    memcpy(InfoStream, kArrayMessage, 248);
    for (int i = 0; i < short_len && i < 248; i++) InfoStream[i] = 0;  // zeroes BYTES (quirk Q18)
    std::vector<uint8_t> cw;
    std::vector<int32_t> idx;
    encode_message(Encoder, InfoStream, 248, cw, idx);
    std::vector<int32_t> pins(idx.begin(), idx.begin() + short_len);  // ... but pins short_len BIT positions
    std::vector<int> log;
    Point p = simulate(Decoder, true, snr, sigma, &cw, &idx, &pins, 7 * (1 << FRAC_WIDTH), COUNT_INFO_BIT_ERRORS, 100, 0, &log);
    for (size_t i = 0; i < log.size(); i++) cout << log[i] << ", ";  // :419 prints every return value
    print_point(p);
    return 0;
}

// PerfTest.cpp:433-517: only db_start is used; the "errors" are iteration counts (quirk Q9)
int ArrayLDPC_PerfTest(double db_start, double db_end, double db_step, char *Filename)
{
    (void)db_end; (void)db_step;
    class FP_Decoder Decoder;
    touch_outputs(Filename);
    double EbN0 = 2 * pow(10.0, db_start / 10) * Decoder.getRate();
    double sigma = sqrt(1 / EbN0);
    print_point(simulate(Decoder, true, EbN0, sigma, NULL, NULL, NULL, 0, COUNT_ITERATIONS, 100, 0, NULL));
    return 0;
}

// PerfTest.cpp:520-607
int ArrayLDPC_TimeTrial(double db, int MaxPckNum, char *Filename)
{
    class FP_Decoder Decoder;
    touch_outputs(Filename);
    double EbN0 = 2 * pow(10.0, db / 10) * Decoder.getRate();
    double sigma = sqrt(1 / EbN0);
    print_point(simulate(Decoder, true, EbN0, sigma, NULL, NULL, NULL, 0, COUNT_ITERATIONS, 0, MaxPckNum, NULL));
    return 0;
}

// PerfTest.cpp:148-192: 100 pre-generated all-zero-codeword frames, MaxPacket timed decodes of frame i%100
int DecodeTrial(double EbN0_dB, int MaxPacket)
{
    class FP_Decoder Decoder;
    double snr = 2 * pow(10.0, EbN0_dB / 10) * Decoder.getRate();
    double sigma = sqrt(1 / snr);
    cout << "equivalent SNR is: " << 10 * log10(snr) << endl;
    ldpc_mc_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.snr = snr; cfg.sigma = sigma; cfg.stream = LDPC_STREAM_REFERENCE; cfg.seed = g_stream_state;
    std::vector<int> pool((size_t)100 * CWD_LENGTH);
    int rc = ldpc_mc_channel(Decoder.engine(true), &cfg, 100, &pool[0]);
    if (rc != LDPC_OK) ldpc_facade::fail("DecodeTrial channel", rc);
    consume_uniforms(100ULL * CWD_LENGTH);
    const size_t chunk = 25600;  // whole multiples of the 100-frame pool
    std::vector<int> llr(chunk * CWD_LENGTH), iters(chunk);
    for (size_t i = 0; i < chunk; i++) memcpy(&llr[i * CWD_LENGTH], &pool[(i % 100) * CWD_LENGTH], sizeof(int) * CWD_LENGTH);
    Decoder.decode_batch(&llr[0], 100, &iters[0], NULL, true);  // first call pays for the device set-up
    auto t0 = std::chrono::steady_clock::now();
    for (long left = MaxPacket; left > 0; left -= (long)chunk) {
        size_t now = left < (long)chunk ? (size_t)left : chunk;
        rc = Decoder.decode_batch(&llr[0], now, &iters[0], NULL, true);
        if (rc != LDPC_OK) ldpc_facade::fail("DecodeTrial decode", rc);
    }
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    cout << sec << "  seconds" << endl;
    cout << 2209.0 * MaxPacket / sec << " bits per second for decoder" << endl;
    return 0;
}

// PerfTest.cpp:193-215 (the encoder is host code in the reference and stays host code here)
int EncodeTrial(char *info, int MaxPacket)
{
    class FP_Encoder Encoder("G_array_forward.txt", 0);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < MaxPacket; i++) Encoder.encode(info, 248);
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    cout << sec << "  seconds" << endl;
    cout << 2209.0 * MaxPacket / sec << " bits per second for encoder" << endl;
    return 0;
}

// The sweep the reference's signature promises (db_start, db_end, db_step) but never runs, with the CSV it
// opens but never writes.  All-zero codeword, every position counted.
int ArrayLDPC_Sweep(double db_start, double db_end, double db_step, const char *Filename, int frame_errors, int short_len)
{
    class FP_Decoder Decoder;
#if LDPC_CODE_VARIANT == 0
    Decoder.ReadH();
    const double rate = 0.5;
    const bool fixpoint = false;
#elif LDPC_CODE_VARIANT == 3
    const double rate = double(INFO_LENGTH) / CWD_LENGTH;
    const bool fixpoint = false;
#else
    const double rate = Decoder.getRate();
    const bool fixpoint = true;
#endif
    // Shortening (BASELINE config 3; the rule of ArrayLDPC_Debug_Shorten, PerfTest.cpp:355, 410-414, for any code): the
    // information positions come from a generator derived from H (ldpc_gen_from_code: the reference ships no G for the
    // cut code), the first short_len of them are known zeros pinned to LLR 7 * 2^FRAC_WIDTH, the channel rate counts
    // the remaining ones, and calculateBER looks at the information positions only.
    std::vector<int32_t> info, pins;
    double sweep_rate = rate;
    if (short_len > 0) {
        int err = LDPC_OK, n = 0, rows = 0, k = 0;
        ldpc_gen *gen = ldpc_gen_from_code(Decoder.code(), NULL, 0, &err);
        if (!gen) ldpc_facade::fail("generator from H", err);
        ldpc_gen_dims(gen, &n, &rows, &k);
        info.resize(k);
        ldpc_gen_indices(gen, &info[0], NULL);
        ldpc_gen_free(gen);
        if (short_len > k) short_len = k;
        pins.assign(info.begin(), info.begin() + short_len);
        sweep_rate = double(k - short_len) / n;
    }
    ofstream csv(Filename, std::ios::app);
    if (!csv) { std::cerr << "failed to open " << Filename << endl; return 1; }
    csv << "EbN0_dB,frames,frame_errors,bit_errors,FER,BER,avg_iters" << endl;
    // the second file the reference opens and leaves empty (PerfTest.cpp:472-480): one line per point with the
    // histogram of the decoder's return values (frames per iteration count 0..MAX_ITER), GPUs used and wall time
    ofstream log((std::string(Filename) + "_log.txt").c_str(), std::ios::app);
    for (double db = db_start; db <= db_end + 1e-9; db += (db_step > 0 ? db_step : 1.0)) {
        double snr = 2 * pow(10.0, db / 10) * sweep_rate, sigma = sqrt(1 / snr);
        Point p = simulate(Decoder, fixpoint, snr, sigma, NULL, short_len > 0 ? &info : NULL, short_len > 0 ? &pins : NULL,
                           7 * (1 << FRAC_WIDTH), COUNT_INFO_BIT_ERRORS, frame_errors, 0, NULL);
        csv << db << "," << p.Counter << "," << p.pckerror << "," << p.biterror << "," << p.pckerror / p.Counter << ","
            << p.biterror / p.Counter / CWD_LENGTH << "," << double(p.iter_sum) / p.Counter << endl;
        if (log) {
            log << "EbN0_dB " << db << " frames " << p.Counter << " gpus " << p.devices << " seconds " << p.seconds << " iterations";
            for (int b = 0; b <= MAX_ITER; ++b) log << " " << p.iter_hist[b];
            log << endl;
        }
        cout << "Eb/N0 " << db << " dB: ";
        print_point(p);
    }
    return 0;
}
