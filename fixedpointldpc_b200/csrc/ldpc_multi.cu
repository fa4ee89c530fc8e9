// ldpc_multi.cu -- one Monte-Carlo point on all GPUs of the box, behind the C ABI (ldpc_mc_group_*, ldpc_mc_run_multi).
//
// Replaces the frame loops of the reference's drivers (PerfTest.cpp:97-135, 276-311, 385-426, 491-512, 580-601) for a
// multi-GPU box.  Frames are independent and addressed by a global index g (the noise of frame g depends on (seed, g)
// only), so device r of R simulates the blocks  g in [(round*R + r)*B, (round*R + r + 1)*B)  with its own decoder handle
// and host thread, and the only communication is
//   * per round, ONE ncclAllReduce(sum) of the round's counters (frames, blocks in error, errors, iteration sum and the
//     iteration histogram: 40 x uint64 = 320 B) over NVLink, after which every thread evaluates the stopping rule on
//     the same numbers;
//   * in the final round only, the per-frame results of every device (host copies), so that the run is cut on exactly
//     the frame on which the reference's sequential `while(pckerror < 100)` / `while(Counter < MaxPckNum)` stops.
// The result therefore does not depend on R or B and, with LDPC_STREAM_REFERENCE, equals the reference's printout.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a process that already carries an NCCL (PyTorch) keeps that one.
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h>

#include "../../include/ldpc_capi.h"
#include "ldpc_code.hpp"

namespace ldpc {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // 1. the library LDPC_NCCL_LIB names (the Python binding points it at the NCCL PyTorch ships, so that a later
        //    `import torch` in the same process finds the NCCL it was built against under the shared soname);
        // 2. an NCCL the process has already loaded; 3. the system's.  Never RTLD_GLOBAL.
        const char *env = getenv("LDPC_NCCL_LIB");
        if (env && *env) api.handle = dlopen(env, RTLD_NOW | RTLD_LOCAL);
        if (!api.handle) api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            if (api.handle) break;
            api.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        }
        if (!api.handle) return;
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(dlsym(api.handle, "ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.handle, "ncclAllReduce"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
        if (!api.CommInitAll || !api.CommDestroy || !api.AllReduce || !api.GetErrorString) api.handle = nullptr;
    });
    return api.handle ? &api : nullptr;
}

constexpr int RED_WORDS = 40;  // 0 frames, 1 blocks in error, 2 errors, 3 iteration sum, 4 failed ranks, 8..39 iteration histogram
constexpr int HIST_BINS = 32;

// the round's counters in all-reduce layout, plus the histogram of the decoder's return values
__global__ void mc_round_counters(const int *iters, long long frames, const unsigned long long *mc4, int count_iterations,
                                  unsigned long long *out)
{
    __shared__ unsigned int hist[HIST_BINS];
    if (threadIdx.x < HIST_BINS) hist[threadIdx.x] = 0u;
    __syncthreads();
    for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < frames; f += (long long)gridDim.x * blockDim.x)
        atomicAdd(&hist[min(max(iters[f], 0), HIST_BINS - 1)], 1u);
    __syncthreads();
    if (threadIdx.x < HIST_BINS && hist[threadIdx.x]) atomicAdd(&out[8 + threadIdx.x], (unsigned long long)hist[threadIdx.x]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        out[0] = mc4[0];
        out[3] = mc4[3];
        if (!count_iterations) { out[1] = mc4[1]; out[2] = mc4[2]; }
    }
}

// ArrayLDPC_PerfTest / ArrayLDPC_TimeTrial count the decoder's return value as the block's errors (quirk Q9,
// PerfTest.cpp:507-511, 596-600): blocks in error = frames with a non-zero return value, errors = the iteration sum
__global__ void mc_round_iteration_rule(unsigned long long *out)
{
    out[1] = out[0] - out[8];
    out[2] = out[3];
}

struct Barrier {
    std::mutex m;
    std::condition_variable cv;
    int n, waiting = 0;
    unsigned long long generation = 0;
    explicit Barrier(int n_) : n(n_) {}
    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        const unsigned long long g = generation;
        if (++waiting == n) { waiting = 0; ++generation; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != g; });
    }
};

}  // namespace ldpc

struct ldpc_mc_group {
    std::vector<ldpc_decoder *> dec;
    std::vector<int> device;
    std::vector<ncclComm_t> comm;
    std::vector<cudaStream_t> stream;
    std::vector<unsigned long long *> d_mc4, d_red, d_sum, h_sum;  // per device
    std::vector<unsigned short *> d_ferr;
    std::vector<int *> d_iters;
    size_t cap = 0;
};

extern "C" int ldpc_decoder_device(const ldpc_decoder *dec);

namespace ldpc {

static int group_alloc(ldpc_mc_group &g, size_t frames)
{
    if (g.cap >= frames) return LDPC_OK;
    for (size_t r = 0; r < g.dec.size(); ++r) {
        if (cudaSetDevice(g.device[r]) != cudaSuccess) { set_error("cudaSetDevice failed"); return LDPC_ERR_CUDA; }
        cudaFree(g.d_ferr[r]); cudaFree(g.d_iters[r]);
        g.d_ferr[r] = nullptr; g.d_iters[r] = nullptr;
        if (cudaMalloc(&g.d_ferr[r], frames * sizeof(unsigned short)) != cudaSuccess ||
            cudaMalloc(&g.d_iters[r], frames * sizeof(int)) != cudaSuccess) {
            cudaGetLastError();
            set_error("Monte-Carlo group: per-frame buffers"); g.cap = 0; return LDPC_ERR_NOMEM;
        }
    }
    g.cap = frames;
    return LDPC_OK;
}

}  // namespace ldpc

extern "C" {

void ldpc_mc_group_destroy(ldpc_mc_group *g)
{
    if (!g) return;
    bool any_comm = false;
    for (ncclComm_t c : g->comm) any_comm |= c != nullptr;
    ldpc::NcclApi *api = any_comm ? ldpc::nccl_api() : nullptr;  // (a single-GPU group never touches NCCL)
    for (size_t r = 0; r < g->dec.size(); ++r) {
        cudaSetDevice(g->device[r]);
        if (r < g->comm.size() && g->comm[r] && api) api->CommDestroy(g->comm[r]);
        if (r < g->stream.size() && g->stream[r]) cudaStreamDestroy(g->stream[r]);
        if (r < g->d_mc4.size()) cudaFree(g->d_mc4[r]);
        if (r < g->d_red.size()) cudaFree(g->d_red[r]);
        if (r < g->d_sum.size()) cudaFree(g->d_sum[r]);
        if (r < g->h_sum.size() && g->h_sum[r]) cudaFreeHost(g->h_sum[r]);
        if (r < g->d_ferr.size()) cudaFree(g->d_ferr[r]);
        if (r < g->d_iters.size()) cudaFree(g->d_iters[r]);
    }
    delete g;
}

ldpc_mc_group *ldpc_mc_group_create(ldpc_decoder *const *decoders, int ndev, int *err)
{
    auto fail = [&](ldpc_mc_group *g, int st) { if (err) *err = st; ldpc_mc_group_destroy(g); return (ldpc_mc_group *)nullptr; };
    if (!decoders || ndev < 1) { ldpc::set_error("no decoders"); return fail(nullptr, LDPC_ERR_ARG); }
    ldpc_mc_group *g = new ldpc_mc_group;
    const size_t R = (size_t)ndev;
    g->dec.assign(decoders, decoders + ndev);
    g->device.resize(R); g->comm.assign(R, nullptr); g->stream.assign(R, nullptr);
    g->d_mc4.assign(R, nullptr); g->d_red.assign(R, nullptr); g->d_sum.assign(R, nullptr); g->h_sum.assign(R, nullptr);
    g->d_ferr.assign(R, nullptr); g->d_iters.assign(R, nullptr);
    for (size_t r = 0; r < R; ++r) {
        if (!decoders[r]) { ldpc::set_error("NULL decoder in the group"); return fail(g, LDPC_ERR_ARG); }
        g->device[r] = ldpc_decoder_device(decoders[r]);
        for (size_t q = 0; q < r; ++q)
            if (g->device[q] == g->device[r]) { ldpc::set_error("two decoders of the group share a device"); return fail(g, LDPC_ERR_ARG); }
    }
    for (size_t r = 0; r < R; ++r) {
        cudaError_t e = cudaSetDevice(g->device[r]);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->stream[r], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&g->d_mc4[r], 4 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMalloc(&g->d_red[r], ldpc::RED_WORDS * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMalloc(&g->d_sum[r], ldpc::RED_WORDS * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaHostAlloc(&g->h_sum[r], ldpc::RED_WORDS * sizeof(unsigned long long), cudaHostAllocDefault);
        if (e != cudaSuccess) { ldpc::set_error(std::string("Monte-Carlo group: ") + cudaGetErrorString(e)); return fail(g, LDPC_ERR_CUDA); }
    }
    if (R > 1) {
        ldpc::NcclApi *api = ldpc::nccl_api();
        if (!api) { ldpc::set_error("libnccl.so.2 not found: a group of more than one GPU needs NCCL"); return fail(g, LDPC_ERR_UNSUPPORTED); }
        ncclResult_t nr = api->CommInitAll(g->comm.data(), ndev, g->device.data());
        if (nr != ncclSuccess) { ldpc::set_error(std::string("ncclCommInitAll: ") + api->GetErrorString(nr)); return fail(g, LDPC_ERR_CUDA); }
    }
    if (err) *err = LDPC_OK;
    return g;
}

int ldpc_mc_group_size(const ldpc_mc_group *g) { return g ? (int)g->dec.size() : 0; }

int ldpc_mc_group_run(ldpc_mc_group *g, const ldpc_mc_cfg *cfg, const ldpc_mc_stop *stop, ldpc_mc_result *res)
{
    if (!g || !cfg || !stop || !res) { ldpc::set_error("NULL group / cfg / stop / result"); return LDPC_ERR_ARG; }
    if (stop->target_block_errors == 0 && stop->max_frames == 0) { ldpc::set_error("no stopping rule"); return LDPC_ERR_ARG; }
    const size_t R = g->dec.size();
    const size_t B = stop->frames_per_round ? stop->frames_per_round : ((size_t)1 << 17);
    std::memset(res, 0, sizeof *res);
    int rc = ldpc::group_alloc(*g, B);
    if (rc != LDPC_OK) return rc;
    ldpc::NcclApi *api = R > 1 ? ldpc::nccl_api() : nullptr;

    ldpc::Barrier barrier((int)R);
    std::vector<std::vector<unsigned short>> h_ferr(R);
    std::vector<std::vector<int>> h_iters(R);
    std::vector<int> status(R, LDPC_OK);
    std::vector<std::string> message(R);
    ldpc_mc_result total;
    std::memset(&total, 0, sizeof total);
    const auto t0 = std::chrono::steady_clock::now();

    auto worker = [&](size_t r) {
        auto cuda_ok = [&](cudaError_t e, const char *what) {
            if (e != cudaSuccess && status[r] == LDPC_OK) { status[r] = LDPC_ERR_CUDA; message[r] = std::string(what) + ": " + cudaGetErrorString(e); }
            return e == cudaSuccess;
        };
        cuda_ok(cudaSetDevice(g->device[r]), "cudaSetDevice");
        cudaStream_t st = g->stream[r];
        // every thread keeps the same running totals (they all see the same reduced numbers)
        unsigned long long frames = 0, block_errors = 0, errors = 0, iter_sum = 0, rounds = 0, hist[ldpc::HIST_BINS] = {0};
        for (unsigned long long round = 0;; ++round) {
            const unsigned long long lo = (round * R + r) * B;  // first frame of this device's block, relative to first_frame
            size_t cnt = B;
            if (stop->max_frames) cnt = lo >= stop->max_frames ? 0 : (size_t)std::min<unsigned long long>(B, stop->max_frames - lo);
            cuda_ok(cudaMemsetAsync(g->d_mc4[r], 0, 4 * sizeof(unsigned long long), st), "memset");
            cuda_ok(cudaMemsetAsync(g->d_red[r], 0, ldpc::RED_WORDS * sizeof(unsigned long long), st), "memset");
            if (cnt > 0 && status[r] == LDPC_OK) {
                ldpc_mc_cfg mine = *cfg;
                mine.first_frame = cfg->first_frame + lo;
                const int e = ldpc_mc_run_device(g->dec[r], &mine, cnt, g->d_ferr[r], g->d_iters[r],
                                                 reinterpret_cast<uint64_t *>(g->d_mc4[r]), st);
                if (e != LDPC_OK) { status[r] = e; message[r] = ldpc_last_error(); }
            }
            if (status[r] == LDPC_OK) {
                ldpc::mc_round_counters<<<64, 256, 0, st>>>(g->d_iters[r], (long long)cnt, g->d_mc4[r], stop->count_iterations, g->d_red[r]);
                if (stop->count_iterations) ldpc::mc_round_iteration_rule<<<1, 1, 0, st>>>(g->d_red[r]);
                cuda_ok(cudaGetLastError(), "counter kernels");
            } else {
                const unsigned long long one = 1;  // tell the other ranks through the all-reduce itself
                cudaMemcpyAsync(g->d_red[r] + 4, &one, sizeof one, cudaMemcpyHostToDevice, st);
            }
            unsigned long long *reduced = g->d_red[r];
            if (R > 1) {
                const ncclResult_t nr = api->AllReduce(g->d_red[r], g->d_sum[r], ldpc::RED_WORDS, ncclUint64, ncclSum, g->comm[r], st);
                if (nr != ncclSuccess && status[r] == LDPC_OK) { status[r] = LDPC_ERR_CUDA; message[r] = std::string("ncclAllReduce: ") + api->GetErrorString(nr); }
                reduced = g->d_sum[r];
            }
            cudaMemcpyAsync(g->h_sum[r], reduced, ldpc::RED_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
            cuda_ok(cudaStreamSynchronize(st), "round");
            const unsigned long long *s = g->h_sum[r];
            if (s[4] != 0 || status[r] != LDPC_OK) {  // some rank failed: everybody leaves after this round
                if (status[r] == LDPC_OK) status[r] = LDPC_ERR_CUDA;
                return;
            }
            ++rounds;
            const bool last = (stop->target_block_errors && block_errors + s[1] >= stop->target_block_errors) ||
                              (stop->max_frames && frames + s[0] >= stop->max_frames) || s[0] == 0;
            const bool log_all = stop->iters_out != nullptr;
            if (last || log_all) {
                h_iters[r].resize(cnt);
                if (cnt) cudaMemcpyAsync(h_iters[r].data(), g->d_iters[r], cnt * sizeof(int), cudaMemcpyDeviceToHost, st);
                if (last) {
                    h_ferr[r].resize(cnt);
                    if (cnt) cudaMemcpyAsync(h_ferr[r].data(), g->d_ferr[r], cnt * sizeof(unsigned short), cudaMemcpyDeviceToHost, st);
                }
                cuda_ok(cudaStreamSynchronize(st), "per-frame results");
                if (log_all && !last)
                    for (size_t i = 0; i < cnt && lo + i < stop->iters_cap; ++i) stop->iters_out[lo + i] = h_iters[r][i];
            }
            if (!last) {
                frames += s[0]; block_errors += s[1]; errors += s[2]; iter_sum += s[3];
                for (int b = 0; b < ldpc::HIST_BINS; ++b) hist[b] += s[8 + b];
                continue;
            }
            // final round: cut on the exact frame, in global frame order (device 0's block first)
            barrier.wait();
            if (r == 0) {
                bool done = false;
                for (size_t q = 0; q < R && !done; ++q) {
                    const unsigned long long qlo = (round * R + q) * B;
                    for (size_t i = 0; i < h_iters[q].size() && !done; ++i) {
                        const unsigned long long blk = stop->count_iterations ? (unsigned long long)std::max(h_iters[q][i], 0) : h_ferr[q][i];
                        ++frames; errors += blk; iter_sum += (unsigned long long)std::max(h_iters[q][i], 0);
                        if (blk) ++block_errors;
                        hist[std::min(std::max(h_iters[q][i], 0), ldpc::HIST_BINS - 1)]++;
                        if (stop->iters_out && qlo + i < stop->iters_cap) stop->iters_out[qlo + i] = h_iters[q][i];
                        if (stop->target_block_errors && block_errors >= stop->target_block_errors) { done = true; total.reached = 1; }
                        if (stop->max_frames && frames >= stop->max_frames) done = true;
                    }
                }
                total.frames = frames; total.block_errors = block_errors; total.errors = errors; total.iter_sum = iter_sum;
                total.rounds = rounds;
                for (int b = 0; b < ldpc::HIST_BINS; ++b) total.iter_hist[b] = hist[b];
            }
            barrier.wait();
            return;
        }
    };

    if (R == 1) {
        worker(0);
    } else {
        std::vector<std::thread> threads;
        for (size_t r = 0; r < R; ++r) threads.emplace_back(worker, r);
        for (auto &t : threads) t.join();
    }
    for (size_t r = 0; r < R; ++r)
        if (status[r] != LDPC_OK) {
            if (!message[r].empty()) ldpc::set_error("device " + std::to_string(g->device[r]) + ": " + message[r]);
            for (size_t q = 0; q < R; ++q)
                if (!message[q].empty()) { ldpc::set_error("device " + std::to_string(g->device[q]) + ": " + message[q]); break; }
            return status[r];
        }
    *res = total;
    res->devices = (int)R;
    res->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return LDPC_OK;
}

int ldpc_mc_run_multi(ldpc_decoder *const *decoders, int ndev, const ldpc_mc_cfg *cfg, const ldpc_mc_stop *stop, ldpc_mc_result *result)
{
    int err = LDPC_OK;
    ldpc_mc_group *g = ldpc_mc_group_create(decoders, ndev, &err);
    if (!g) return err;
    const int rc = ldpc_mc_group_run(g, cfg, stop, result);
    ldpc_mc_group_destroy(g);
    return rc;
}

}  // extern "C"
