// ldpc_encode.cu -- batched GF(2) encoder: parity_i = popcount(G_i & info) & 1 on bit-packed rows.
//
// Reference: FP_Encoder::encode (ArrayLDPC_Encoder.cpp:160-225) evaluates every parity equation of the
// Format B file as an XOR over its information members, one message at a time on the host.  Here the
// equations are packed once into a dense bit matrix G[row][k] over the k information positions (in
// InfoIndex order, which is also the bit order of the message bytes: LSB first, :171-183), stored
// word-major so that consecutive threads (rows) read consecutive addresses, and one CTA encodes one frame:
//   1. the message words go to shared memory,
//   2. thread r computes parity r = XOR_w popc(G[w][r] & info[w]) & 1,
//   3. thread v assembles codeword bit v (information bit or parity bit) and a warp ballot packs 32 of them.
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/ldpc_capi.h"
#include "ldpc_code.hpp"

namespace ldpc {

struct EncTables {
    int device = -1;
    uint32_t *d_g = nullptr;   // [kw][rows]
    int *d_src = nullptr;      // [n]  >= 0: information bit index, < 0: ~parity row
    cudaStream_t stream = nullptr;
    uint32_t *d_info = nullptr, *d_cw = nullptr;
    size_t cap = 0;
};

__global__ void encode_kernel(const uint32_t *__restrict__ g, const int *__restrict__ src, int n, int rows, int kw,
                              const uint32_t *__restrict__ info, long long frames, uint32_t *__restrict__ cw, int nw32)
{
    extern __shared__ uint32_t sm[];
    uint32_t *msg = sm;              // [kw]
    uint32_t *par = sm + kw;         // [ceil(rows/32)]
    const int prw = (rows + 31) / 32;
    for (long long f = blockIdx.x; f < frames; f += gridDim.x) {
        for (int w = threadIdx.x; w < kw; w += blockDim.x) msg[w] = info[(size_t)f * kw + w];
        __syncthreads();
        for (int r0 = 0; r0 < rows; r0 += blockDim.x) {
            const int r = r0 + threadIdx.x;
            uint32_t acc = 0;
            if (r < rows)
                for (int w = 0; w < kw; ++w) acc ^= g[(size_t)w * rows + r] & msg[w];
            const uint32_t word = __ballot_sync(0xffffffffu, __popc(acc) & 1);
            if ((threadIdx.x & 31) == 0 && r < rows) par[r >> 5] = word;
        }
        __syncthreads();
        for (int v0 = 0; v0 < n; v0 += blockDim.x) {
            const int v = v0 + threadIdx.x;
            uint32_t bit = 0;
            if (v < n) {
                const int s = src[v];
                bit = s >= 0 ? (msg[s >> 5] >> (s & 31)) & 1u : (par[(~s) >> 5] >> ((~s) & 31)) & 1u;
            }
            const uint32_t word = __ballot_sync(0xffffffffu, bit);
            if ((threadIdx.x & 31) == 0 && v < n) cw[(size_t)f * nw32 + (v >> 5)] = word;
        }
        __syncthreads();
        (void)prw;
    }
}

static void enc_release(EncTables &t)
{
    if (t.device >= 0) cudaSetDevice(t.device);
    cudaFree(t.d_g); cudaFree(t.d_src); cudaFree(t.d_info); cudaFree(t.d_cw);
    if (t.stream) cudaStreamDestroy(t.stream);
    t = EncTables();
}

// Makes `device` current and (re)builds the device tables there.  A generator lives on one device at a time: a call
// that names another device first releases everything the previous one holds (tables, staging, stream).
static int enc_prepare(const ldpc_gen &g, int device, EncTables &t)
{
    if (t.d_g && t.device == device) {
        if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return LDPC_ERR_CUDA; }
        return LDPC_OK;
    }
    if (t.d_g || t.stream) enc_release(t);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); set_error("no CUDA device visible"); return LDPC_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) { set_error("device ordinal out of range"); return LDPC_ERR_ARG; }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return LDPC_ERR_CUDA; }
    const int k = g.n - g.rows, kw = (k + 31) / 32;
    std::vector<int> pos(g.n, -1);
    for (int j = 0; j < k; ++j) pos[g.info_index[j]] = j;
    std::vector<uint32_t> dense((size_t)kw * g.rows, 0u);
    for (int r = 0; r < g.rows; ++r)
        for (int v : g.eq[r])
            if (g.flag[v] == 0) dense[(size_t)(pos[v] >> 5) * g.rows + r] ^= 1u << (pos[v] & 31);
    std::vector<int> src(g.n);
    for (int j = 0; j < k; ++j) src[g.info_index[j]] = j;
    for (int r = 0; r < g.rows; ++r) src[g.parity_index[r]] = ~r;
    cudaError_t e = cudaMalloc(&t.d_g, dense.size() * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&t.d_src, src.size() * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(t.d_g, dense.data(), dense.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t.d_src, src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error(std::string("encoder tables: ") + cudaGetErrorString(e)); return LDPC_ERR_CUDA; }
    t.device = device;
    return LDPC_OK;
}

static int enc_launch(const ldpc_gen &g, EncTables &t, const uint32_t *d_info, size_t frames, uint32_t *d_cw, cudaStream_t st)
{
    const int k = g.n - g.rows, kw = (k + 31) / 32, nw32 = (g.n + 31) / 32;
    const int smem = (kw + (g.rows + 31) / 32) * (int)sizeof(uint32_t);
    const unsigned grid = (unsigned)(frames < 148 * 8 ? frames : 148 * 8);
    encode_kernel<<<grid, 256, smem, st>>>(t.d_g, t.d_src, g.n, g.rows, kw, d_info, (long long)frames, d_cw, nw32);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(std::string("encode_kernel: ") + cudaGetErrorString(e)); return LDPC_ERR_CUDA; }
    return LDPC_OK;
}

}  // namespace ldpc

// the device tables hang off the generator object (created on first use, one device at a time)
struct ldpc_gen_device { ldpc::EncTables t; };

extern "C" {

int ldpc_encode_batch_device(ldpc_gen *gen, int device, const uint32_t *d_info, size_t frames, uint32_t *d_codewords, void *stream)
{
    if (!gen || !d_info || !d_codewords) { ldpc::set_error("NULL generator / info / codewords"); return LDPC_ERR_ARG; }
    if (frames == 0) return LDPC_OK;
    if (!gen->dev) gen->dev = new ldpc_gen_device;
    int rc = ldpc::enc_prepare(*gen, device, gen->dev->t);
    if (rc != LDPC_OK) return rc;
    return ldpc::enc_launch(*gen, gen->dev->t, d_info, frames, d_codewords, stream ? (cudaStream_t)stream : gen->dev->t.stream);
}

int ldpc_encode_batch(ldpc_gen *gen, int device, const uint8_t *info, size_t frames, uint32_t *codewords)
{
    if (!gen || !info || !codewords) { ldpc::set_error("NULL generator / info / codewords"); return LDPC_ERR_ARG; }
    if (frames == 0) return LDPC_OK;
    if (!gen->dev) gen->dev = new ldpc_gen_device;
    ldpc::EncTables &t = gen->dev->t;
    int rc = ldpc::enc_prepare(*gen, device, t);
    if (rc != LDPC_OK) return rc;
    const int k = gen->n - gen->rows, kw = (k + 31) / 32, kb = (k + 7) / 8, nw32 = (gen->n + 31) / 32;
    if (t.cap < frames) {
        cudaFree(t.d_info); cudaFree(t.d_cw);
        t.d_info = nullptr; t.d_cw = nullptr; t.cap = 0;
        if (cudaMalloc(&t.d_info, frames * kw * sizeof(uint32_t)) != cudaSuccess ||
            cudaMalloc(&t.d_cw, frames * nw32 * sizeof(uint32_t)) != cudaSuccess) {
            ldpc::set_error("encoder staging allocation failed"); return LDPC_ERR_NOMEM;
        }
        t.cap = frames;
    }
    // message bytes (LSB first) are the little-endian bytes of the packed words; pad each frame to whole words
    std::vector<uint32_t> packed(frames * kw, 0u);
    for (size_t f = 0; f < frames; ++f) {
        std::memcpy(&packed[f * kw], info + f * kb, kb);
        if (k % 32) packed[f * kw + kw - 1] &= (1u << (k % 32)) - 1u;  // bits past k do not exist (:179-183)
    }
    cudaError_t e = cudaMemcpyAsync(t.d_info, packed.data(), packed.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, t.stream);
    if (e == cudaSuccess) { rc = ldpc::enc_launch(*gen, t, t.d_info, frames, t.d_cw, t.stream); if (rc != LDPC_OK) return rc; }
    if (e == cudaSuccess) e = cudaMemcpyAsync(codewords, t.d_cw, frames * nw32 * sizeof(uint32_t), cudaMemcpyDeviceToHost, t.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(t.stream);
    if (e != cudaSuccess) { ldpc::set_error(std::string("ldpc_encode_batch: ") + cudaGetErrorString(e)); return LDPC_ERR_CUDA; }
    return LDPC_OK;
}

void ldpc_gen_release_device(ldpc_gen *gen)
{
    if (!gen || !gen->dev) return;
    ldpc::enc_release(gen->dev->t);
    delete gen->dev;
    gen->dev = nullptr;
}

}  // extern "C"
