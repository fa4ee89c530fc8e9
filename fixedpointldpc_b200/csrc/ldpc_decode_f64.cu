// ldpc_decode_f64.cu -- the reference's floating-point decoder FP_Decoder::decode_general(const double *)
// (ArrayLDPC_Decoder.cpp:735-933) with the exact box-plus sxor(double, double) (:724-732) and checkPost() (:335-372),
// batched on the GPU in FP64.
//
// Same two-phase flooding schedule, the same forward/backward recursion and the same evaluation order of every sum as
// the reference (the posterior is ((0 + c2v_0) + c2v_1 + ...) + LLR, v2c = posterior - c2v_k), so the only source of
// difference is libm: CUDA's log/exp against the host's, at most an ulp or two per call.  The reference never calls
// this decoder from a live driver ("dead code path"); it is the tool that quantifies the fixed-point
// loss, so the design is the simple one: a CTA per frame, messages in an FP64 workspace in global memory (L2
// resident: E doubles per CTA), thread per check / per variable.  Parity bar (tests/test_gpu_f64.py): iteration counts
// and decoded bits equal to the reference's on the golden frames, posteriors and messages within a stated relative
// tolerance.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/ldpc_capi.h"
#include "ldpc_code.hpp"

extern "C" int ldpc_decoder_device(const ldpc_decoder *dec);
extern "C" const ldpc_code *ldpc_decoder_code(const ldpc_decoder *dec);
extern "C" int ldpc_decoder_max_iter(const ldpc_decoder *dec);

namespace ldpc {

struct F64Tables {
    int *cdeg = nullptr, *vdeg = nullptr, *clist = nullptr;  // [m], [n], [m][dc_max]
    unsigned int *vedge = nullptr;                            // [dv_max][n]: slot*m + check (EdgeRAM index)
};

__device__ __forceinline__ int sgn_ref(double x) { return x > 0 ? 1 : -1; }  // ArrayLDPCMacro.h:218-224

// ArrayLDPC_Decoder.cpp:724-732, every operation rounded on its own (no contraction)
__device__ __forceinline__ double sxor_f64(double x, double y)
{
    const double v1 = fabs(x), v2 = fabs(y);
    const double sum_abs = __dadd_rn(v1, v2);
    const double diff_abs = fabs(__dsub_rn(v1, v2));
    const double mn = v2 < v1 ? v2 : v1;  // std::min
    const double a = log(__dadd_rn(1.0, exp(-sum_abs)));
    const double b = log(__dadd_rn(1.0, exp(-diff_abs)));
    const double mag = __dsub_rn(__dadd_rn(mn, a), b);
    return __dmul_rn((double)(sgn_ref(x) * sgn_ref(y)), mag);
}

constexpr int F64_MAX_DC = 64, F64_MAX_DV = 32;

__global__ void __launch_bounds__(256) decode_f64_kernel(F64Tables t, int n, int m, int dc_max, int dv_max, int max_iter,
                                                         const double *llr_all, long long frames, int *iters, uint32_t *bits,
                                                         int nw32, double *post_out, double *v2c_out, double *workspace)
{
    extern __shared__ unsigned char hd[];  // [n] hard decisions of the posteriors
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t words = (size_t)dc_max * m;
    double *edge = workspace + (size_t)blockIdx.x * (words + n);  // EdgeRAM[slot].BRAM[check]
    double *post = edge + words;                                    // Posteriori
    for (long long f = blockIdx.x; f < frames; f += gridDim.x) {
        const double *llr = llr_all + (size_t)f * n;
        // :757-774 EdgeRAM[k][c] = LLR[clist[c][k]]
        for (size_t i = tid; i < words; i += nt) {
            const int k = (int)(i / m), c = (int)(i % m);
            edge[i] = k < t.cdeg[c] ? llr[t.clist[(size_t)c * dc_max + k]] : 0.0;
        }
        __syncthreads();
        int it = 0;
        while (it < max_iter) {
            // check phase :779-833
            for (int c = tid; c < m; c += nt) {
                const int d = t.cdeg[c];
                double fwd[F64_MAX_DC];
                fwd[0] = edge[c];
                for (int k = 1; k < d - 1; ++k) fwd[k] = sxor_f64(fwd[k - 1], edge[(size_t)k * m + c]);
                double bwd = edge[(size_t)(d - 1) * m + c];  // Backward[d-1]
                edge[(size_t)(d - 1) * m + c] = fwd[d - 2];  // c2v[d-1] = Forward[d-2]
                for (int k = d - 2; k >= 1; --k) {
                    const double mk = edge[(size_t)k * m + c];
                    edge[(size_t)k * m + c] = sxor_f64(fwd[k - 1], bwd);  // c2v[k] = sxor(Forward[k-1], Backward[k+1])
                    bwd = sxor_f64(bwd, mk);                              // Backward[k]
                }
                edge[c] = bwd;  // c2v[0] = Backward[1]
            }
            __syncthreads();
            // variable phase :838-872
            for (int v = tid; v < n; v += nt) {
                const int d = t.vdeg[v];
                double x[F64_MAX_DV];
                double accum = 0;
                for (int j = 0; j < d; ++j) {
                    x[j] = edge[t.vedge[(size_t)j * n + v]];
                    accum = __dadd_rn(accum, x[j]);
                }
                accum = __dadd_rn(accum, llr[v]);
                post[v] = accum;
                hd[v] = accum > 0 ? 0 : 1;  // checkPost :345
                for (int j = 0; j < d; ++j) edge[t.vedge[(size_t)j * n + v]] = __dsub_rn(accum, x[j]);
            }
            __syncthreads();
            ++it;
            // checkPost :335-372
            int fail = 0;
            for (int c = tid; c < m; c += nt) {
                const int d = t.cdeg[c];
                unsigned int sum = 0;
                for (int k = 0; k < d; ++k) sum ^= hd[t.clist[(size_t)c * dc_max + k]];
                fail |= (int)sum;
            }
            if (!__syncthreads_or(fail)) break;
        }
        if (tid == 0) iters[f] = it;
        for (int v0 = 0; v0 < n; v0 += nt) {
            const int v = v0 + tid;
            const uint32_t word = __ballot_sync(0xffffffffu, v < n && hd[v]);
            if ((tid & 31) == 0 && v < n && bits) bits[(size_t)f * nw32 + (v >> 5)] = word;
        }
        if (post_out) for (int v = tid; v < n; v += nt) post_out[(size_t)f * n + v] = post[v];
        if (v2c_out) for (size_t i = tid; i < words; i += nt) v2c_out[(size_t)f * words + i] = edge[i];
        __syncthreads();
    }
}

}  // namespace ldpc

extern "C" int ldpc_decode_batch_f64(ldpc_decoder *dec, const double *llr, size_t frames, int32_t *iters, uint32_t *bits,
                                     double *post, double *v2c)
{
    using namespace ldpc;
    if (!dec || !llr || !iters) { set_error("NULL decoder / llr / iters"); return LDPC_ERR_ARG; }
    if (frames == 0) return LDPC_OK;
    const ldpc_code &c = *ldpc_decoder_code(dec);
    if (c.dc_max > F64_MAX_DC || c.dv_max > F64_MAX_DV) { set_error("check degree > 64 or variable degree > 32"); return LDPC_ERR_UNSUPPORTED; }
    const int max_iter = ldpc_decoder_max_iter(dec);
    cudaError_t e = cudaSetDevice(ldpc_decoder_device(dec));
    const int nw32 = (c.n + 31) / 32;
    const size_t words = (size_t)c.dc_max * c.m;
    std::vector<unsigned int> vedge((size_t)c.dv_max * c.n, 0u);
    for (int v = 0; v < c.n; ++v)
        for (int j = 0; j < c.vdeg[v]; ++j)
            vedge[(size_t)j * c.n + v] = (unsigned int)(c.vslot[(size_t)v * c.dv_max + j] * c.m + c.vlist[(size_t)v * c.dv_max + j]);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ldpc_decoder_device(dec));
    const int grid = (int)std::min<size_t>(frames, (size_t)sms * 4);
    F64Tables t;
    double *d_llr = nullptr, *d_post = nullptr, *d_v2c = nullptr, *d_ws = nullptr;
    int *d_iters = nullptr;
    uint32_t *d_bits = nullptr;
    auto release = [&] {
        cudaFree(t.cdeg); cudaFree(t.vdeg); cudaFree(t.clist); cudaFree(t.vedge);
        cudaFree(d_llr); cudaFree(d_post); cudaFree(d_v2c); cudaFree(d_ws); cudaFree(d_iters); cudaFree(d_bits);
    };
    auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; return e == cudaSuccess; };
    ok(cudaMalloc(&t.cdeg, c.m * sizeof(int))) && ok(cudaMalloc(&t.vdeg, c.n * sizeof(int))) &&
        ok(cudaMalloc(&t.clist, c.clist.size() * sizeof(int))) && ok(cudaMalloc(&t.vedge, vedge.size() * sizeof(unsigned int))) &&
        ok(cudaMalloc(&d_llr, frames * c.n * sizeof(double))) && ok(cudaMalloc(&d_iters, frames * sizeof(int))) &&
        ok(cudaMalloc(&d_bits, frames * nw32 * sizeof(uint32_t))) && ok(cudaMalloc(&d_ws, (size_t)grid * (words + c.n) * sizeof(double)));
    if (post) ok(cudaMalloc(&d_post, frames * c.n * sizeof(double)));
    if (v2c) ok(cudaMalloc(&d_v2c, frames * words * sizeof(double)));
    ok(cudaMemcpy(t.cdeg, c.cdeg.data(), c.m * sizeof(int), cudaMemcpyHostToDevice));
    ok(cudaMemcpy(t.vdeg, c.vdeg.data(), c.n * sizeof(int), cudaMemcpyHostToDevice));
    ok(cudaMemcpy(t.clist, c.clist.data(), c.clist.size() * sizeof(int), cudaMemcpyHostToDevice));
    ok(cudaMemcpy(t.vedge, vedge.data(), vedge.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
    ok(cudaMemcpy(d_llr, llr, frames * c.n * sizeof(double), cudaMemcpyHostToDevice));
    if (e == cudaSuccess) {
        decode_f64_kernel<<<grid, 256, c.n>>>(t, c.n, c.m, c.dc_max, c.dv_max, max_iter, d_llr, (long long)frames, d_iters,
                                              bits ? d_bits : nullptr, nw32, d_post, d_v2c, d_ws);
        ok(cudaGetLastError());
    }
    ok(cudaMemcpy(iters, d_iters, frames * sizeof(int), cudaMemcpyDeviceToHost));
    if (bits) ok(cudaMemcpy(bits, d_bits, frames * nw32 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (post) ok(cudaMemcpy(post, d_post, frames * c.n * sizeof(double), cudaMemcpyDeviceToHost));
    if (v2c) ok(cudaMemcpy(v2c, d_v2c, frames * words * sizeof(double), cudaMemcpyDeviceToHost));
    release();
    if (e != cudaSuccess) { set_error(std::string("ldpc_decode_batch_f64: ") + cudaGetErrorString(e)); return LDPC_ERR_CUDA; }
    return LDPC_OK;
}
