// Pipe-throughput microbenchmark for the integer ops the decode kernels are built from.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on a B200.
// Prints thread-ops per clock per SM for each op (8 independent chains/thread, 1024 thr/SM).
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

#define ITER 4096
#define NCH 8

#define DEF_KERNEL(NAME, BODY)                                                        \
__global__ void __launch_bounds__(1024, 1) k_##NAME(uint32_t* out, const uint32_t* in, long long* cyc) { \
    uint32_t r[NCH];                                                                   \
    uint32_t y = in[threadIdx.x & 31], z = in[32 + (threadIdx.x & 31)];                \
    _Pragma("unroll") for (int j = 0; j < NCH; ++j) r[j] = in[64 + j] + threadIdx.x;   \
    __syncthreads();                                                                   \
    long long t0 = clock64();                                                          \
    for (int it = 0; it < ITER; ++it) {                                                \
        _Pragma("unroll") for (int j = 0; j < NCH; ++j) { uint32_t x = r[j]; BODY; r[j] = x; } \
    }                                                                                  \
    long long t1 = clock64();                                                          \
    uint32_t s = 0;                                                                    \
    _Pragma("unroll") for (int j = 0; j < NCH; ++j) s ^= r[j];                         \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;                                    \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                   \
}

#define ASM1(ins) asm volatile(ins : "+r"(x) : "r"(y), "r"(z))

DEF_KERNEL(iadd3,   ASM1("add.u32 %0, %0, %1;"))
DEF_KERNEL(lop3,    ASM1("lop3.b32 %0, %0, %1, %2, 0x96;"))
DEF_KERNEL(shf,     ASM1("shf.r.wrap.b32 %0, %0, %1, %2;"))
DEF_KERNEL(shr,     asm volatile("shr.u32 %0, %0, 2; add.u32 %0, %0, %1;" : "+r"(x) : "r"(y)))
DEF_KERNEL(vminu2,  ASM1("min.u16x2 %0, %0, %1;"))
DEF_KERNEL(vadd2,   ASM1("add.u16x2 %0, %0, %1;"))
DEF_KERNEL(viaddmax,x = __viaddmax_s16x2(x, y, z))
DEF_KERNEL(vimax3,  x = __vimax3_s16x2(x, y, z))
DEF_KERNEL(imnmx,   ASM1("min.s32 %0, %0, %1;"))
DEF_KERNEL(prmt,    ASM1("prmt.b32 %0, %0, %1, 0xbb99;"))
DEF_KERNEL(imad,    ASM1("mad.lo.u32 %0, %0, %1, %2;"))
DEF_KERNEL(imadhi,  ASM1("mad.hi.u32 %0, %0, %1, %2;"))
DEF_KERNEL(dp2a,    x = (uint32_t)__dp2a_lo((int)x, (int)y, (int)z))
DEF_KERNEL(dp4a,    x = (uint32_t)__dp4a((int)x, (int)y, (int)z))
// mixes: one ALU op + one FMA-pipe op per step (2 ops)
DEF_KERNEL(mix_lop_imad,  asm volatile("lop3.b32 %0, %0, %1, %2, 0x96; mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z)))
DEF_KERNEL(mix_min_imad,  asm volatile("min.u16x2 %0, %0, %1; mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z)))
DEF_KERNEL(mix_lop_imadhi,asm volatile("lop3.b32 %0, %0, %1, %2, 0x96; mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z)))
DEF_KERNEL(mix_2lop_imad, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96; lop3.b32 %0, %0, %2, %1, 0xe8; mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z)))
DEF_KERNEL(mix_lop_2imad, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96; mad.lo.u32 %0, %0, %1, %2; mad.lo.u32 %0, %0, %2, %1;" : "+r"(x) : "r"(y), "r"(z)))
DEF_KERNEL(hmnmx2,  ASM1("min.f16x2 %0, %0, %1;"))
DEF_KERNEL(mix_lop_hmnmx2, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96; min.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y), "r"(z)))
DEF_KERNEL(mix_vmin_hmnmx2, asm volatile("min.u16x2 %0, %0, %2; min.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y), "r"(z)))
DEF_KERNEL(hadd2,   ASM1("add.f16x2 %0, %0, %1;"))
DEF_KERNEL(mix_lop_hadd2, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96; add.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y), "r"(z)))
DEF_KERNEL(mix_min_dp2a,  { asm volatile("min.u16x2 %0, %0, %1;" : "+r"(x) : "r"(y)); x = (uint32_t)__dp2a_lo((int)x, (int)y, (int)z); })

// g(a,b) candidates: x = chain value, y = message magnitude (packed 16x2, two frames per register)
__device__ __forceinline__ uint32_t g_packed_alu(uint32_t a, uint32_t b) {
    uint32_t mn = __vminu2(a, b), mx = __vmaxu2(a, b);
    uint32_t s = a + b, d = mx - mn;
    uint32_t qs = (s >> 2) & 0x003F003Fu, qd = (d >> 2) & 0x003F003Fu;
    uint32_t us = __vminu2(qs, 0x000A000Au), ud = __vminu2(qd, 0x000A000Au);
    return mn + ud - us;
}
__device__ __forceinline__ uint32_t g_packed_mix(uint32_t a, uint32_t b) {
    uint32_t mn = __vminu2(a, b);
    uint32_t s, d, hs, hd;
    asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(s) : "r"(a), "r"(b));
    asm("mad.lo.u32 %0, %1, 0xfffffffe, %2;" : "=r"(d) : "r"(mn), "r"(s));
    asm("mul.hi.u32 %0, %1, 0x40000000;" : "=r"(hs) : "r"(s));
    asm("mul.hi.u32 %0, %1, 0x40000000;" : "=r"(hd) : "r"(d));
    uint32_t qs = hs & 0x003F003Fu, qd = hd & 0x003F003Fu;
    uint32_t us = __vminu2(qs, 0x000A000Au), ud = __vminu2(qd, 0x000A000Au);
    uint32_t t;
    asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(t) : "r"(ud), "r"(mn));
    return t - us;
}
__device__ __forceinline__ uint32_t g_scalar(uint32_t a, uint32_t b) {  // one frame per register, int32
    int mn = min((int)a, (int)b), mx = max((int)a, (int)b);
    int s = mn + mx, d = mx - mn;
    int us = min(10, (s >> 2) & 63), ud = min(10, (d >> 2) & 63);
    return (uint32_t)(mn + ud - us);
}
__device__ __forceinline__ uint32_t g_scalar_dp(uint32_t a, uint32_t b) {  // s,d packed as 16x2, finish with IDP.2A
    int mn = min((int)a, (int)b);
    uint32_t s = a + b, d = s - 2u * (uint32_t)mn;
    uint32_t p = __byte_perm(s, d, 0x5410);
    uint32_t q = (p >> 2) & 0x003F003Fu;
    uint32_t u = __vminu2(q, 0x000A000Au);
    return (uint32_t)__dp2a_lo((int)u, 0x000001FF, mn);   // mn - u.lo + u.hi  (b0=-1 for s, b1=+1 for d)
}
DEF_KERNEL(g_packed_alu, x = g_packed_alu(x, y + j))
DEF_KERNEL(g_packed_mix, x = g_packed_mix(x, y + j))
DEF_KERNEL(g_scalar,     x = g_scalar(x, y + j))
DEF_KERNEL(g_scalar_dp,  x = g_scalar_dp(x, y + j))

// shared-memory load throughput (conflict-free 32-bit)
__global__ void __launch_bounds__(1024, 1) k_lds(uint32_t* out, const uint32_t* in, long long* cyc) {
    __shared__ uint32_t sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = in[i & 63] + i;
    __syncthreads();
    uint32_t acc = 0; int idx = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        #pragma unroll
        for (int j = 0; j < NCH; ++j) acc ^= sm[(idx + j * 1024 + it) & 8191];
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename K> void run(const char* name, K kern, int ops_per_step, int nsm, uint32_t* out, uint32_t* in, long long* cyc) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<nsm, 1024>>>(out, in, cyc); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<<<nsm, 1024>>>(out, in, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[nsm]; cudaMemcpy(h, cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nsm; ++i) avg += (double)h[i]; avg /= nsm; delete[] h;
    double steps = (double)ITER * NCH * 1024;
    printf("%-18s steps/clk/SM %8.2f  ops/clk/SM %8.2f  (ops/step %d)  cycles %.0f  ms %.3f  eff_MHz %.0f  err=%d\n",
           name, steps / avg, steps * ops_per_step / avg, ops_per_step, avg, ms, avg / ms / 1e3, (int)cudaGetLastError());
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    printf("device %s SMs %d clock %d kHz\n", p.name, nsm, p.clockRate);
    uint32_t *out, *in; long long* cyc;
    cudaMalloc(&out, (size_t)nsm * 1024 * 4); cudaMalloc(&in, 8192 * 4); cudaMalloc(&cyc, nsm * 8);
    uint32_t h[8192]; for (int i = 0; i < 8192; ++i) h[i] = 0x01230457u * (i + 1) & 0x3fff3fffu;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
#define RUN(NAME, OPS) run(#NAME, k_##NAME, OPS, nsm, out, in, cyc)
    RUN(iadd3, 1); RUN(lop3, 1); RUN(shf, 1); RUN(shr, 2); RUN(vminu2, 1); RUN(vadd2, 1); RUN(viaddmax, 1); RUN(vimax3, 1);
    RUN(imnmx, 1); RUN(prmt, 1); RUN(imad, 1); RUN(imadhi, 1); RUN(dp2a, 1); RUN(dp4a, 1);
    RUN(mix_lop_imad, 2); RUN(mix_min_imad, 2); RUN(mix_lop_imadhi, 2); RUN(mix_2lop_imad, 3); RUN(mix_lop_2imad, 3); RUN(mix_min_dp2a, 2);
    RUN(hmnmx2, 1); RUN(mix_lop_hmnmx2, 2); RUN(mix_vmin_hmnmx2, 2); RUN(hadd2, 1); RUN(mix_lop_hadd2, 2);
    RUN(g_packed_alu, 1); RUN(g_packed_mix, 1); RUN(g_scalar, 1); RUN(g_scalar_dp, 1);
    RUN(lds, 1);
    return 0;
}
