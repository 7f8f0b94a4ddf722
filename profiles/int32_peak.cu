// int32_peak.cu -- issue-rate microbenchmark for the pipes the 2048 kernels live on (SURVEY 8(d):
// "the builder must add an INT32 microbenchmark", MEASURED_PEAKS.json has none).
//
// Every kernel runs kChains independent dependency chains per thread of one instruction kind, with one
// wave of 2 x 1024-thread blocks per SM (16 warps per scheduler), and reports warp-instructions per clock
// per scheduler (SMSP) from clock64() spans of the blocks, plus warp-instructions per second from CUDA events.  These
// are the denominators for the `issue` object in bench.py: a pure ALU-pipe stream (LOP3/SHF/IADD3/PRMT)
// tops out at the ALU pipe's rate, an ALU+FMA mix at the scheduler's one instruction per clock.
//
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o int32_peak int32_peak.cu
// Run:   ./int32_peak            (prints one JSON object)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                  \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

constexpr int kChains = 8;
constexpr int kUnroll = 8;
constexpr int kThreads = 1024;

enum Op { LOP3, SHF, IADD3, PRMT, POPC, IMAD, FFMA, DADD, DMUL, MIX_ALU_FMA, MIX_LOP_SHF, LDS16, IMADHI, IMADWIDE, SEL, I2FD, kOps };
static const char *kOpName[kOps] = {"lop3",      "shf",  "iadd", "prmt", "popc", "imad", "ffma",
                                    "dadd",      "dmul", "mix_alu_fma",   "mix_lop3_shf", "lds_u16",
                                    "imad_hi",   "imad_wide", "sel", "i2f_f64"};
// instructions per chain step
static const int kOpInstr[kOps] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 1, 1, 2, 2, 2};   // imad_wide, sel, i2f carry one helper op

template <int OP>
__device__ __forceinline__ void step(uint32_t &x, double &d, uint32_t a, uint32_t b, const uint16_t *tab) {
    if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(a));
    if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(x));
    if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == FFMA) {
        float f = __uint_as_float(x);
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)));
        x = __float_as_uint(f);
    }
    if (OP == DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d) : "d"((double)a));
    if (OP == DMUL) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d) : "d"(1.0000001));
    if (OP == MIX_ALU_FMA) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    }
    if (OP == MIX_LOP_SHF) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
        asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    }
    if (OP == IMADHI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(a));
    if (OP == IMADWIDE) {
        uint64_t w;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x), "r"(a));
        x = (uint32_t)(w >> 32) ^ (uint32_t)w;      // one extra LOP3 per step (counted below as 1 instruction pair)
    }
    if (OP == SEL) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(x) : "r"(a), "r"(b));
    if (OP == I2FD) {
        double t;
        asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(t) : "r"(x));
        x = (uint32_t)__double_as_longlong(t) + (uint32_t)(__double_as_longlong(t) >> 32);
    }
    if (OP == LDS16) x = tab[x];   // dependent 16-bit table lookups, lane-random addresses
}

template <int OP>
__global__ void __launch_bounds__(kThreads) peak_kernel(uint32_t *out, long long *cycles, uint32_t a, uint32_t b,
                                                       int iters) {
    extern __shared__ uint16_t tab[];
    if (OP == LDS16) {
        // a 64 Ki-entry permutation-like table so that chains keep wandering over the whole table
        for (int i = threadIdx.x; i < 65536; i += kThreads) tab[i] = (uint16_t)(i * 40503u + 12345u);
        __syncthreads();
    }
    uint32_t x[kChains];
    double d[kChains];
    const uint32_t tid = blockIdx.x * kThreads + threadIdx.x;
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
        x[c] = (tid * 2654435761u + c * 40503u) & (OP == LDS16 ? 0xffffu : 0xffffffffu);
        d[c] = 1.0 + c;
    }
    __shared__ long long t_first, t_last;
    if (threadIdx.x == 0) {
        t_first = 0x7fffffffffffffffll;
        t_last = 0;
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
#pragma unroll
            for (int c = 0; c < kChains; ++c)   // adds take a neighbouring chain as operand so ptxas cannot fold them
                step<OP>(x[c], d[c], OP == IADD3 ? x[(c + 1) % kChains] : a, b, tab);
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc ^= x[c] ^ (uint32_t)__double2ll_rn(d[c]);
    out[tid] = acc;
    // block time = first warp's start to last warp's end (clock64 is one counter per SM)
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&t_first, t0);
        atomicMax(&t_last, t1);
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t_last - t_first;
}

template <int OP>
static void run(int sms, int clock_khz, bool last) {
    const int blocks = sms * 2;   // 2 x 1024 threads = a full SM, one wave
    const int iters = (OP == DADD || OP == DMUL) ? 256 : 2048;
    const size_t smem = (OP == LDS16) ? 65536 * sizeof(uint16_t) : 0;
    const int grid = (OP == LDS16) ? sms : blocks;   // 128 KiB of table: one block per SM
    uint32_t *out;
    long long *cycles;
    CK(cudaMalloc(&out, (size_t)blocks * kThreads * sizeof(uint32_t)));
    CK(cudaMalloc(&cycles, blocks * sizeof(long long)));
    if (smem) CK(cudaFuncSetAttribute(peak_kernel<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        peak_kernel<OP><<<grid, kThreads, smem>>>(out, cycles, 0x9e3779b9u + rep, 7u, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    CK(cudaGetLastError());
    long long *h = (long long *)malloc(grid * sizeof(long long));
    CK(cudaMemcpy(h, cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int i = 0; i < grid; ++i) mean += (double)h[i];
    mean /= grid;
    free(h);
    const double warps_per_sm = (double)grid / sms * (kThreads / 32);
    const double inst_per_warp = (double)iters * kUnroll * kChains * kOpInstr[OP];
    const double per_clk_smsp = warps_per_sm * inst_per_warp / mean / 4.0;
    const double per_s = (double)grid * (kThreads / 32) * inst_per_warp / (best_ms * 1e-3);
    // per clock at the maximum SM clock (bench.py samples 1,965 MHz under load on these boxes); the
    // clock64()-based figure is kept beside it: on B200 that counter does not tick at the SM clock.
    const double per_clk_max = per_s / ((double)sms * 4.0 * clock_khz * 1e3);
    printf("  \"%s\": {\"warp_inst_per_s\": %.4e, \"per_clk_per_smsp_at_max_clock\": %.4f, "
           "\"per_clock64_tick_per_smsp\": %.4f, \"ms\": %.4f, \"warps_per_smsp\": %.0f}%s\n",
           kOpName[OP], per_s, per_clk_max, per_clk_smsp, best_ms, warps_per_sm / 4.0, last ? "" : ",");
    CK(cudaFree(out));
    CK(cudaFree(cycles));
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clock_khz = 0;
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"sm_max_khz\": %d, \"chains_per_thread\": %d,\n", p.name,
           p.multiProcessorCount, clock_khz, kChains);
    const int sms = p.multiProcessorCount;
    run<LOP3>(sms, clock_khz, false);
    run<SHF>(sms, clock_khz, false);
    run<IADD3>(sms, clock_khz, false);
    run<PRMT>(sms, clock_khz, false);
    run<POPC>(sms, clock_khz, false);
    run<IMAD>(sms, clock_khz, false);
    run<FFMA>(sms, clock_khz, false);
    run<DADD>(sms, clock_khz, false);
    run<DMUL>(sms, clock_khz, false);
    run<MIX_ALU_FMA>(sms, clock_khz, false);
    run<MIX_LOP_SHF>(sms, clock_khz, false);
    run<LDS16>(sms, clock_khz, false);
    run<IMADHI>(sms, clock_khz, false);
    run<IMADWIDE>(sms, clock_khz, false);
    run<SEL>(sms, clock_khz, false);
    run<I2FD>(sms, clock_khz, true);
    printf("}\n");
    return 0;
}
