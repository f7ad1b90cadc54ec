// Integer / DPX pipe micro-benchmark for B200 (sm_100a).
//
// Measures, on the GPU box, the per-SM throughput (thread-ops per clock) and the
// dependent-issue latency of the instructions the NW inner loops are built from:
// VIMNMX3, VIADDMNMX (s32 and .S16x2), IMAD, IDP.4A, PRMT, SHFL, LDS, and the
// mixes the kernels actually issue.  The numbers are the "R" of the integer-issue
// roofline in DESIGN.md (SURVEY.md 8d asks for R to be measured, not assumed).
//
// Build (on the GPU box; -cudart shared keeps the static runtime out of the binary):
//        nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -cudart shared -o /tmp/pipe_microbench tools/pipe_microbench.cu
// Run:   /tmp/pipe_microbench          (prints one JSON object per line)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ILP = 8;
constexpr int ITERS = 2048;

enum Op {
    OP_VIMNMX3 = 0, OP_VIADDMNMX, OP_VIMNMX3_16, OP_VIADDMNMX_16, OP_VIMNMX2_16, OP_IMAD, OP_DP4A, OP_PRMT,
    OP_IADD, OP_LOP3,
    OP_MIX_IMAD_MAX3,      // d = diag*1+s (IMAD) ; c = max3(d, up, left)
    OP_MIX_DP4A_MAX3,      // d = dp4a(word, sel, diag) ; c = max3(d, up, left)
    OP_MIX_ADDMAX_MAX,     // t = viaddmax(diag, s, up) ; c = max(t, left)
    OP_MIX_16_PRMT,        // s = prmt(w0,w1) ; t = viaddmax16(diag, s, up) ; c = vimax16(t,left)
    OP_MIX_16_NOPRMT,      // t = viaddmax16(diag, s, up) ; c = vimax16(t,left)
    OP_SHFL, OP_LDS32, OP_LDS64, OP_LDS128,
    OP_MIX_MAX3_SHFL,      // 4x (imad+max3) + 1 shfl
    OP_MIX_MAX3_LDS,       // 4x (imad+max3) + 1 lds32
    OP_DP2A,               // IDP.2A.LO.U16.U8
    OP_MIX_PACKED2,        // two pairs per register: d = dp4a(wA, sel, diag) ; d = dp2a(wB, sel2, d) ; c = vimax3_u16x2(d, up, left)
    OP_MIX_MERGED2,        // round 2: m = prmt(wA, wB) per TWO rows ; d = dp2a_lo/hi(0x80000001, m, diag) ; c = vimax3_u16x2(d, up, left)
    OP_MIX_DP2A_MAX3_16,   // the same without the permute (the ceiling if the two profile words came pre-merged)
    OP_MIX_ADDMERGE2,      // the merge as ONE integer add of two words with disjoint bytes (profiles stored two rows per word)
    OP_MIX_LOPMERGE2,      // ... as a LOP3
    OP_MIX_IMADMERGE2,     // ... as an IMAD (fma pipe)
    OP_COUNT
};

static const char* op_names[OP_COUNT] = {
    "VIMNMX3.s32", "VIADDMNMX.s32", "VIMNMX3.s16x2", "VIADDMNMX.s16x2", "VIMNMX.s16x2", "IMAD", "IDP.4A", "PRMT",
    "IADD(add.s32)", "LOP3",
    "mix: IMAD+VIMNMX3 (per cell)", "mix: IDP4A+VIMNMX3 (per cell)", "mix: VIADDMNMX+VIMNMX (per cell)",
    "mix16: PRMT+VIADDMNMX16+VIMNMX16 (per 2 cells)", "mix16: VIADDMNMX16+VIMNMX16 (per 2 cells)",
    "SHFL.UP", "LDS.32", "LDS.64", "LDS.128",
    "mix: 4x(IMAD+VIMNMX3)+SHFL (per 4 cells)", "mix: 4x(IMAD+VIMNMX3)+LDS32 (per 4 cells)",
    "IDP.2A", "mix16: IDP4A+IDP2A+VIMNMX3.U16x2 (per 2 cells)",
    "mix16: 0.5 PRMT+IDP2A+VIMNMX3.U16x2 (per 2 cells)", "mix16: IDP2A+VIMNMX3.U16x2 (per 2 cells)",
    "mix16: 0.5 IADD+IDP2A+VIMNMX3.U16x2 (per 2 cells)", "mix16: 0.5 LOP3+IDP2A+VIMNMX3.U16x2 (per 2 cells)",
    "mix16: 0.5 IMAD+IDP2A+VIMNMX3.U16x2 (per 2 cells)"};

// "units" per inner-loop body per accumulator (what the reported rate counts).
template <int OP>
__device__ __forceinline__ void body(int (&a)[ILP], int (&b)[ILP], int p, int q, int sel, const int* sm, int lane) {
    unsigned smaddr = (unsigned)__cvta_generic_to_shared(sm);
#pragma unroll
    for (int k = 0; k < ILP; k++) {
        if (OP == OP_VIMNMX3) a[k] = __vimax3_s32(a[k], b[k], p);
        else if (OP == OP_VIADDMNMX) a[k] = __viaddmax_s32(a[k], p, b[k]);
        else if (OP == OP_VIMNMX3_16) a[k] = __vimax3_s16x2(a[k], b[k], p);
        else if (OP == OP_VIADDMNMX_16) a[k] = __viaddmax_s16x2(a[k], p, b[k]);
        else if (OP == OP_VIMNMX2_16) a[k] = __vmaxs2(a[k], b[k]), b[k] = __vmaxs2(b[k], p);
        else if (OP == OP_IMAD) a[k] = a[k] * p + q;
        else if (OP == OP_DP4A) a[k] = __dp4a(b[k], sel, a[k]);
        else if (OP == OP_PRMT) a[k] = __byte_perm(a[k], b[k], sel);
        else if (OP == OP_DP2A) a[k] = __dp2a_lo((unsigned)b[k], (unsigned)sel, (unsigned)a[k]);
        else if (OP == OP_MIX_PACKED2) {
            unsigned d = __dp4a((unsigned)p, (unsigned)sel, (unsigned)b[k]);
            d = __dp2a_lo((unsigned)q, (unsigned)sel, d);
            unsigned c = __vimax3_u16x2(d, (unsigned)a[k], (unsigned)b[k]);
            b[k] = a[k]; a[k] = c;
        }
        else if (OP == OP_MIX_MERGED2) {
            if ((k & 1) == 0) {
                const unsigned m = __byte_perm((unsigned)p + (unsigned)b[k], (unsigned)q, (unsigned)sel);
                unsigned d0 = __dp2a_lo(0x80000001u, m, (unsigned)b[k]);
                unsigned d1 = __dp2a_hi(0x80000001u, m, (unsigned)b[k + 1]);
                unsigned c0 = __vimax3_u16x2(d0, (unsigned)a[k], (unsigned)b[k]);
                unsigned c1 = __vimax3_u16x2(d1, (unsigned)a[k + 1], (unsigned)b[k + 1]);
                b[k] = a[k]; a[k] = c0; b[k + 1] = a[k + 1]; a[k + 1] = c1;
            }
        }
        else if (OP == OP_MIX_ADDMERGE2 || OP == OP_MIX_LOPMERGE2 || OP == OP_MIX_IMADMERGE2) {
            if ((k & 1) == 0) {
                unsigned m;
                if (OP == OP_MIX_ADDMERGE2) asm volatile("add.u32 %0, %1, %2;" : "=r"(m) : "r"((unsigned)b[k]), "r"((unsigned)q));
                else if (OP == OP_MIX_LOPMERGE2) asm volatile("lop3.b32 %0, %1, %2, %3, 0xfe;" : "=r"(m) : "r"((unsigned)b[k]), "r"((unsigned)q), "r"((unsigned)p));
                else asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(m) : "r"((unsigned)b[k]), "r"((unsigned)q), "r"((unsigned)p));
                unsigned d0 = __dp2a_lo(0x80000001u, m, (unsigned)b[k]);
                unsigned d1 = __dp2a_hi(0x80000001u, m, (unsigned)b[k + 1]);
                unsigned c0 = __vimax3_u16x2(d0, (unsigned)a[k], (unsigned)b[k]);
                unsigned c1 = __vimax3_u16x2(d1, (unsigned)a[k + 1], (unsigned)b[k + 1]);
                b[k] = a[k]; a[k] = c0; b[k + 1] = a[k + 1]; a[k + 1] = c1;
            }
        }
        else if (OP == OP_MIX_DP2A_MAX3_16) {
            unsigned d = __dp2a_lo(0x80000001u, (unsigned)p, (unsigned)b[k]);
            unsigned c = __vimax3_u16x2(d, (unsigned)a[k], (unsigned)b[k]);
            b[k] = a[k]; a[k] = c;
        }
        else if (OP == OP_IADD) { asm volatile("add.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(b[k])); }
        else if (OP == OP_LOP3) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b[k]), "r"(p)); }
        else if (OP == OP_MIX_IMAD_MAX3) {
            int d = b[k] * q + p;              // q == 1 at run time -> IMAD (fma pipe)
            int c = __vimax3_s32(d, a[k], b[k]);
            b[k] = a[k]; a[k] = c;
        } else if (OP == OP_MIX_DP4A_MAX3) {
            int d = __dp4a(p, sel, b[k]);
            int c = __vimax3_s32(d, a[k], b[k]);
            b[k] = a[k]; a[k] = c;
        } else if (OP == OP_MIX_ADDMAX_MAX) {
            int t = __viaddmax_s32(b[k], p, a[k]);
            int c = max(t, q + a[k] - a[k]);   // 2-input max with a run-time operand
            b[k] = a[k]; a[k] = c;
        } else if (OP == OP_MIX_16_PRMT) {
            unsigned s = __byte_perm(b[k], q, sel);
            unsigned t = __viaddmax_s16x2(b[k], s, a[k]);
            unsigned c = __vmaxs2(t, q);
            b[k] = a[k]; a[k] = c;
        } else if (OP == OP_MIX_16_NOPRMT) {
            unsigned t = __viaddmax_s16x2(b[k], p, a[k]);
            unsigned c = __vmaxs2(t, q);
            b[k] = a[k]; a[k] = c;
        } else if (OP == OP_SHFL) a[k] = __shfl_up_sync(0xffffffffu, a[k], 1);
        else if (OP == OP_LDS32) a[k] ^= ((const volatile int*)sm)[lane + k * 32];
        else if (OP == OP_LDS64) { int x, y; asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(smaddr + (lane + k * 32) * 8)); a[k] ^= x; b[k] ^= y; }
        else if (OP == OP_LDS128) { int x, y, z, w; asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(smaddr + (lane + (k & 3) * 32) * 16)); a[k] ^= x ^ z; b[k] ^= y ^ w; }
    }
    if (OP == OP_MIX_MAX3_SHFL || OP == OP_MIX_MAX3_LDS) {
        // 2 accumulator groups of 4 "rows"; each group: 4 x (IMAD + VIMNMX3) + one SHFL or LDS
#pragma unroll
        for (int g = 0; g < 2; g++) {
            int up;
            if (OP == OP_MIX_MAX3_SHFL) up = __shfl_up_sync(0xffffffffu, a[g * 4 + 3], 1);
            else up = ((const volatile int*)sm)[lane + g * 32];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                int k = g * 4 + r;
                int d = b[k] * q + p;
                int c = __vimax3_s32(d, up, a[k]);
                b[k] = up; up = c; a[k] = c;
            }
        }
    }
}

template <int OP>
__global__ void __launch_bounds__(1024) bench_kernel(int* out, long long* cycles, int p, int q, int sel) {
    __shared__ int sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i * p;
    __syncthreads();
    int a[ILP], b[ILP];
    int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < ILP; k++) { a[k] = threadIdx.x * (k + 1) + p; b[k] = threadIdx.x ^ (k * q); }
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
        body<OP>(a, b, p, q, sel, sm, lane);
    }
    long long t1 = clock64();
    int acc = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) acc ^= a[k] ^ b[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Latency: one warp, one dependent chain.
template <int OP>
__global__ void latency_kernel(int* out, long long* cycles, int p, int q, int sel) {
    __shared__ int sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = (i + 1) & 2047;
    __syncthreads();
    int a = threadIdx.x + p, b = q;
    long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < 4096; it++) {
        if (OP == OP_VIMNMX3) a = __vimax3_s32(a, b, p);
        else if (OP == OP_VIADDMNMX) a = __viaddmax_s32(a, p, b);
        else if (OP == OP_VIADDMNMX_16) a = __viaddmax_s16x2(a, p, b);
        else if (OP == OP_VIMNMX3_16) a = __vimax3_s16x2(a, b, p);
        else if (OP == OP_IMAD) a = a * p + q;
        else if (OP == OP_DP4A) a = __dp4a(b, sel, a);
        else if (OP == OP_PRMT) a = __byte_perm(a, b, sel);
        else if (OP == OP_SHFL) a = __shfl_up_sync(0xffffffffu, a, 1);
        else if (OP == OP_LDS32) a = sm[a & 2047];
        else if (OP == OP_MIX_IMAD_MAX3) { int d = a * q + p; a = __vimax3_s32(d, b, p); }   // IMAD -> VIMNMX3 chain
    }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int OP>
static void run_tp(int* d_out, long long* d_cyc, int sms, int threads, double units_per_body, const char* unit) {
    int blocks = sms;  // one block per SM
    bench_kernel<OP><<<blocks, threads>>>(d_out, d_cyc, 1, 1, 0x4140);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    bench_kernel<OP><<<blocks, threads>>>(d_out, d_cyc, 1, 1, 0x4140);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long* h = (long long*)malloc(sizeof(long long) * blocks);
    CK(cudaMemcpy(h, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    double mx = 0, sum = 0;
    for (int i = 0; i < blocks; i++) { if (h[i] > mx) mx = (double)h[i]; sum += (double)h[i]; }
    free(h);
    double units = (double)threads * ITERS * units_per_body;
    printf("{\"kind\":\"throughput\",\"op\":\"%s\",\"threads_per_sm\":%d,\"%s_per_clk_per_sm\":%.2f,\"cycles_max\":%.0f,\"cycles_mean\":%.0f,\"ms\":%.4f}\n",
           op_names[OP], threads, unit, units / mx, mx, sum / blocks, ms);
}

template <int OP>
static void run_lat(int* d_out, long long* d_cyc) {
    latency_kernel<OP><<<1, 32>>>(d_out, d_cyc, 1, 1, 0x4140);
    CK(cudaDeviceSynchronize());
    latency_kernel<OP><<<1, 32>>>(d_out, d_cyc, 1, 1, 0x4140);
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost));
    printf("{\"kind\":\"latency\",\"op\":\"%s\",\"cycles_per_dependent_op\":%.2f}\n", op_names[OP], (double)c / 4096.0);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("{\"kind\":\"device\",\"name\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", prop.name, sms, clk_khz);
    int* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, sizeof(int) * sms * 1024));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * sms));
    for (int threads : {128, 256, 512, 1024}) {
        run_tp<OP_VIMNMX3>(d_out, d_cyc, sms, threads, ILP, "ops");
        run_tp<OP_VIADDMNMX>(d_out, d_cyc, sms, threads, ILP, "ops");
        run_tp<OP_VIMNMX3_16>(d_out, d_cyc, sms, threads, ILP, "ops");
        run_tp<OP_VIADDMNMX_16>(d_out, d_cyc, sms, threads, ILP, "ops");
        run_tp<OP_VIMNMX2_16>(d_out, d_cyc, sms, threads, 2 * ILP, "ops");
        run_tp<OP_IMAD>(d_out, d_cyc, sms, threads, ILP, "ops");
        run_tp<OP_DP4A>(d_out, d_cyc, sms, threads, ILP, "ops");
        run_tp<OP_PRMT>(d_out, d_cyc, sms, threads, ILP, "ops");
        run_tp<OP_IADD>(d_out, d_cyc, sms, threads, ILP, "ops");
        run_tp<OP_LOP3>(d_out, d_cyc, sms, threads, ILP, "ops");
        run_tp<OP_MIX_IMAD_MAX3>(d_out, d_cyc, sms, threads, ILP, "cells");
        run_tp<OP_MIX_DP4A_MAX3>(d_out, d_cyc, sms, threads, ILP, "cells");
        run_tp<OP_MIX_ADDMAX_MAX>(d_out, d_cyc, sms, threads, ILP, "cells");
        run_tp<OP_MIX_16_PRMT>(d_out, d_cyc, sms, threads, 2 * ILP, "cells");
        run_tp<OP_MIX_16_NOPRMT>(d_out, d_cyc, sms, threads, 2 * ILP, "cells");
        run_tp<OP_SHFL>(d_out, d_cyc, sms, threads, ILP, "ops");
        run_tp<OP_LDS32>(d_out, d_cyc, sms, threads, ILP, "ops");
        run_tp<OP_LDS64>(d_out, d_cyc, sms, threads, ILP, "ops");
        run_tp<OP_LDS128>(d_out, d_cyc, sms, threads, ILP, "ops");
        run_tp<OP_MIX_MAX3_SHFL>(d_out, d_cyc, sms, threads, 8, "cells");
        run_tp<OP_MIX_MAX3_LDS>(d_out, d_cyc, sms, threads, 8, "cells");
        run_tp<OP_DP2A>(d_out, d_cyc, sms, threads, ILP, "ops");
        run_tp<OP_MIX_PACKED2>(d_out, d_cyc, sms, threads, 2 * ILP, "cells");
        run_tp<OP_MIX_MERGED2>(d_out, d_cyc, sms, threads, 2 * ILP, "cells");
        run_tp<OP_MIX_DP2A_MAX3_16>(d_out, d_cyc, sms, threads, 2 * ILP, "cells");
        run_tp<OP_MIX_ADDMERGE2>(d_out, d_cyc, sms, threads, 2 * ILP, "cells");
        run_tp<OP_MIX_LOPMERGE2>(d_out, d_cyc, sms, threads, 2 * ILP, "cells");
        run_tp<OP_MIX_IMADMERGE2>(d_out, d_cyc, sms, threads, 2 * ILP, "cells");
    }
    run_lat<OP_VIMNMX3>(d_out, d_cyc);
    run_lat<OP_VIADDMNMX>(d_out, d_cyc);
    run_lat<OP_VIMNMX3_16>(d_out, d_cyc);
    run_lat<OP_VIADDMNMX_16>(d_out, d_cyc);
    run_lat<OP_IMAD>(d_out, d_cyc);
    run_lat<OP_DP4A>(d_out, d_cyc);
    run_lat<OP_PRMT>(d_out, d_cyc);
    run_lat<OP_SHFL>(d_out, d_cyc);
    run_lat<OP_LDS32>(d_out, d_cyc);
    run_lat<OP_MIX_IMAD_MAX3>(d_out, d_cyc);
    return 0;
}
