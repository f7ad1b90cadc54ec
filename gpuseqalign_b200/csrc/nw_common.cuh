// nw_common.cuh -- shared device helpers for the B200 NW engine (sm_100a only).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 900)
#error "gpuseqalign_b200 kernels are written for sm_100a (B200) only"
#endif

namespace nwb {

constexpr int kWarp = 32;
constexpr int kMaxLetters = 63;       // substitution alphabet limit (reference uses 25)
constexpr unsigned kFull = 0xffffffffu;

// ---- DPX / integer-pipe primitives -------------------------------------------------------
// One cell of the recurrence in "shifted" coordinates P[i][j] = H[i][j] - (i+j)*gap
// (the reference's own half-way form is nwalign_gpu1_ml_diag.cu:65-70):
//     P[i][j] = max3(P[i-1][j-1] + s', P[i-1][j], P[i][j-1]),   s' = max(subst - 2*gap, 0)
// s' is fetched as one byte of a packed profile word and added with IDP.4A (fma pipe), the
// 3-way max is a single VIMNMX3 (alu pipe): two issue slots per cell on two different pipes.
__device__ __forceinline__ int add_byte(unsigned word, unsigned onehot, int acc)
{
    return (int)__dp4a(word, onehot, (unsigned)acc);      // IDP.4A.U8.U8
}
__device__ __forceinline__ int max3(int a, int b, int c) { return __vimax3_s32(a, b, c); }   // VIMNMX3

// ---- inter-CTA flags ---------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_volatile(const int* p)
{
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// ---- tagged 64-bit elements: value and ready-flag in ONE naturally atomic store -------------
__device__ __forceinline__ unsigned long long pack_tagged(int v, unsigned tag)
{
    return ((unsigned long long)tag << 32) | (unsigned)v;
}
__device__ __forceinline__ unsigned long long ld_relaxed64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// the same element from its two 32-bit halves: ONE naturally aligned 64-bit store (single-copy atomic: value and tag arrive together;
// a .v2.u32 store is two 32-bit accesses in unspecified order as far as the memory model goes).  mov.b64 packs a register pair: no arithmetic.
__device__ __forceinline__ void st_tagged(unsigned long long* p, int v, unsigned tag)
{
    asm volatile("{\n\t.reg .b64 t;\n\tmov.b64 t, {%1, %2};\n\tst.relaxed.gpu.global.u64 [%0], t;\n\t}" ::"l"(p), "r"((unsigned)v), "r"(tag) : "memory");
}
// Set when a wait for a tagged element gave up (a producer that never publishes must not hang the GPU): the host
// reports errorInvalidResult.  ~2^21 polls of >= 20 ns + an L2 round trip each are several seconds.
static __device__ int g_wait_timeout = 0;      // one copy per translation unit (no relocatable device code): the batch kernels report through BatchArgs::err instead
constexpr unsigned kMaxPolls = 1u << 22;
// Polls do NOT sleep: __nanosleep(20) between two polls made single warps oversleep by 1-3 ms every few launches on B200 (one
// straggling map unit then held the whole fill launch: 1.2 ms -> 4.5 ms, profiles/r1q_*); a poll is an L2 round trip anyway.
constexpr unsigned kPollSleepNs = 0;

// returns the value once its tag matches; `first` is an earlier (prefetched) read of *p
__device__ __forceinline__ int wait_tagged(const unsigned long long* p, unsigned long long first, unsigned tag)
{
    unsigned long long v = first;
    unsigned polls = 0;
    while ((unsigned)(v >> 32) != tag) {
        if (kPollSleepNs) __nanosleep(kPollSleepNs);
        v = ld_relaxed64(p);
        if (++polls > kMaxPolls) { g_wait_timeout = 1; break; }
    }
    return (int)(unsigned)v;
}

__device__ __forceinline__ int wait_tagged_count(const unsigned long long* p, unsigned long long first, unsigned tag, unsigned& spins)
{
    unsigned long long v = first;
    unsigned polls = 0;
    while ((unsigned)(v >> 32) != tag) {
        if (kPollSleepNs) __nanosleep(kPollSleepNs);
        v = ld_relaxed64(p);
        spins++;
        if (++polls > kMaxPolls) { g_wait_timeout = 1; break; }
    }
    return (int)(unsigned)v;
}

// streaming (evict-first) global accesses for header traffic that is written once / read once
__device__ __forceinline__ void st_cs(int* p, int v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs4(int4* p, int4 v) { __stcs(p, v); }

}  // namespace nwb
