// nw_scan.cuh -- score of a matrix with FEW ROWS and VERY MANY COLUMNS (BASELINE config 4: 2 048 x 4 194 304) by the
// row-parallel prefix-max formulation (the reference has no counterpart -- its tile wavefront has at
// most min(trows, tcols) = 16 blocks in flight for this shape, nwalign_gpu9_mlsp_diagdiagdiag.cu:589).
//
// In shifted coordinates (P = H - (i+j)*gap, nw_sweep.cuh) a whole row is an elementwise step and a prefix maximum:
//     A[j]    = max(P[i-1][j-1] + s'(y_i, x_j), P[i-1][j])              (no dependence inside the row)
//     P[i][j] = max(A[1..j])                                            (max is associative and exact on integers)
// so the dependence along the long dimension disappears into ONE int per row that crosses a column boundary: the running
// maximum P[i][last column left of the boundary], which is also the diagonal input of the first column to the right for row i+1.
//
// Round 2: the pipeline unit is a WARP, not a CTA, and nothing inside a CTA waits on a chain.  A warp owns a STRIP of 512 columns
// (16 per lane, previous row in registers) for ALL rows and never meets a CTA barrier inside the row loop.  The 16 warps of a CTA
// own 16 consecutive strips (a GROUP of 8 192 columns, taken from a ticket in left-to-right order).  Per row a warp
//   1. runs the IDP.4A + VIMNMX3 cell over its columns (local prefix) and a 5-step shuffle scan of the lane maxima;
//   2. PUBLISHES its strip maximum (row | value, one 8-byte store into a 32-row ring in shared memory) -- it depends on the previous
//      row only, not on any carry of this row;
//   3. gathers what lies to its left with ONE load per lane: lane l < w reads the strip maximum of warp l, lane 31 the carry that
//      came into the group (warp 0 fetched it from the group to the left: a tagged 64-bit element (epoch | value) in global memory,
//      requested a few rows ahead; across GPUs the right neighbour's array is mapped over NVLink with CUDA IPC), polls until the row
//      tags match, and reduces with one REDUX.MAX: no warp-to-warp chain inside the group (round 2's first version had one: 0.2 us
//      per warp and row-1 hop, 10.7 us of skew per group, 40 of 148 SMs busy -- profiles/r2m_*);
//   4. folds that carry into its 16 values; the last warp hands max(carry, all strip maxima) to the next group.
// Back-pressure: every warp reads the last warp's progress counter every 8 rows (the last warp waits for everybody by construction).
// Round 1 had one CTA per 4 096 columns with two __syncthreads and a CTA-wide scan per row: ~1 000 clk per row, 1.5 TCUPS.
#pragma once
#include "nw_sweep.cuh"

namespace nwb {

constexpr int kScanMaxWarps = 16;                       // warps (= strips) per CTA / group: 16, 8 or 4 (template parameter WARPS).  An SM holds 16 warps'
                                                        // profiles either way; narrower groups are more CTAs per SM, whose per-row rounds overlap
constexpr int kScanC = 16;                              // columns per lane
constexpr int kScanStrip = 32 * kScanC;                 // columns per strip (one warp); a group (the unit the host deals to the ranks) is WARPS strips
// (1 / 2 rows: one GPU 4.83 -> 4.66 ms for cfg4, 8 GPUs 3.64 -> 2.53 ms -- there every rank has its groups in flight at once and the
//  start-to-start distance of neighbouring groups, not throughput, sets the time; 4 / 6 rows were the first round-2 values)
constexpr int kScanAhead = 1;                           // rows the carry of the left group is requested ahead (global memory)
constexpr int kScanLag = 2;                             // rows a group falls back behind its left neighbour when a request came back empty
constexpr int kScanRing = 32;                           // rows of the shared-memory rings (strip maxima, group carry)
__host__ __device__ constexpr size_t scan_warp_smem(int S) { return (size_t)(S + 1) * kScanStrip; }
__host__ __device__ constexpr size_t scan_cta_smem(int S, int warps) { return (size_t)warps * scan_warp_smem(S) + (size_t)kScanRing * (warps + 1) * 8; }

struct ScanArgs {
    const uint8_t* y;                 // n row letters
    const uint8_t* x;                 // m column letters
    int n;
    long long m;
    const uint8_t* sprime;
    int S;
    int chunk0;                       // first group of this rank (global group index)
    int nchunks;                      // groups of this rank
    int total_chunks;                 // groups of the whole matrix
    unsigned long long* carry;        // carry[(1 + local group) * n + (i-1)] = (tag << 32 | P[i][last column of that group]); slot 0 = from the left rank
    unsigned long long* peer_carry0;  // the right rank's slot 0 (peer memory), nullptr when this rank owns the last group
    unsigned tag;                     // epoch tag, the same on every rank
    int* ticket;
    unsigned long long* score;        // (tag << 32 | P[n][m]) written by the owner of the last column
    unsigned long long timeout_ns;
    int* err;
    unsigned long long* dbg;          // developer aid (nullable): globaltimer stamps [group < 8][warp][8]
};

__device__ __forceinline__ unsigned long long lds_volatile64(unsigned addr)
{
    unsigned long long v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_volatile64(unsigned addr, unsigned long long v)
{
    asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

// predicated stores without a branch: a lane-divergent `if` around a store makes the compiler wrap every warp collective that follows
// (shuffles, votes, REDUX) in convergence barriers -- a third of the row loop's instructions in the first version of this kernel
__device__ __forceinline__ void sts64_if(bool p, unsigned addr, unsigned long long v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.volatile.shared.u64 [%1], %2;\n\t}" ::"r"((unsigned)p), "r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void sts32_if(bool p, unsigned addr, int v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.volatile.shared.s32 [%1], %2;\n\t}" ::"r"((unsigned)p), "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void stg64_relaxed_if(bool p, unsigned long long* addr, unsigned long long v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.relaxed.gpu.global.u64 [%1], %2;\n\t}" ::"r"((unsigned)p), "l"(addr), "l"(v) : "memory");
}

// The row loop of one strip.  FROM_GLOBAL: this warp (warp 0 of a group that has a group to its left) fetches the group's carry
// from global memory and publishes it in shared memory for the other warps.  HAS_GCIN: the group has a carry coming in at all.
template <bool FROM_GLOBAL, bool HAS_GCIN, int WARPS>
__device__ __forceinline__ void scan_rows(const ScanArgs& a, const int lane, const int w, const unsigned prof_lane_s, const unsigned tot_s, const unsigned gcin_s,
                                          const unsigned prog_s, const unsigned long long* __restrict__ cin_g, unsigned long long* cout_g,
                                          const bool to_global, int (&prev)[kScanC])
{
    const int n = a.n;
    const unsigned S = (unsigned)a.S, tag = a.tag;
    const uint8_t* __restrict__ y = a.y;
    int carry_prev = 0;                        // P[i-1][column left of the strip]: the carry that came in for the previous row
    unsigned long long pend = 0;               // FROM_GLOBAL: lanes 0..7 hold the outstanding request of "their" row
    const unsigned long long t0 = a.timeout_ns ? globaltimer_ns() : 0ull;
    if (FROM_GLOBAL && lane < kScanAhead && 1 + lane <= n) pend = ld_relaxed64(cin_g + lane);
    unsigned ynext = (unsigned)__ldg(y);
    // what this lane gathers per row: the strip maximum of warp `lane` (lanes < w), the group's carry (lane 31), nothing (others)
    const bool gather = lane < w || (HAS_GCIN && lane == 31);
    const unsigned gather_s = (lane == 31 ? gcin_s : tot_s + 8u * (unsigned)lane);
    const unsigned gather_step = (lane == 31 ? 8u : 8u * WARPS);           // bytes from one row's slot to the next
    const unsigned mytot_s = tot_s + 8u * (unsigned)w;
    const bool pub_tot = lane == 31, pub_out = to_global && lane == 31, pub_prog = w == WARPS - 1 && lane == 0;
    const unsigned at0 = gather ? gather_s : mytot_s;                           // lanes that gather nothing load their own warp's slot (and ignore it)
    const unsigned step0 = gather ? gather_step : 8u * WARPS;

    for (int i = 1; i <= n; i++) {
        const unsigned slot = (unsigned)i & (kScanRing - 1);
        // ---- warp 0 of the group: the carry that comes into the group for this row, from global memory into the shared ring
        if (FROM_GLOBAL) {
            const int own = (i - 1) & 7;
            unsigned long long v = pend;
            const bool miss = lane == own && (unsigned)(v >> 32) != tag;
            if (__any_sync(kFull, miss)) {
                // Too close behind the producer: a request issued kScanAhead rows ahead found nothing, and so would the next ones (every
                // row would pay round trips to L2).  Fall back kScanLag rows behind it, then re-issue all outstanding requests at once.
                if (lane == 0) {
                    const int far = min(n, i + kScanLag) - 1;
                    unsigned long long u = ld_relaxed64(cin_g + far);
                    while ((unsigned)(u >> 32) != tag) {      // no __nanosleep: it oversleeps by milliseconds now and then (nw_common.cuh)
                        u = ld_relaxed64(cin_g + far);
                        if (a.timeout_ns && globaltimer_ns() - t0 > a.timeout_ns) { atomicExch(a.err, 1); break; }
                    }
                }
                __syncwarp();
                const int ahead = (lane - own) & 7;                 // lane `own + d` holds the request of row i + d
                if (lane < 8 && ahead < kScanAhead && i + ahead <= n) pend = ld_relaxed64(cin_g + (i - 1 + ahead));
                v = pend;
                if (lane == own) {
                    while ((unsigned)(v >> 32) != tag) {            // (relaxed stores of one thread may become visible out of order)
                        v = ld_relaxed64(cin_g + (i - 1));
                        if (a.timeout_ns && globaltimer_ns() - t0 > a.timeout_ns) { atomicExch(a.err, 1); break; }
                    }
                }
            }
            sts64_if(lane == own, gcin_s + 8u * slot, ((unsigned long long)(unsigned)i << 32) | (unsigned)v);
            if (lane == ((i - 1 + kScanAhead) & 7) && i + kScanAhead <= n) pend = ld_relaxed64(cin_g + (i - 1 + kScanAhead));
        }
        // ---- local prefix over this lane's columns
        const unsigned yl = min(ynext, S);      // (letters were validated at upload)
        ynext = (unsigned)__ldg(y + min(i, n - 1));
        unsigned swv[4];
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(swv[0]), "=r"(swv[1]), "=r"(swv[2]), "=r"(swv[3])
                     : "r"(prof_lane_s + yl * (unsigned)kScanStrip));
        int diag = __shfl_up_sync(kFull, prev[kScanC - 1], 1);
        if (lane == 0) diag = carry_prev;
        int run = 0;                           // P >= 0: neutral start of the running maximum
#pragma unroll
        for (int k = 0; k < kScanC; k++) {
            const int t = add_byte(swv[k >> 2], 1u << (8 * (k & 3)), diag);
            diag = prev[k];
            run = max3(t, prev[k], run);
            prev[k] = run;
        }
        // ---- lane maxima -> inclusive / exclusive prefix maxima across the warp (shfl.up hands a lane its own value when there is
        //      no source lane: the max is then a no-op, no predicate needed); the strip maximum goes out at once
        int incl = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) incl = max(incl, __shfl_up_sync(kFull, incl, d));
        sts64_if(pub_tot, mytot_s + 8u * WARPS * slot, ((unsigned long long)(unsigned)i << 32) | (unsigned)incl);
        int excl = __shfl_up_sync(kFull, incl, 1);
        if (lane == 0) excl = 0;
        // ---- everything to the left of the strip: one load per lane, poll until the row tags match, one REDUX.MAX
        int cin = 0;
        if (w > 0 || HAS_GCIN) {                // (warp-uniform)
            const unsigned at = at0 + step0 * slot;
            unsigned long long v = lds_volatile64(at);
            unsigned polls = 0;
            while (!__all_sync(kFull, !gather || (unsigned)(v >> 32) == (unsigned)i)) {
                v = lds_volatile64(at);
                if (++polls > (1u << 24)) { atomicExch(a.err, 1); break; }
            }
            cin = __reduce_max_sync(kFull, gather ? (int)(unsigned)v : 0);
        }
        // ---- the last warp hands the group's carry on: max(carry into the group, all strip maxima)
        stg64_relaxed_if(pub_out, cout_g + (i - 1), pack_tagged(max(cin, incl), tag));
        // ---- fold what lies to the left into this lane's values
        const int cfix = max(cin, excl);
#pragma unroll
        for (int k = 0; k < kScanC; k++) prev[k] = max(prev[k], cfix);
        carry_prev = cin;
        // ---- progress / back-pressure, every 8 rows: nobody runs more than 16 + 7 rows ahead of the last warp (which waits for everybody)
        if ((i & 7) == 0) {
            sts32_if(pub_prog, prog_s, i);
            if (w != WARPS - 1) {          // (warp-uniform)
                unsigned polls = 0;
                while (lds_volatile1(prog_s) < i - 16) { if (++polls > (1u << 24)) { atomicExch(a.err, 1); break; } }
            }
        }
    }
}

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS, kScanMaxWarps / WARPS) nw_scan_kernel(const ScanArgs a)
{
    constexpr int kScanWarps = WARPS, kScanT = 32 * WARPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_prog;                        // rows the last warp has completed (multiples of 8)
    __shared__ int s_group;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int n = a.n, S = a.S;
    // prof[letter][column of the strip] bytes s'(letter, x[column]); row S is all zero
    const unsigned prof_s = (unsigned)__cvta_generic_to_shared(smem_raw + (size_t)w * scan_warp_smem(S));      // (kScanWarps, kScanT: this instance's)
    const unsigned tot_s = (unsigned)__cvta_generic_to_shared(smem_raw + (size_t)kScanWarps * scan_warp_smem(S));      // [kScanRing][kScanWarps] (row | strip maximum)
    const unsigned gcin_s = tot_s + (unsigned)(kScanRing * kScanWarps * 8);                                            // [kScanRing] (row | carry into the group)
    const unsigned prog_s = (unsigned)__cvta_generic_to_shared(&s_prog);

    for (;;) {
        __syncthreads();                          // every warp is done with the previous group (rings, counters)
        if (tid == 0) s_group = atomicAdd(a.ticket, 1);
        __syncthreads();
        const int lg = s_group;                   // local group
        if (lg >= a.nchunks) break;
        const int gg = a.chunk0 + lg;             // global group
        const long long c0 = ((long long)gg * kScanWarps + w) * kScanStrip + (long long)lane * kScanC;      // this lane's first column
        const bool to_global = w == kScanWarps - 1 && gg + 1 < a.total_chunks;
        const unsigned long long* cin_g = a.carry + (long long)lg * n;                       // written by the group to the left
        unsigned long long* cout_g = (lg + 1 == a.nchunks && a.peer_carry0 != nullptr) ? a.peer_carry0 : a.carry + (long long)(lg + 1) * n;

        // ---- the rings start out empty (row tags 0; rows count from 1), the profile of this strip's columns is built
        for (int k = tid; k < kScanRing * (kScanWarps + 1); k += kScanT) sts_volatile64(tot_s + 8u * (unsigned)k, 0ull);
        if (tid == 0) s_prog = 0;
        {
            unsigned xl[kScanC];
#pragma unroll
            for (int k = 0; k < kScanC; k++) {
                const long long j = c0 + k;
                const unsigned v = (j < a.m) ? (unsigned)__ldg(a.x + j) : (unsigned)S;
                xl[k] = min(v, (unsigned)S);       // (letters were validated at upload)
            }
            for (int yl = 0; yl <= S; yl++) {
                unsigned wd[4] = {0u, 0u, 0u, 0u};
                if (yl < S) {
#pragma unroll
                    for (int k = 0; k < kScanC; k++) {
                        const unsigned b = xl[k] < (unsigned)S ? (unsigned)__ldg(a.sprime + yl * S + xl[k]) : 0u;
                        wd[k >> 2] |= b << (8 * (k & 3));
                    }
                }
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(prof_s + (unsigned)yl * kScanStrip + 16u * (unsigned)lane),
                             "r"(wd[0]), "r"(wd[1]), "r"(wd[2]), "r"(wd[3]) : "memory");
            }
        }
        __syncthreads();                          // rings and counters are reset before anybody writes into them
        unsigned long long* dbg = (a.dbg != nullptr && gg < 8 && lane == 0) ? a.dbg + (gg * kScanWarps + w) * 8 : nullptr;
        if (dbg) dbg[0] = globaltimer_ns();

        int prev[kScanC];                          // P[i-1][this lane's columns]
#pragma unroll
        for (int k = 0; k < kScanC; k++) prev[k] = 0;
        const unsigned prof_lane_s = prof_s + 16u * (unsigned)lane;
        if (gg == 0) scan_rows<false, false, WARPS>(a, lane, w, prof_lane_s, tot_s, gcin_s, prog_s, cin_g, cout_g, to_global, prev);
        else if (w == 0) scan_rows<true, true, WARPS>(a, lane, w, prof_lane_s, tot_s, gcin_s, prog_s, cin_g, cout_g, to_global, prev);
        else scan_rows<false, true, WARPS>(a, lane, w, prof_lane_s, tot_s, gcin_s, prog_s, cin_g, cout_g, to_global, prev);
        if (dbg) dbg[5] = globaltimer_ns();
        // ---- the score lives at column m of the last row
        if (a.m - 1 >= c0 && a.m - 1 < c0 + kScanC) {
            const int k = (int)(a.m - 1 - c0);
            int v = 0;
#pragma unroll
            for (int q = 0; q < kScanC; q++) if (q == k) v = prev[q];
            *a.score = pack_tagged(v, a.tag);
        }
    }
}

}  // namespace nwb
