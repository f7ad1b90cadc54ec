// nw_batch3.cuh -- many independent short pairs, TWO pairs per warp in packed 16-bit halves (BASELINE config 3), round 2.
//
// Same sweep as nw_batch2.cuh (one band of 32*R rows per pair, lanes one column apart, every register = cell of pair A in its low and
// cell of pair B in its high 16 bits, ONE VIMNMX3.U16x2 and ONE rotate-shuffle per packed row), rebuilt around the instruction count
// of a step, which is what bounds it (ncu, profiles/r1z_ncu_batch2_*.txt: issue slots 76 % busy, 33.2 warp instructions per step of 16
// cells per lane of which 24 were the cells):
//
//  (a) ONE fma-pipe instruction adds BOTH substitution terms of a packed cell.  A byte permute interleaves the two profile words of a
//      step into (s'_A[r], 2 s'_B[r], s'_A[r+1], 2 s'_B[r+1]) -- one PRMT per TWO rows -- and IDP.2A with the constant halves
//      (1, 32768) computes diag + 1 * s'_A + 32768 * 2 s'_B = diag + s'_A + (s'_B << 16) (.LO: bytes 0, 1; .HI: bytes 2, 3).
//      Per 4 rows: 2 PRMT + 4 IDP.2A + 4 VIMNMX3.U16x2 = 2.5 instructions per packed cell (nw_batch2: 3) and no integer adds.
//      MQ of the R/4 row quads of a lane take this route, the others the two-IDP route of nw_batch2 (IDP.4A + IDP.2A, no permute):
//      a knob for the balance between the alu pipe (VIMNMX3, PRMT) and the fma pipe (IDP).
//  (b) The letter ring holds ONE 32-bit entry per column (the two profile-row offsets as 16-bit halves) instead of a 64-bit pair:
//      half the shared-memory traffic of the ring; the address of a profile word is again one instruction (IDP.2A.LO with the
//      constant bytes (1, 0) resp. (0, 1): base + offset half).
//  (c) Column letters arrive 128 columns at a time: a lane fetches FOUR letters per pair with one 32-bit load, validates them with
//      three word-wide operations, turns them into four ring entries with four permutes and stores them with one 16-byte store --
//      every fourth chunk, instead of a byte load + checks + two stores in EVERY chunk (55 -> ~12 instructions of overhead per chunk).
//  (d) The profile build reads the s' table 16 letters at a time (LDS.128 from 48-byte rows; pair B from a second, doubled table) and
//      addresses everything through 32-bit shared-space addresses: 131 -> ~45 instructions per 4 letters x 16 rows.
//
// Layout of a warp's shared memory:  [profile A: S rows][the zero row][profile B: S rows][ring: 256 + 32 entries], a profile row =
// 32 lanes x R bytes.  Pair A's letter S (= past the end / padding) is the zero row by construction; pair B's ring offsets are
// stored as (letter + 1) rows from the zero row, 0 = past the end.
#pragma once
#include "nw_batch.cuh"

namespace nwb {

// G = lanes that sweep one packed pair of pairs (a GROUP): 32 (one group per warp) or 16 (two groups per warp side by side: FOUR pairs
// per warp, R = 16 rows per lane for a 256-row band).  The narrow group halves the fill / drain of the systolic array (15 instead of
// 31 idle steps per sweep) and spreads the per-step overhead (ring load, two address computations, two profile loads, the multiply
// and the shuffle: ~7 instructions) over 16 packed rows instead of 8 -- 2.96 instead of 3.43 instructions per packed cell -- at the
// price of twice the shared memory per warp (the profiles of four pairs): 8 warps per SM instead of 16.
//
// SPLIT (G = 16, R = 16): with two warps per scheduler the 16 dependent VIMNMX3 of a lane's step (16 x 4.5 clk) are exposed.  The lane
// is therefore cut into an upper and a lower half of R / 2 rows, the lower half ONE COLUMN BEHIND the upper one: its inputs (the upper
// half's last row at that column and the one before) were computed a tick earlier, so a tick consists of TWO INDEPENDENT chains of
// R / 2 cells, issued alternately.  The bottom row of a lane is then one tick late: the lanes sit two columns apart
// (LAG = 2 (G - 1) + 1 = 31 ticks, the fill / drain of the 32-lane groups), the profile words of a column serve the upper half in the
// tick they arrive and the lower half in the next one.
//
// EXP (G = 32, R = 8): the alu pipe is what the cells saturate (VIMNMX3 + the merging PRMT: 1.5 alu instructions per packed cell;
// tools/pipe_microbench.cu: 58.7 cells/clk/SM for that mix, 68.4 when the merge is an integer add, which ptxas spreads over the alu
// and fma pipes).  The profiles are therefore stored EXPANDED, two rows per word with the other pair's bytes zero -- pair A
// (s'[r], 0, s'[r+1], 0), pair B (0, 2 s'[r], 0, 2 s'[r+1]) -- so that the merged word is wa + wb.  Twice the shared memory per warp.
template <int R, int K = 1, int G = 32, bool SPLIT = false, bool EXP = false>
struct Sched3 {
    static_assert(R == 4 || R == 8 || R == 16, "rows per lane");
    static_assert(K == 1 || K == 2, "lane skew in columns");
    static_assert(G == 32 || G == 16, "lanes per group");
    static constexpr int NG = 32 / G;                  // groups per warp
    static constexpr int By = G * R;
    static_assert(!EXP || !SPLIT, "one experiment at a time");
    static constexpr int WA = EXP ? R / 2 : R / 4;     // words per lane and letter and pair (bytes: s' for pair A, 2*s' for pair B)
    static constexpr int STRIDE = 128 * WA;            // bytes between the profile rows of two letters (all 32 lanes of the warp)
    static_assert(!SPLIT || (K == 1 && R >= 8), "split lanes: the shuffle feeds the next tick");
    static constexpr int KS = SPLIT ? 2 : K;           // columns between two lanes
    static constexpr int LAG = (G - 1) * KS + (SPLIT ? 1 : 0);      // K = 2: the shuffle of a step is issued one step early (off the dependent chain); twice the fill / drain
    static constexpr int CH = G;                       // steps per unrolled chunk
    static constexpr int XR = 256, XM = CH;            // letter ring of a group (32-bit entries) + mirror of its first CH entries
    static constexpr int BLK = 4 * G;                  // columns fetched at a time (4 per lane of the group)
    __host__ __device__ static constexpr int nlc(int m) { return (m + LAG + CH - 1) / CH; }
    // (the profile build writes letters in groups of four: up to three rows past pair B's last letter land on the ring, which must cover them)
    static constexpr size_t RING_BYTES = (size_t)NG * (XR + XM) * 4 > 3 * (size_t)STRIDE ? (size_t)NG * (XR + XM) * 4 : 3 * (size_t)STRIDE;
    __host__ __device__ static constexpr size_t warp_smem_bytes(int S) { return (size_t)(2 * S + 1) * STRIDE + RING_BYTES; }
};

constexpr int kBatch3MaxLetters = 31;      // rows of the CTA's s' tables (static shared memory)
constexpr int kBatch3MaxSprime = 127;      // 2 * s' must fit a byte, and 256 * s' must stay below 2^15
constexpr int kB3Pitch = 12;               // words per s' table row: 16-byte aligned; 3 (mod 8) quads, so the rows of a quarter warp spread over the banks

__device__ __forceinline__ unsigned b3_u4c(const uint4& v, int j) { return j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w; }

// prmt.b32 with the full 4-bit selector nibbles (bit 3 = replicate the sign bit of the selected byte); __byte_perm masks them to 3 bits
__device__ __forceinline__ unsigned b3_prmt(unsigned a, unsigned b, unsigned sel)
{
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__device__ __forceinline__ uint4 b3_lds128(unsigned addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// ---- 5-bit packed letters (PK instances; include/nwb200.h, nwb200_align_batch_packed5): letter k of a sequence sits in bits
// [5k, 5k + 5) of the little-endian bit stream that starts at the sequence's BYTE offset in the pool (8 letters in 5 bytes).  The
// pool is allocated with slack, so aligned 32-bit loads may run a few bytes past a sequence's end.
// 32 bits of the stream starting at bit `bit` (any alignment)
__device__ __forceinline__ unsigned b3_bits32(const uint8_t* __restrict__ base, long long bit)
{
    const uint8_t* p = base + (bit >> 3);
    const unsigned long long ad = reinterpret_cast<unsigned long long>(p);
    const unsigned* w = reinterpret_cast<const unsigned*>(ad & ~3ull);
    const unsigned sh = (unsigned)(ad & 3ull) * 8u + (unsigned)(bit & 7);
    const unsigned w0 = __ldg(w), w1 = __ldg(w + 1);
    return __funnelshift_r(w0, w1, sh);              // sh <= 31
}
// four 5-bit letters (the low 20 bits of v) -> one letter per byte
__device__ __forceinline__ unsigned b3_unpack4(unsigned v)
{
    return (v & 31u) | ((v << 3) & 0x1f00u) | ((v << 6) & 0x1f0000u) | ((v << 9) & 0x1f000000u);
}

// The R row letters of a lane, fetched ahead of their use (for the NEXT pair of a warp while the current one is swept): a lane whose
// R = 8 rows are all inside the sequence and 8-byte aligned fetches them with one load (`fast`); other lanes fetch when they build.
struct B3Rows { uint4 v; bool fast; };
template <int R, bool PK = false>
__device__ __forceinline__ B3Rows b3_fetch_rows(const uint8_t* __restrict__ y, int i0, int n)
{
    B3Rows q;
    q.v = make_uint4(0u, 0u, 0u, 0u);
    q.fast = false;
    if constexpr (PK) {                              // any alignment: R letters = 5 R bits from bit 5 i0
        static_assert(!PK || R == 8, "packed letters: 8 rows per lane");
        q.fast = i0 >= 0 && i0 + 8 <= n;
        if (q.fast) {
            const unsigned lo = b3_bits32(y, 5LL * i0), hi = b3_bits32(y, 5LL * i0 + 20);
            q.v.x = b3_unpack4(lo); q.v.y = b3_unpack4(hi);
        }
        return q;
    }
    if constexpr (R == 8) {
        q.fast = i0 >= 0 && i0 + 8 <= n && ((reinterpret_cast<unsigned long long>(y + i0) & 7ull) == 0ull);
        if (q.fast) { const uint2 t = __ldg(reinterpret_cast<const uint2*>(y + i0)); q.v.x = t.x; q.v.y = t.y; }
    }
    if constexpr (R == 16) {
        q.fast = i0 >= 0 && i0 + 16 <= n && ((reinterpret_cast<unsigned long long>(y + i0) & 15ull) == 0ull);
        if (q.fast) q.v = __ldg(reinterpret_cast<const uint4*>(y + i0));
    }
    return q;
}

// Row offsets (bytes into an s' table) of a lane's R matrix rows; row S = the zero row for padding rows.  Returns true when a letter is >= S.
template <int R, bool PK = false>
__device__ __forceinline__ bool b3_table_rows(unsigned* row_off, int S, const uint8_t* __restrict__ y, int i0, int n, const B3Rows& pre)
{
    unsigned yl[R];
    if (pre.fast) {
#pragma unroll
        for (int r = 0; r < R; r++) yl[r] = __byte_perm(r < 4 ? pre.v.x : r < 8 ? pre.v.y : r < 12 ? pre.v.z : pre.v.w, 0u, 0x4440u + (unsigned)(r & 3));
    } else {
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = i0 + r;
            if constexpr (PK) yl[r] = (i >= 0 && i < n) ? (b3_bits32(y, 5LL * i) & 31u) : kPastEnd;
            else yl[r] = (i >= 0 && i < n) ? (unsigned)__ldg(y + i) : kPastEnd;
        }
    }
    bool bad = false;
#pragma unroll
    for (int r = 0; r < R; r++) {          // a real letter must be < S; S is the internal code of the zero row
        bad |= yl[r] >= (unsigned)S && yl[r] != kPastEnd;
        row_off[r] = min(yl[r], (unsigned)S) * (unsigned)(kB3Pitch * 4);
    }
    return bad;
}

// One byte profile: prof[letter][lane][q] = bytes t(y[row 4q..4q+3], letter) for the table t at shared address tab_s.  Letters are
// written in whole groups of four: up to three profile rows past letter S - 1 are scribbled on (see the caller for who owns them).
// Rows are handled eight at a time (R = 16: two passes, the second one fills the upper 8 bytes of a lane's 16).
template <int R>
__device__ __forceinline__ void b3_build_profile(unsigned prof_lane_s, unsigned tab_s, int S, const unsigned (&row_off)[R])
{
    constexpr int WA = R / 4;
    constexpr unsigned STRIDE = 128 * WA;
    constexpr int RO = R < 8 ? R : 8;                         // rows per pass
    constexpr int WO = RO / 4;                                // words per pass
#pragma unroll
    for (int o = 0; o < R / RO; o++) {
#pragma unroll
        for (int half = 0; half < 2; half++) {                    // letters 16*half .. 16*half + 15 (S <= 31)
            if (16 * half < S) {
                uint4 v[RO];
#pragma unroll
                for (int r = 0; r < RO; r++) v[r] = b3_lds128(tab_s + row_off[RO * o + r] + 16u * (unsigned)half);
                const unsigned dst = prof_lane_s + (unsigned)(16 * half) * STRIDE + (unsigned)(RO * o);      // bytes RO*o .. of the lane's R
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (16 * half + 4 * j < S) {
                        unsigned ow[WO][4];
#pragma unroll
                        for (int q = 0; q < WO; q++) {
                            const unsigned w0 = b3_u4c(v[4 * q], j), w1 = b3_u4c(v[4 * q + 1], j), w2 = b3_u4c(v[4 * q + 2], j), w3 = b3_u4c(v[4 * q + 3], j);
                            const unsigned t0 = __byte_perm(w0, w1, 0x5140), t1 = __byte_perm(w2, w3, 0x5140);
                            const unsigned t2 = __byte_perm(w0, w1, 0x7362), t3 = __byte_perm(w2, w3, 0x7362);
                            ow[q][0] = __byte_perm(t0, t1, 0x5410); ow[q][1] = __byte_perm(t0, t1, 0x7632);
                            ow[q][2] = __byte_perm(t2, t3, 0x5410); ow[q][3] = __byte_perm(t2, t3, 0x7632);
                        }
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const unsigned at = dst + (unsigned)(4 * j + k) * STRIDE;
                            if constexpr (WO == 2) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(at), "r"(ow[0][k]), "r"(ow[1][k]) : "memory");
                            else asm volatile("st.shared.u32 [%0], %1;" ::"r"(at), "r"(ow[0][k]) : "memory");
                        }
                    }
                }
            }
        }
    }
}

// The expanded profile (Sched3<..., EXP>): prof[letter][lane][j] = (t(y[row 2j], letter), 0, t(y[row 2j+1], letter), 0) for pair A
// (HIGH = false), the same bytes one position up for pair B.  One permute per output word, straight from two table rows: selector
// nibbles with bit 3 set replicate the sign bit of a byte, which is 0 for table entries < 128.
template <int R, bool HIGH>
__device__ __forceinline__ void b3_build_profile_exp(unsigned prof_lane_s, unsigned tab_s, int S, const unsigned (&row_off)[R])
{
    constexpr int WP = R / 2;
    constexpr unsigned STRIDE = 128 * WP;
    static_assert(R == 8, "expanded profiles: 16 bytes per lane and letter");
#pragma unroll
    for (int half = 0; half < 2; half++) {                    // letters 16*half .. 16*half + 15 (S <= 31)
        if (16 * half < S) {
            uint4 v[R];
#pragma unroll
            for (int r = 0; r < R; r++) v[r] = b3_lds128(tab_s + row_off[r] + 16u * (unsigned)half);
            const unsigned dst = prof_lane_s + (unsigned)(16 * half) * STRIDE;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (16 * half + 4 * j < S) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const unsigned z = 8u | (unsigned)k;                                       // a zero byte
                        const unsigned sel = HIGH ? (z | ((unsigned)k << 4) | (z << 8) | ((4u + (unsigned)k) << 12))
                                                  : ((unsigned)k | (z << 4) | ((4u + (unsigned)k) << 8) | (z << 12));
                        unsigned o[WP];
#pragma unroll
                        for (int q = 0; q < WP; q++) o[q] = b3_prmt(b3_u4c(v[2 * q], j), b3_u4c(v[2 * q + 1], j), sel);
                        const unsigned at = dst + (unsigned)(4 * j + k) * STRIDE;
                        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(at), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                    }
                }
            }
        }
    }
}

template <int R, int WARPS, int MQ, int K = 1, int G = 32, bool SPLIT = false, bool EXP = false, bool PK = false>
__global__ void __launch_bounds__(WARPS * 32) nw_batch3_kernel(const BatchArgs a)
{
    using S3 = Sched3<R, K, G, SPLIT, EXP>;
    constexpr int By = S3::By, WA = S3::WA, XR = S3::XR, XM = S3::XM, BLK = S3::BLK, CH = S3::CH, NG = S3::NG, KS = S3::KS;
    static_assert(!(SPLIT || EXP) || MQ == R / 4, "split lanes / expanded profiles use the merged route for every row");
    constexpr unsigned STRIDE = S3::STRIDE;
    static_assert(MQ >= 0 && MQ <= R / 4, "row quads on the merged route");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(16) unsigned tabA[(kBatch3MaxLetters + 1) * kB3Pitch];      // s'
    __shared__ __align__(16) unsigned tabB[(kBatch3MaxLetters + 1) * kB3Pitch];      // 2 * s' (pair B's IDP.2A multiplies by 32768)
    stage_sprime(tabA, a.sprime, a.S, smem_raw, (kBatch3MaxLetters + 1) * kB3Pitch, kB3Pitch);
    for (int i = threadIdx.x; i < (kBatch3MaxLetters + 1) * kB3Pitch; i += blockDim.x) tabB[i] = tabA[i] << 1;      // s' <= 127: no carry between bytes
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int gl = lane & (G - 1), gi = lane / G;                               // lane within its group, group within the warp
    const int S = a.S;
    const unsigned tabA_s = (unsigned)__cvta_generic_to_shared(tabA), tabB_s = (unsigned)__cvta_generic_to_shared(tabB);
    const unsigned base_s = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)w * (unsigned)S3::warp_smem_bytes(S);
    const unsigned laneA_s = base_s + lane * 4 * WA;                            // profile A, this lane's words; row S = the zero row
    const unsigned zero_s = laneA_s + (unsigned)S * STRIDE;                    // = pair B's base: row 0 of B's offsets is the zero row
    const unsigned laneB_s = zero_s + STRIDE;                                  // profile B, letter 0
    const unsigned ring_s = base_s + (unsigned)(2 * S + 1) * STRIDE + (unsigned)gi * (unsigned)((XR + XM) * 4);      // this group's [XR + XM] 32-bit entries
    const int src_lane = (lane + 31) & 31;
    const bool last = gl == G - 1;
    const unsigned ezero = (unsigned)S * STRIDE;                               // ring entry of a column outside both pairs
    // ring entries are built as (letter << 8) halves and shifted to letter * STRIDE
    constexpr int SHR = (STRIDE == 128) ? 1 : 0, SHL = (STRIDE == 512) ? 1 : 0;

    // Tickets and per-pair metadata run ahead of the sweep (as in nw_batch2.cuh).  A ticket is a unit of NG packed pairs of pairs; a
    // lane carries the metadata of its own group's two pairs.
    struct Meta { unsigned long long p0, pA, oyA, oxA, oyB, oxB; unsigned nA, mA, nB, mB; };
    auto load_meta = [&](unsigned long long t) {
        Meta q;
        q.p0 = a.first + (unsigned long long)(2 * NG) * t;                     // first pair of the unit (warp-uniform)
        q.pA = q.p0 + 2ull * (unsigned)gi;
        q.oyA = q.oxA = q.oyB = q.oxB = 0; q.nA = q.mA = q.nB = q.mB = 0;
        if (q.pA < a.npairs) {
            q.nA = a.lenY[q.pA]; q.mA = a.lenX[q.pA]; q.oyA = a.offY[q.pA]; q.oxA = a.offX[q.pA];
            if (q.pA + 1 < a.npairs) { q.nB = a.lenY[q.pA + 1]; q.mB = a.lenX[q.pA + 1]; q.oyB = a.offY[q.pA + 1]; q.oxB = a.offX[q.pA + 1]; }
        }
        return q;
    };
    // the row letters of a pair are requested as soon as its metadata is known: one pair ahead of the sweep
    auto fetch_rows = [&](unsigned long long off, unsigned n) {
        const int nn = n > (unsigned)By ? 0 : (int)n;             // taller than the band: swept as an empty pair
        return b3_fetch_rows<R, PK>(a.letters + off, gl * R - (By - nn), nn);
    };
    unsigned long long tk = 0;                       // lane 0: the ticket drawn ahead
    if (lane == 0) tk = atomicAdd(a.ticket, 1ull);
    Meta nx = load_meta(__shfl_sync(kFull, tk, 0));
    B3Rows nrA = fetch_rows(nx.oyA, nx.nA), nrB = fetch_rows(nx.oyB, nx.nB);
    if (lane == 0) tk = atomicAdd(a.ticket, 1ull);
    unsigned keep;                                   // 0 in the last lane of a group, 1 elsewhere -- opaque to the compiler, which would turn the multiply back into a select
    asm volatile("mov.u32 %0, %1;" : "=r"(keep) : "r"(last ? 0u : 1u));
    unsigned one;
    asm volatile("mov.u32 %0, 1;" : "=r"(one));

    for (;;) {
        const Meta cu = nx;
        const B3Rows crA = nrA, crB = nrB;
        const unsigned long long pA = cu.pA, pB = pA + 1;
        if (cu.p0 >= a.npairs) break;
        const bool hasA = pA < a.npairs, hasB = pB < a.npairs;
        int nA = (int)cu.nA, mA = (int)cu.mA, nB = (int)cu.nB, mB = (int)cu.mB;
        const int gapsA = (nA + mA) * a.gap, gapsB = (nB + mB) * a.gap;
        const bool tallA = nA > By, tallB = nB > By;
        if (tallA) { nA = 0; mA = 0; }                     // swept as an empty pair; the host re-runs it as a single pair
        if (tallB) { nB = 0; mB = 0; }
        if (nA == 0) mA = 0;
        if (nB == 0) mB = 0;
        const uint8_t* yA = a.letters + cu.oyA;
        const uint8_t* xA = a.letters + cu.oxA;
        const uint8_t* yB = a.letters + cu.oyB;
        const uint8_t* xB = a.letters + cu.oxB;
        int m = max(mA, mB);
        if constexpr (NG == 2) m = max(m, __shfl_xor_sync(kFull, m, 16));       // the sweep's length is the warp's
        // a lane fetches the four letters of columns c0 .. c0 + 3 with one load when the sequence start is 4-byte aligned (the letter
        // pool is allocated with slack, so the load may run past the end of the sequence; such letters are masked out below)
        const bool al = (((reinterpret_cast<unsigned long long>(xA) | reinterpret_cast<unsigned long long>(xB)) & 3ull) == 0ull);
        auto fetch4 = [&](const uint8_t* x, int mm, int c0) -> unsigned {
            if (c0 >= mm) return 0u;
            if constexpr (PK) return b3_unpack4(b3_bits32(x, 5LL * c0));      // (letters past the end are masked out below)
            if (al) return __ldg(reinterpret_cast<const unsigned*>(x + c0));
            unsigned v = 0u;
#pragma unroll
            for (int k = 0; k < 4; k++) if (c0 + k < mm) v |= (unsigned)__ldg(x + c0 + k) << (8 * k);
            return v;
        };
        // four ring entries from the letters of columns c0 .. c0 + 3 (wa, wb: one letter per byte): validated, past-the-end columns
        // mapped to the zero row, stored with one 16-byte store (+ the mirror of the ring's first CH entries)
        auto put_block = [&](int c0, unsigned wa, unsigned wb) {
            // letters of this word inside the sequences: >= 4 all, <= 0 none
            const int ka = mA - c0, kb = mB - c0;
            const unsigned keepA = ka >= 4 ? 0xffffffffu : (ka <= 0 ? 0u : (0xffffffffu >> (8 * (4 - ka))));
            const unsigned keepB = kb >= 4 ? 0xffffffffu : (kb <= 0 ? 0u : (0xffffffffu >> (8 * (4 - kb))));
            wa &= keepA; wb &= keepB;                                                // the real letters, zeros elsewhere
            // validation, word-wide: a byte v < 0x80 has bit 7 of v + (0x80 - S) set exactly when v >= S; bytes >= 0x80 flag themselves
            const unsigned KV = (unsigned)(0x80 - S) * 0x01010101u;
            if ((((wa | (wa + KV)) | (wb | (wb + KV))) & 0x80808080u) != 0u) { *a.err = 1; wa = 0u; wb = 0u; }      // (keeps the offsets inside the profile)
            // columns past the end of a sequence: letter S for pair A (the zero row by layout), 0xff for pair B (stored + 1 = 0 = the zero row)
            wa |= ((unsigned)S * 0x01010101u) & ~keepA;
            wb |= ~keepB;
            const unsigned wbp = ((wb & 0x7f7f7f7fu) + 0x01010101u) ^ (wb & 0x80808080u);      // bytewise + 1 (mod 256)
            unsigned e[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                // (0, letter A, 0, letter B + 1): selector nibbles with bit 3 set replicate the sign bit of a byte (0 for values < 128)
                const unsigned sel = (8u | (unsigned)k) | ((unsigned)k << 4) | ((8u | (unsigned)k) << 8) | ((4u + (unsigned)k) << 12);
                e[k] = (b3_prmt(wa, wbp, sel) >> SHR) << SHL;
            }
            const unsigned p = (unsigned)c0 & (unsigned)(XR - 1);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(ring_s + 4u * p), "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]) : "memory");
            if (p < (unsigned)XM)
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(ring_s + 4u * (p + XR)), "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]) : "memory");
        };
        __syncwarp();
        // ---- the first two letter blocks are requested before the profiles are built and land under the build
        unsigned fa0 = fetch4(xA, mA, 4 * gl), fb0 = fetch4(xB, mB, 4 * gl);
        unsigned fa1 = fetch4(xA, mA, BLK + 4 * gl), fb1 = fetch4(xB, mB, BLK + 4 * gl);
        // ---- the two byte profiles; rows are aligned to the bottom of the band per pair.  Pair A first: its last letter group
        //      scribbles on the zero row and pair B's first rows; pair B's on the start of the ring; zero row and ring come last.
        {
            unsigned ro[R];
            const bool badA = b3_table_rows<R, PK>(ro, S, yA, gl * R - (By - nA), nA, crA);
            if constexpr (EXP) b3_build_profile_exp<R, false>(laneA_s, tabA_s, S, ro);
            else b3_build_profile<R>(laneA_s, tabA_s, S, ro);
            const bool badB = b3_table_rows<R, PK>(ro, S, yB, gl * R - (By - nB), nB, crB);
            if constexpr (EXP) b3_build_profile_exp<R, true>(laneB_s, tabB_s, S, ro);
            else b3_build_profile<R>(laneB_s, tabB_s, S, ro);
            if (badA || badB) *a.err = 1;
            if constexpr (WA == 4) asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(zero_s), "r"(0u) : "memory");
            else if constexpr (WA == 2) asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(zero_s), "r"(0u) : "memory");
            else asm volatile("st.shared.u32 [%0], %1;" ::"r"(zero_s), "r"(0u) : "memory");
        }
        __syncwarp();
        // ---- letter ring: columns -G KS .. -1 are outside (zero row), then block 0
#pragma unroll
        for (int k = 1; k <= KS; k++) asm volatile("st.shared.u32 [%0], %1;" ::"r"(ring_s + 4u * (unsigned)(XR - G * k + gl)), "r"(ezero) : "memory");
        put_block(4 * gl, fa0, fb0);
        __syncwarp();
        unsigned h[R];
#pragma unroll
        for (int r = 0; r < R; r++) h[r] = 0u;
        unsigned dprev = 0u, up_next = 0u;
        unsigned hmid_prev = 0u;                     // SPLIT: the upper half's last row, one column behind the lower half
        unsigned lwa[WA], lwb[WA];                   // SPLIT: the profile words of the previous tick's column (the lower half's column)
#pragma unroll
        for (int q = 0; q < WA; q++) { lwa[q] = 0u; lwb[q] = 0u; }
        const int nlc = S3::nlc(m);
        nx = load_meta(__shfl_sync(kFull, tk, 0));
        if (lane == 0) tk = atomicAdd(a.ticket, 1ull);
        nrA = fetch_rows(nx.oyA, nx.nA); nrB = fetch_rows(nx.oyB, nx.nB);
        for (int lc = 0; lc < nlc; lc++) {
            // Shared-space addresses (32-bit) for the loads of the chunk: ring entries are requested three steps, profile words two
            // steps ahead of their use.
            const unsigned xs_s = ring_s + 4u * (unsigned)((CH * lc - KS * gl) & (XR - 1));
            unsigned xo[CH];
            unsigned wa[CH][WA], wb[CH][WA];
            auto load_xo = [&](int s) {
                if (s < CH) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(xo[s]) : "r"(xs_s + 4u * (unsigned)s));
            };
            auto load_pw = [&](int s) {
                if (s >= CH) return;
                const unsigned qa = __dp2a_lo(xo[s], 0x00000001u, laneA_s);        // base + low half of the entry
                const unsigned qb = __dp2a_lo(xo[s], 0x00000100u, zero_s);         // base + high half
                if constexpr (WA == 1) {
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wa[s][0]) : "r"(qa));
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wb[s][0]) : "r"(qb));
                } else if constexpr (WA == 2) {
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(wa[s][0]), "=r"(wa[s][1]) : "r"(qa));
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(wb[s][0]), "=r"(wb[s][1]) : "r"(qb));
                } else {
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(wa[s][0]), "=r"(wa[s][1]), "=r"(wa[s][2]), "=r"(wa[s][3]) : "r"(qa));
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(wb[s][0]), "=r"(wb[s][1]), "=r"(wb[s][2]), "=r"(wb[s][3]) : "r"(qb));
                }
            };
#pragma unroll
            for (int s = 0; s < 3; s++) load_xo(s);
#pragma unroll
            for (int s = 0; s < 2; s++) load_pw(s);
#pragma unroll
            for (int s = 0; s < CH; s++) {
                load_xo(s + 3);
                load_pw(s + 2);
                // rotate-shuffle: the bottom row of the lane above; row 0 of P (above the first lane of a group) is zero.  K = 2: the value
                // shuffled now is consumed in the NEXT step (the lane above is two columns ahead), so the shuffle's latency is off the dependent chain
                unsigned up;
                if constexpr (K == 1) up = __shfl_sync(kFull, h[R - 1] * keep, src_lane);        // (a multiply: the fma pipe has room, the alu pipe has not)
                else { up = up_next; up_next = __shfl_sync(kFull, h[R - 1] * keep, src_lane); }
                unsigned diag = dprev;
                dprev = up;
                if constexpr (EXP) {
                    unsigned mg[WA];
#pragma unroll
                    for (int j = 0; j < WA; j++) {                                 // (A[2j], 2B[2j], A[2j+1], 2B[2j+1]): disjoint bytes, no carries
                        if (j & 1) mg[j] = wa[s][j] * one + wb[s][j];              // fma pipe (IMAD; `one` is opaque to the compiler)
                        else mg[j] = wa[s][j] | wb[s][j];                          // alu pipe (LOP3)
                    }
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const unsigned left = h[r];
                        const unsigned tt = (r & 1) ? __dp2a_hi(0x80000001u, mg[r >> 1], diag) : __dp2a_lo(0x80000001u, mg[r >> 1], diag);
                        const unsigned nv = __vimax3_u16x2(tt, up, left);
                        diag = left; up = nv; h[r] = nv;
                    }
                } else if constexpr (!SPLIT) {
                    unsigned mg[R / 4][2];
#pragma unroll
                    for (int q = 0; q < MQ; q++) {
                        mg[q][0] = __byte_perm(wa[s][q], wb[s][q], 0x5140);        // (A[4q], 2B[4q], A[4q+1], 2B[4q+1])
                        mg[q][1] = __byte_perm(wa[s][q], wb[s][q], 0x7362);        // (A[4q+2], 2B[4q+2], A[4q+3], 2B[4q+3])
                    }
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const unsigned left = h[r];
                        unsigned tt;
                        if ((r >> 2) < MQ) {
                            // diag + 1 * s'_A + 32768 * (2 s'_B): both halves in one IDP.2A
                            const unsigned m2 = mg[r >> 2][(r >> 1) & 1];
                            tt = (r & 1) ? __dp2a_hi(0x80000001u, m2, diag) : __dp2a_lo(0x80000001u, m2, diag);
                        } else {
                            tt = __dp4a(wa[s][r >> 2], 1u << (8 * (r & 3)), diag);                          // low half += s'_A
                            const unsigned half = (r & 1) ? 0x80000000u : 0x00008000u;                      // 32768 * (2 s'_B) = s'_B << 16
                            tt = (r & 2) ? __dp2a_hi(half, wb[s][r >> 2], tt) : __dp2a_lo(half, wb[s][r >> 2], tt);
                        }
                        const unsigned nv = __vimax3_u16x2(tt, up, left);
                        diag = left; up = nv; h[r] = nv;
                    }
                } else {
                    constexpr int RH = R / 2, WH = WA / 2;
                    // the lower half works on the previous tick's column: its top inputs are the upper half's last row as it stands now
                    // (that column) and as it stood a tick ago (the column before)
                    unsigned up_l = h[RH - 1], diag_l = hmid_prev;
                    hmid_prev = up_l;
                    unsigned mgu[WH][2], mgl[WH][2];
#pragma unroll
                    for (int q = 0; q < WH; q++) {
                        mgu[q][0] = __byte_perm(wa[s][q], wb[s][q], 0x5140);
                        mgu[q][1] = __byte_perm(wa[s][q], wb[s][q], 0x7362);
                        mgl[q][0] = __byte_perm(lwa[WH + q], lwb[WH + q], 0x5140);
                        mgl[q][1] = __byte_perm(lwa[WH + q], lwb[WH + q], 0x7362);
                    }
#pragma unroll
                    for (int r = 0; r < RH; r++) {                       // two independent chains, issued alternately
                        {
                            const unsigned left = h[r];
                            const unsigned m2 = mgu[r >> 2][(r >> 1) & 1];
                            const unsigned tt = (r & 1) ? __dp2a_hi(0x80000001u, m2, diag) : __dp2a_lo(0x80000001u, m2, diag);
                            const unsigned nv = __vimax3_u16x2(tt, up, left);
                            diag = left; up = nv; h[r] = nv;
                        }
                        {
                            const unsigned left = h[RH + r];
                            const unsigned m2 = mgl[r >> 2][(r >> 1) & 1];
                            const unsigned tt = (r & 1) ? __dp2a_hi(0x80000001u, m2, diag_l) : __dp2a_lo(0x80000001u, m2, diag_l);
                            const unsigned nv = __vimax3_u16x2(tt, up_l, left);
                            diag_l = left; up_l = nv; h[RH + r] = nv;
                        }
                    }
#pragma unroll
                    for (int q = WH; q < WA; q++) { lwa[q] = wa[s][q]; lwb[q] = wb[s][q]; }
                }
            }
            // ---- every fourth chunk: the next letter block goes into the ring (it was requested four chunks ago) and the one after is requested
            if ((lc & 3) == 2) {
                const int b = (lc >> 2) + 1;                                   // block b: columns BLK b .. BLK b + BLK - 1, read from chunk 4 b - 1 on
                __syncwarp();
                put_block(BLK * b + 4 * gl, fa1, fb1);
                fa1 = fetch4(xA, mA, BLK * (b + 1) + 4 * gl);
                fb1 = fetch4(xB, mB, BLK * (b + 1) + 4 * gl);
                __syncwarp();
            }
        }
        // the last row of a group's last lane is row n of both matrices, frozen behind their last columns: un-shift H = P + (n+m)*gap
        if (last) {
            if (hasA) a.scores[pA] = tallA ? kBatchTooTall : (int)(h[R - 1] & 0xffffu) + gapsA;
            if (hasB) a.scores[pB] = tallB ? kBatchTooTall : (int)(h[R - 1] >> 16) + gapsB;
        }
    }
}

}  // namespace nwb
