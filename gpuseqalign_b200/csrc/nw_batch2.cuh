// nw_batch2.cuh -- many independent short pairs, TWO pairs per warp in packed 16-bit halves (BASELINE config 3).
//
// Same systolic sweep as nw_batch.cuh (one band of 32*R rows per pair, lanes one column apart), but every register holds the
// cell of pair A in its low and the cell of pair B in its high 16 bits, so that ONE DPX instruction (VIMNMX3.U16x2) and ONE
// warp shuffle serve two cells.  In shifted coordinates P = H - (i+j)*gap >= 0 and P[i][j] <= min(i,j) * max(s'); with
// s' <= 127 and min(i,j) <= 256 that is <= 32 512, so neither half overflows and the packed add never carries from the low
// half into the high one.  (The host picks this kernel only when both bounds hold, nwb200_capi_batch.inc.)
//
// The two pairs have different row AND column letters, so the substitution term of a packed cell is two table lookups:
//   pair A: byte r of a packed per-lane profile word, added to the LOW half with IDP.4A (selector byte = 1), as in nw_sweep.cuh
//   pair B: byte r of a second per-lane profile word that holds 2*s', added to the HIGH half with IDP.2A: the 16-bit operand is
//           the constant 32768 in the half that selects the byte (32768 * 2*s' = s' << 16), .LO / .HI pick the byte pair
// Both adds run on the fma pipe, the packed 3-way max on the alu pipe: 3 instructions per 2 cells instead of 4.
// Ragged pairs need no extra code: rows are aligned to the bottom of the band per pair (padding rows have zero profile bytes),
// columns past the end of the shorter pair read the all-zero profile row and freeze that half.
#pragma once
#include "nw_batch.cuh"

namespace nwb {

template <int R>
struct Sched2 {
    static_assert(R == 4 || R == 8, "rows per lane");
    static constexpr int By = 32 * R;
    static constexpr int WA = R / 4;                   // words per lane and letter and pair (bytes: s' for pair A, 2*s' for pair B)
    static constexpr int STRIDE = 128 * WA;            // bytes between the profile rows of two letters
    static constexpr int LAG = 31;
    static constexpr int PD = 2;                       // letter groups prefetched ahead
    static constexpr int XR = 128, XM = 32;            // letter ring of (offset A, offset B) entries + mirror: a chunk reads columns
                                                       // 32*lc-31 .. 32*lc+31 and columns up to 32*lc+95 are written ahead
    __host__ __device__ static constexpr int nlc(int m) { return (m + LAG + 31) / 32; }
    __host__ __device__ static constexpr size_t warp_smem_bytes(int S)
    {
        return (size_t)(2 * S + 1) * STRIDE + (size_t)(XR + XM) * 8;      // profile A | profile B | the zero row they share | ring
    }
};

constexpr int kBatch2MaxLetters = 31;      // rows of the CTA's s' table (static shared memory): 16 warps of S = 25 fit one SM with it
constexpr int kBatch2MaxSprime = 127;      // 2 * s' must fit a byte, and 256 * s' must stay below 2^15

// Both packed byte profiles of a warp's two pairs in ONE pass over the letters: prof[letter][lane][q] = bytes s'(y[row 4q..4q+3], letter)
// (layout of build_profile, nw_sweep.cuh).  All 2R row letters are fetched up front (one global round trip), the words of a lane go
// out as one 8-byte store, and the two independent transposes interleave.  SHLB: bit q set = pair B's word q is stored doubled.
template <int R, int SHLB>
__device__ __forceinline__ void build_profiles2(unsigned char* profA_lane, unsigned char* profB_lane, const unsigned* __restrict__ sp_tab, int S,
                                                const uint8_t* __restrict__ yA, int i0A, int nA,
                                                const uint8_t* __restrict__ yB, int i0B, int nB, int* err, unsigned char* zero_lane)
{
    constexpr int WA = R / 4;
    constexpr unsigned STRIDE = 128 * WA;
    // letter S is the INTERNAL code of a padding row (the zero row of the s' table); a real letter must be < S
    unsigned ya[R], yb[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int ia = i0A + r, ib = i0B + r;
        ya[r] = (ia >= 0 && ia < nA) ? (unsigned)__ldg(yA + ia) : kPastEnd;
        yb[r] = (ib >= 0 && ib < nB) ? (unsigned)__ldg(yB + ib) : kPastEnd;
    }
    const unsigned* rowA[R];
    const unsigned* rowB[R];
    bool bad = false;
#pragma unroll
    for (int r = 0; r < R; r++) {
        if (ya[r] >= (unsigned)S) { bad |= ya[r] != kPastEnd; ya[r] = (unsigned)S; }
        if (yb[r] >= (unsigned)S) { bad |= yb[r] != kPastEnd; yb[r] = (unsigned)S; }
        rowA[r] = sp_tab + ya[r] * kSpPitch;
        rowB[r] = sp_tab + yb[r] * kSpPitch;
    }
    if (bad) *err = 1;
#pragma unroll 1
    for (int g4 = 0; 4 * g4 < S; g4++) {
        unsigned oa[WA][4], ob[WA][4];
#pragma unroll
        for (int q = 0; q < WA; q++) {
            {
                const unsigned w0 = rowA[4 * q][g4], w1 = rowA[4 * q + 1][g4], w2 = rowA[4 * q + 2][g4], w3 = rowA[4 * q + 3][g4];
                const unsigned t0 = __byte_perm(w0, w1, 0x5140), t1 = __byte_perm(w2, w3, 0x5140);
                const unsigned t2 = __byte_perm(w0, w1, 0x7362), t3 = __byte_perm(w2, w3, 0x7362);
                oa[q][0] = __byte_perm(t0, t1, 0x5410); oa[q][1] = __byte_perm(t0, t1, 0x7632);
                oa[q][2] = __byte_perm(t2, t3, 0x5410); oa[q][3] = __byte_perm(t2, t3, 0x7632);
            }
            {
                const unsigned w0 = rowB[4 * q][g4], w1 = rowB[4 * q + 1][g4], w2 = rowB[4 * q + 2][g4], w3 = rowB[4 * q + 3][g4];
                const unsigned t0 = __byte_perm(w0, w1, 0x5140), t1 = __byte_perm(w2, w3, 0x5140);
                const unsigned t2 = __byte_perm(w0, w1, 0x7362), t3 = __byte_perm(w2, w3, 0x7362);
                const int sh = (SHLB >> q) & 1;
                ob[q][0] = __byte_perm(t0, t1, 0x5410) << sh; ob[q][1] = __byte_perm(t0, t1, 0x7632) << sh;
                ob[q][2] = __byte_perm(t2, t3, 0x5410) << sh; ob[q][3] = __byte_perm(t2, t3, 0x7632) << sh;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const unsigned xl = 4u * (unsigned)g4 + (unsigned)j;
            if (xl < (unsigned)S) {
                if constexpr (WA == 2) {
                    *reinterpret_cast<uint2*>(profA_lane + xl * STRIDE) = make_uint2(oa[0][j], oa[1][j]);
                    *reinterpret_cast<uint2*>(profB_lane + xl * STRIDE) = make_uint2(ob[0][j], ob[1][j]);
                } else {
                    *reinterpret_cast<unsigned*>(profA_lane + xl * STRIDE) = oa[0][j];
                    *reinterpret_cast<unsigned*>(profB_lane + xl * STRIDE) = ob[0][j];
                }
            }
        }
    }
    if constexpr (WA == 2) *reinterpret_cast<uint2*>(zero_lane) = make_uint2(0u, 0u);
    else *reinterpret_cast<unsigned*>(zero_lane) = 0u;
}

__device__ __forceinline__ unsigned prmt_generic(unsigned a, unsigned b, unsigned sel)
{
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// NP: the last NP rows of a lane (whole profile words) take the substitution term through the alu pipe instead of the fma pipe:
// ONE byte permute builds (s'_A, 0, s'_B, 0) from the two profile words (selector nibbles with bit 3 set replicate the sign bit of
// a byte, which is 0 for s' <= 127), one IADD3 adds it.  The words of those rows hold plain s' for pair B as well.
template <int R, int WARPS, int NP = 0>
__global__ void __launch_bounds__(WARPS * 32) nw_batch2_kernel(const BatchArgs a)
{
    static_assert(NP % 4 == 0 && NP <= R, "whole profile words");
    constexpr int SHLB = ((1 << ((R - NP) / 4)) - 1);          // words of pair B that are stored doubled (IDP.2A rows)
    using S2 = Sched2<R>;
    constexpr int By = S2::By, WA = S2::WA, PD = S2::PD, XR = S2::XR, XM = S2::XM;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned sp_tab[(kBatch2MaxLetters + 1) * kSpPitch];
    stage_sprime(sp_tab, a.sprime, a.S, smem_raw, (kBatch2MaxLetters + 1) * kSpPitch);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int S = a.S;
    unsigned char* base = smem_raw + (size_t)w * S2::warp_smem_bytes(S);
    unsigned char* profA = base;                                     // [S][32 lanes][WA words]
    unsigned char* profB = base + (size_t)S * S2::STRIDE;            // [S][32][WA]; row S of B = row 2S of A = the zero row
    uint2* ring = reinterpret_cast<uint2*>(base + (size_t)(2 * S + 1) * S2::STRIDE);  // [XR + XM]
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    const unsigned laneA_s = (unsigned)__cvta_generic_to_shared(profA) + lane * 4 * WA;
    const unsigned laneB_s = (unsigned)__cvta_generic_to_shared(profB) + lane * 4 * WA;
    const int src_lane = (lane + 31) & 31;
    const bool last = lane == 31;

    auto put_letters = [&](int c, unsigned la, unsigned lb) {
        const int p = c & (XR - 1);
        const uint2 v = make_uint2((la < (unsigned)S ? la : 2u * (unsigned)S) * S2::STRIDE, lb * S2::STRIDE);      // letter S = the zero row
        ring[p] = v;
        if (p < XM) ring[p + XR] = v;
    };

    // Tickets and per-pair metadata run ahead of the sweep: the atomic for the warp's second-next ticket and the metadata loads of
    // its next ticket are issued before the chunk loop and land while it runs (three dependent global round trips otherwise).
    // (Tried and dropped: the profile build fully unrolled over the letter groups -- fewer instructions, but ptxas then caps the
    // kernel at 64 registers and the sweep loses 5 %.)
    struct Meta { unsigned long long pA, oyA, oxA, oyB, oxB; unsigned nA, mA, nB, mB; };
    auto load_meta = [&](unsigned long long t) {
        Meta q;
        q.pA = a.first + 2 * t;
        q.oyA = q.oxA = q.oyB = q.oxB = 0; q.nA = q.mA = q.nB = q.mB = 0;
        if (q.pA < a.npairs) {
            q.nA = a.lenY[q.pA]; q.mA = a.lenX[q.pA]; q.oyA = a.offY[q.pA]; q.oxA = a.offX[q.pA];
            if (q.pA + 1 < a.npairs) { q.nB = a.lenY[q.pA + 1]; q.mB = a.lenX[q.pA + 1]; q.oyB = a.offY[q.pA + 1]; q.oxB = a.offX[q.pA + 1]; }
        }
        return q;
    };
    unsigned long long tk = 0;                       // lane 0: the ticket drawn ahead
    if (lane == 0) tk = atomicAdd(a.ticket, 1ull);
    Meta nx = load_meta(__shfl_sync(kFull, tk, 0));
    if (lane == 0) tk = atomicAdd(a.ticket, 1ull);

    for (;;) {
        const Meta cu = nx;
        const unsigned long long pA = cu.pA, pB = pA + 1;
        if (pA >= a.npairs) break;
        const bool hasB = pB < a.npairs;
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
        const int m = max(mA, mB);
        __syncwarp();
        // ---- the first PD letter groups are requested before the profiles are built and land under the build
        auto fetch = [&](int c, unsigned& la, unsigned& lb) {
            la = c < mA ? (unsigned)__ldg(xA + c) : kPastEnd;
            lb = c < mB ? (unsigned)__ldg(xB + c) : kPastEnd;
        };
        auto check_put = [&](int c, unsigned la, unsigned lb) {        // a real letter must be < S; S itself is the internal zero-row code
            if (la >= (unsigned)S) { if (la != kPastEnd) *a.err = 1; la = (unsigned)S; }
            if (lb >= (unsigned)S) { if (lb != kPastEnd) *a.err = 1; lb = (unsigned)S; }
            put_letters(c, la, lb);
        };
        unsigned fa[PD], fb[PD];
#pragma unroll
        for (int g = 0; g < PD; g++) fetch(32 * g + lane, fa[g], fb[g]);
        // ---- the two byte profiles (pair B's IDP.2A words doubled); rows are aligned to the bottom of the band per pair
        build_profiles2<R, SHLB>(profA + lane * 4 * WA, profB + lane * 4 * WA, sp_tab, S,
                                 yA, lane * R - (By - nA), nA, yB, lane * R - (By - nB), nB, a.err, profB + (size_t)S * S2::STRIDE + lane * 4 * WA);
        // ---- letter ring: columns -32..-1 are outside (zero row), then the PD groups
        put_letters(-32 + lane, (unsigned)S, (unsigned)S);
#pragma unroll
        for (int g = 0; g < PD; g++) check_put(32 * g + lane, fa[g], fb[g]);
        __syncwarp();
        unsigned h[R];
#pragma unroll
        for (int r = 0; r < R; r++) h[r] = 0u;
        unsigned dprev = 0u;
        const int nlc = S2::nlc(m);
        nx = load_meta(__shfl_sync(kFull, tk, 0));
        if (lane == 0) tk = atomicAdd(a.ticket, 1ull);
        for (int lc = 0; lc < nlc; lc++) {
            const int cp = 32 * (lc + PD) + lane;
            unsigned pf_a, pf_b;
            fetch(cp, pf_a, pf_b);                                               // checked when it lands, after the chunk
            // Shared-space addresses (32-bit) for the loads of the chunk: letter offsets are requested three steps, profile words two
            // steps ahead of their use (ptxas still sinks the profile loads towards their use to save registers; the 16 warps of an
            // SM cover what is left of the shared-memory latency: short_sb is 6 % of the stall samples).
            const unsigned xs_s = ring_s + 8u * (unsigned)((32 * lc - lane) & (XR - 1));
            uint2 xo[32];
            unsigned wa[32][WA], wb[32][WA];
            auto load_xo = [&](int s) {
                if (s < 32) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(xo[s].x), "=r"(xo[s].y) : "r"(xs_s + 8u * (unsigned)s));
            };
            auto load_pw = [&](int s) {
                if (s >= 32) return;
                const unsigned qa = laneA_s + xo[s].x, qb = laneB_s + xo[s].y;
                if constexpr (WA == 1) {
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wa[s][0]) : "r"(qa));
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wb[s][0]) : "r"(qb));
                } else {
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(wa[s][0]), "=r"(wa[s][1]) : "r"(qa));
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(wb[s][0]), "=r"(wb[s][1]) : "r"(qb));
                }
            };
#pragma unroll
            for (int s = 0; s < 3; s++) load_xo(s);
#pragma unroll
            for (int s = 0; s < 2; s++) load_pw(s);
#pragma unroll
            for (int s = 0; s < 32; s++) {
                load_xo(s + 3);
                load_pw(s + 2);
                // rotate-shuffle: the bottom row of the lane above; row 0 of P (above lane 0) is zero
                unsigned up = __shfl_sync(kFull, last ? 0u : h[R - 1], src_lane);
                unsigned diag = dprev;
                dprev = up;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const unsigned left = h[r];
                    unsigned tt;
                    if (r < R - NP) {
                        tt = __dp4a(wa[s][r >> 2], 1u << (8 * (r & 3)), diag);                          // low half += s'_A
                        const unsigned half = (r & 1) ? 0x80000000u : 0x00008000u;                      // 32768 * (2 s'_B) = s'_B << 16
                        tt = (r & 2) ? __dp2a_hi(half, wb[s][r >> 2], tt) : __dp2a_lo(half, wb[s][r >> 2], tt);
                    } else {
                        constexpr unsigned j = 0;
                        const unsigned jj = (unsigned)(r & 3) + j;
                        const unsigned sel = jj | ((8u | jj) << 4) | ((4u + jj) << 8) | ((8u | (4u + jj)) << 12);
                        tt = diag + prmt_generic(wa[s][r >> 2], wb[s][r >> 2], sel);
                    }
                    const unsigned nv = __vimax3_u16x2(tt, up, left);
                    diag = left; up = nv; h[r] = nv;
                }
            }
            __syncwarp();
            check_put(cp, pf_a, pf_b);
            __syncwarp();
        }
        // lane 31's last row is row n of both matrices, frozen behind their last columns: un-shift H = P + (n+m)*gap
        if (last) {
            a.scores[pA] = tallA ? kBatchTooTall : (int)(h[R - 1] & 0xffffu) + gapsA;
            if (hasB) a.scores[pB] = tallB ? kBatchTooTall : (int)(h[R - 1] >> 16) + gapsB;
        }
    }
}

}  // namespace nwb
