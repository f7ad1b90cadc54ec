// nw_fill.cuh -- score-matrix fill for ONE pair: the replacement of Nw_Gpu9_KernelA/KernelB
// (reference nwalign_gpu9_mlsp_diagdiagdiag.cu:15-63,69-360) and of the per-diagonal CUDA graph
// that drives them (:555-684).
//
// Design (B200-first, not a translation):
//   * The matrix is cut into horizontal BANDS of By = W*32*R rows.  One CTA (W warps) owns one
//     band and sweeps it left to right over ALL columns in a single pass; CTAs are persistent
//     and take bands from an atomic ticket, so band b-1 is always running (or done) when band b
//     starts: no launch-per-diagonal, no global barrier, no cooperative launch.
//   * Inside a warp every ROW is one stage of a register-resident systolic array: lane l keeps
//     R consecutive rows, row r of lane l works on column  s - KS*l - r  at step s (KS = R+K-1).
//     Each row lags the row above by one column, so the R cell updates a lane issues in one step
//     are mutually independent (no serial max chain inside a step) and the whole warp advances
//     one anti-diagonal of its band per step.  The value crossing a lane boundary moves with ONE
//     SHFL.UP per step (K = 2: issued a step early, off the critical path); the value crossing a
//     warp boundary goes through a small shared-memory ring; the value crossing a CTA (band)
//     boundary is the band's bottom row, which IS the tile header row the traceback needs anyway:
//     it is streamed through L2 as 64-bit (epoch-tag | value) elements, so data and ready-flag
//     arrive in one atomic store and the consumer never fences.
//   * All arithmetic is done in shifted coordinates P = H - (i+j)*gap, so a
//     cell is one IDP.4A (adds the byte score s' = max(subst-2*gap,0) picked out of a packed
//     per-lane profile word; fma pipe) and one VIMNMX3 (alu pipe).  Row 0 / column 0 are all
//     zeros in P, which is why there is no header-init kernel (reference KernelA).
//   * Only header rows (one per band), header columns (one per Bx columns, optional) and the
//     last column leave the SM.
#pragma once
#include "nw_common.cuh"

namespace nwb {

__host__ __device__ constexpr int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// Compile-time schedule of one kernel shape.
template <int R, int W, int K>
struct Sched {
    static constexpr int WPL = R / 4;                      // profile words per lane
    static constexpr int By = W * 32 * R;                  // rows per band (= tile height)
    static constexpr int KS = R + K - 1;                   // column skew between consecutive lanes
    static constexpr int OFF = 31 * KS + R - 1;            // column lag of the warp's last row behind its first
    static constexpr int G = (31 * KS + 31) / 32;          // sequence groups a chunk reaches back
    static constexpr int DW = (31 + OFF) / 32;             // chunks until a 32-column group of the bottom row is complete
    static constexpr int D = DW + 1;                       // chunks of lag between consecutive warps of a CTA
    static constexpr int PHI = (32 * 64 - OFF) % 32;       // ring positions of a chunk that may wrap
    static constexpr int VR = 128;                         // ints per value ring (writer/reader span <= 94 columns)
    static constexpr int XR = next_pow2(32 * (D * (W - 1) + G + 5));   // bytes of sequence window ring
    static constexpr int XM = 32 * G + 32;                 // mirror bytes behind the sequence ring
    __host__ __device__ static constexpr size_t smem_bytes(int S)
    {
        return ((size_t)W * (S + 1) * 32 * WPL + (size_t)(W + 1) * VR) * 4 + XR + XM + (size_t)S * S + 16;
    }
    __host__ __device__ static constexpr int nlc(int m) { return (m + OFF + 31) / 32; }   // local chunks per warp: steps 0 .. m-1+OFF
};

struct FillArgs {
    const uint8_t* y;        // lenY letters
    const uint8_t* x;        // lenX letters
    int n, m;                // lenY, lenX
    const uint8_t* sprime;   // S*S bytes: s'[y*S+x]
    int S;                   // alphabet size (<= kMaxLetters)
    unsigned long long* HR;  // header rows, one 64-bit element (tag << 32 | P) per cell:
                             // HR[b*ldr + c] = P[b*By][c+1], b = 1..trows-1; the tag is the run epoch, so
                             // data and "ready" flag travel in one atomic store (no fences, no separate flag)
    long long ldr;           // >= 32*nlc
    int* HC;                 // P-space header cols: HC[q*ldc + i0] = P[i0+1][(q+1)*Bx], q = 0..tcols-2 (may be null)
    long long ldc;
    int* lastcol;            // P[i0+1][m] for every (padded) row i0
    unsigned tag;            // epoch tag of this run (never 0)
    int* ticket;             // band ticket counter
    int Bx;                  // header column spacing (multiple of 32)
    int trows;               // number of bands
    int keep_hdr;            // write HC
    unsigned backoff_ns;     // sleep after an in-loop miss of the band above (drops this band one chunk behind)
    unsigned long long* dbg; // optional [trows][4] globaltimer stamps: band start, first chunk, last chunk, end
};

template <int R, int WPL>
struct Lane {
    int h[R];            // h[r] = P[row r][col_r - 1]: the "left" neighbour of the cell row r does next
    int g[R];            // g[r] = h[r] one step earlier = diagonal neighbour for row r+1
    int upprev;          // diagonal neighbour for row 0 (the previous step's `up`)
    int up_next;         // K == 2 only: the shuffled upper neighbour for the next step
    unsigned pwh[R > 1 ? R - 1 : 1][WPL];   // profile words of the previous R-1 steps (row r uses the word of step s-r)
};

// smallest header/last target column >= c
__device__ __forceinline__ int next_target(int c, int Bx, int m, int keep_hdr)
{
    int last = m - 1;
    if (!keep_hdr) return last;
    int t = ((c + Bx) / Bx) * Bx - 1;
    return t < last ? t : last;
}

struct Target {
    int col;       // next header column (c = k*Bx-1 < m-1) at which this lane stores its R values, or INT_MAX
    int* dst;      // where (already offset to this lane's first row)
};

// One 32-step chunk of one warp.  cb = column of this lane's row 0 at s = 0.
template <int R, int W, int K, bool CHECKED>
__device__ __forceinline__ void run_chunk(Lane<R, R / 4>& st, const int lane, const int cb,
                                          const uint8_t* __restrict__ xs_lane,    // letters of row 0, index s
                                          const unsigned* __restrict__ prof_lane, // &prof[w][0][lane*WPL]
                                          const int* __restrict__ rin_chunk,      // lane 0 input, index s
                                          int* __restrict__ rout, const int rout_base,
                                          Target& tgt, const FillArgs& a, const long long row0)
{
    using SC = Sched<R, W, K>;
    constexpr int WPL = SC::WPL;
    constexpr int VR = SC::VR;
    constexpr int H = R - 1;     // history depth
    // software pipeline: letters 3 steps ahead, profile words 2 steps ahead (LDS latency ~34 clk each)
    unsigned xl[32 + 3];
    unsigned pw[H + 32 + 2][WPL];     // pw[H + s] is the word of step s; pw[0..H-1] come from the previous chunk
    int rv[32];                   // lane-0 input (row above the band), broadcast load, 2-3 steps ahead
    const int src_lane = (lane + 31) & 31;
#pragma unroll
    for (int i = 0; i < H; i++)
#pragma unroll
        for (int q = 0; q < WPL; q++) pw[i][q] = st.pwh[i][q];
#pragma unroll
    for (int s = 0; s < 3; s++) xl[s] = xs_lane[s];
#pragma unroll
    for (int s = 0; s < 3; s++) rv[s] = rin_chunk[s];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const unsigned* pp = prof_lane + xl[s] * (32 * WPL);
        if constexpr (WPL == 1) pw[H + s][0] = pp[0];
        else { uint2 v = *reinterpret_cast<const uint2*>(pp); pw[H + s][0] = v.x; pw[H + s][1] = v.y; }
    }
#pragma unroll
    for (int s = 0; s < 32; s++) {
        if (s + 3 < 32) xl[s + 3] = xs_lane[s + 3];
        if (s + 3 < 32) rv[s + 3] = rin_chunk[s + 3];
        if (s + 2 < 32) {
            const unsigned* pp = prof_lane + xl[s + 2] * (32 * WPL);
            if constexpr (WPL == 1) pw[H + s + 2][0] = pp[0];
            else { uint2 v = *reinterpret_cast<const uint2*>(pp); pw[H + s + 2][0] = v.x; pw[H + s + 2][1] = v.y; }
        }
        // Rotate-shuffle: lanes 0..30 hand their bottom row to the lane below; lane 31 (whose bottom row
        // goes to the ring instead) hands lane 0 its next input from the row above the band, so the
        // shuffle result feeds VIMNMX3 directly (no select behind the SHFL latency).
        int up;
        if constexpr (K == 1) {
            up = __shfl_sync(kFull, lane == 31 ? rv[s] : st.h[R - 1], src_lane);
        } else {
            up = st.up_next;
            if (s == 0 && lane == 0) up = rv[0];      // the value for a chunk's first step cannot be pre-shuffled (not yet produced)
            st.up_next = __shfl_sync(kFull, lane == 31 ? rv[s < 31 ? s + 1 : 31] : st.h[R - 1], src_lane);   // consumed at step s+1
        }
        // R independent cell updates: row r is at column cb + s - r and uses the profile word of step s - r
        int nh[R];
        nh[0] = max3(add_byte(pw[H + s][0], 1u, st.upprev), up, st.h[0]);
#pragma unroll
        for (int r = 1; r < R; r++)
            nh[r] = max3(add_byte(pw[H + s - r][r >> 2], 1u << (8 * (r & 3)), st.g[r - 1]), st.h[r - 1], st.h[r]);
        st.upprev = up;
#pragma unroll
        for (int r = 0; r < R; r++) { st.g[r] = st.h[r]; st.h[r] = nh[r]; }
        if (lane == 31) rout[(s < 32 - SC::PHI) ? rout_base + s : ((rout_base + s) & (VR - 1))] = st.h[R - 1];
        if (CHECKED) {
            const int c0 = cb + s;
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (c0 - r == tgt.col) st_cs(tgt.dst + r, st.h[r]);              // header column
                if (c0 - r == a.m - 1) st_cs(a.lastcol + row0 + r, st.h[r]);      // last column: the score lives here
            }
            if (c0 - (R - 1) == tgt.col) {
                tgt.col += a.Bx; tgt.dst += a.ldc;
                if (tgt.col >= a.m - 1) tgt.col = 0x7fffffff;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < H; i++)
#pragma unroll
        for (int q = 0; q < WPL; q++) st.pwh[i][q] = pw[32 + i][q];
}

template <int R, int W, int K>
__global__ void __launch_bounds__(W * 32) nw_fill_kernel(const FillArgs a)
{
    using SC = Sched<R, W, K>;
    constexpr int WPL = SC::WPL, By = SC::By, KS = SC::KS, OFF = SC::OFF, G = SC::G, DW = SC::DW, D = SC::D;
    constexpr int VR = SC::VR, XR = SC::XR, XM = SC::XM;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: prof [W][(S+1)][32][WPL] words | rings [W][VR] | rin [VR] | xs [XR+XM] | sp [S*S]
    const int S1 = a.S + 1;
    unsigned* prof = reinterpret_cast<unsigned*>(smem_raw);
    int* rings = reinterpret_cast<int*>(prof + (size_t)W * S1 * 32 * WPL);
    int* rin = rings + W * VR;
    uint8_t* xs = reinterpret_cast<uint8_t*>(rin + VR);
    uint8_t* sp = xs + XR + XM;
    __shared__ int s_band;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int m = a.m, n = a.n;
    const int nlc = SC::nlc(m);
    const int ngc = nlc + D * (W - 1);             // global chunks per band
    const uint8_t ZL = (uint8_t)a.S;               // the all-zero profile row

    for (int i = tid; i < a.S * a.S; i += W * 32) sp[i] = a.sprime[i];

    for (;;) {
        __syncthreads();
        if (tid == 0) s_band = atomicAdd(a.ticket, 1);
        __syncthreads();
        const int b = s_band;
        if (b >= a.trows) break;

        if (a.dbg && tid == 0) a.dbg[4 * b + 0] = globaltimer_ns();
        const long long row0 = (long long)b * By + (long long)w * 32 * R + (long long)lane * R;   // 0-based first row
        // ---- per-lane profile: prof[w][xl][lane] = bytes s'(y[row0+r], xl), r = 0..R-1; row S is all zero
        {
            unsigned yl[R];
#pragma unroll
            for (int r = 0; r < R; r++) yl[r] = (row0 + r < n) ? a.y[row0 + r] : 0u;
            unsigned* pl = prof + ((size_t)w * S1 * 32 + lane) * WPL;
            for (int xl = 0; xl < a.S; xl++) {
#pragma unroll
                for (int q = 0; q < WPL; q++) {
                    unsigned word = 0;
#pragma unroll
                    for (int r = 0; r < 4; r++) word |= (unsigned)sp[yl[q * 4 + r] * a.S + xl] << (8 * r);
                    pl[(size_t)xl * 32 * WPL + q] = word;
                }
            }
#pragma unroll
            for (int q = 0; q < WPL; q++) pl[(size_t)a.S * 32 * WPL + q] = 0u;
        }
        // ---- sequence window: groups -G..-1 (zero letters), 0, 1, 2; value rings cleared
        for (int i = tid; i < (G + 3) * 32; i += W * 32) {
            int c = i - 32 * G;
            uint8_t v = (c >= 0 && c < m) ? a.x[c] : ZL;
            int p = c & (XR - 1);
            xs[p] = v;
            if (p < XM) xs[p + XR] = v;
        }
        for (int i = tid; i < (W + 1) * VR; i += W * 32) rings[i] = 0;
        __syncthreads();

        Lane<R, WPL> st;
#pragma unroll
        for (int r = 0; r < R; r++) { st.h[r] = 0; st.g[r] = 0; }
        st.upprev = 0;
        st.up_next = 0;
#pragma unroll
        for (int i = 0; i < (R > 1 ? R - 1 : 1); i++)
#pragma unroll
            for (int q = 0; q < WPL; q++) st.pwh[i][q] = 0u;
        int hdr_next = a.Bx - 1;
        Target tgt;
        tgt.col = (a.keep_hdr && a.Bx - 1 < m - 1) ? a.Bx - 1 : 0x7fffffff;
        tgt.dst = a.HC + row0;
        const unsigned* prof_lane = prof + ((size_t)w * S1 * 32 + lane) * WPL;
        int* rout = rings + w * VR;
        const int* rprev = (w == 0) ? rin : rings + (w - 1) * VR;
        const bool consumer = (b > 0);                       // band 0 has P = 0 above it
        const bool producer = (b + 1 < a.trows);             // last band feeds nobody
        const unsigned long long* hr_in = a.HR + (long long)b * a.ldr;
        unsigned long long* hr_out = a.HR + (long long)(b + 1) * a.ldr;
        // warps entirely below the matrix only keep the barriers company
        const bool wactive = ((long long)b * By + (long long)w * 32 * R) < n;

        // prefetch registers (loaded in one chunk, stored to shared memory at its end)
        unsigned long long pf_hr = 0; int pf_hr_grp = -1;
        unsigned dbg_spins = 0;      // lane-private: (chunks with a miss << 16) + re-polls
        uint8_t pf_x = ZL; int pf_x_grp = -1;

        // warp 0 prologue: header-row groups 0 and 1 of the band above
        if (w == 0 && consumer) {
            for (int g = 0; g < 2 && g < nlc; g++)
                rin[(32 * g + lane) & (VR - 1)] = wait_tagged(hr_in + 32 * g + lane, ld_relaxed64(hr_in + 32 * g + lane), a.tag);
        }
        __syncthreads();

        if (a.dbg && tid == 0) a.dbg[4 * b + 1] = globaltimer_ns();
        for (int gc = 0; gc < ngc; gc++) {
            const int lc = gc - D * w;
            // ---- issue prefetches for later chunks (warp 0: header row of band above; warp W-1: x letters)
            if (w == 0 && consumer) {
                const int g = gc + 2;
                if (g < nlc) {
                    pf_hr = ld_relaxed64(hr_in + 32 * g + lane);      // checked (and re-polled if early) at the end of the chunk
                    pf_hr_grp = g;
                }
            }
            if (w == W - 1) {
                const int g = gc + 3;
                const int c = 32 * g + lane;
                pf_x = (c < m) ? __ldg(a.x + c) : ZL;
                pf_x_grp = g;
            }
            // ---- the chunk itself
            if (lc >= 0 && lc < nlc && wactive) {
                const int cb = 32 * lc - KS * lane;
                const uint8_t* xs_lane = xs + ((32 * (lc - G)) & (XR - 1)) + 32 * G - KS * lane;
                const int* rin_chunk = rprev + ((32 * lc) & (VR - 1));
                const int rout_base = (32 * lc - OFF) & (VR - 1);
                // does any row of this warp meet a target column in this chunk?
                const int lo = 32 * lc - OFF, hi = 32 * lc + 31;
                while (hdr_next < lo) hdr_next += a.Bx;             // next header column >= lo (warp-uniform)
                const bool hit_hdr = a.keep_hdr && hdr_next <= hi && hdr_next < m - 1;
                const bool hit_last = (m - 1 >= lo) && (m - 1 <= hi);
                if (hit_hdr || hit_last)
                    run_chunk<R, W, K, true>(st, lane, cb, xs_lane, prof_lane, rin_chunk, rout, rout_base, tgt, a, row0);
                else
                    run_chunk<R, W, K, false>(st, lane, cb, xs_lane, prof_lane, rin_chunk, rout, rout_base, tgt, a, row0);
                // ---- bottom warp streams the band's bottom row (= header row of band b+1)
                if (w == W - 1 && producer && lc >= DW) {
                    __syncwarp();
                    const int g = lc - DW;
                    st_relaxed64(hr_out + 32 * g + lane, pack_tagged(rout[(32 * g + lane) & (VR - 1)], a.tag));
                }
            }
            // ---- land the prefetches in shared memory for the chunks after the barrier
            if (w == 0 && pf_hr_grp >= 0) {
                rin[(32 * pf_hr_grp + lane) & (VR - 1)] = wait_tagged_backoff(hr_in + 32 * pf_hr_grp + lane, pf_hr, a.tag, a.backoff_ns, dbg_spins);
                pf_hr_grp = -1;
            }
            if (w == W - 1 && pf_x_grp >= 0) {
                int p = (32 * pf_x_grp + lane) & (XR - 1);
                xs[p] = pf_x;
                if (p < XM) xs[p + XR] = pf_x;
                pf_x_grp = -1;
            }
            __syncthreads();
            if (a.dbg && tid == 0 && b < 4 && gc < 600) a.dbg[4 * a.trows + b * 600 + gc] = globaltimer_ns();
        }
        if (a.dbg && tid == 0) { a.dbg[4 * b + 2] = globaltimer_ns(); a.dbg[4 * b + 3] = dbg_spins; }
        // ---- publish the last header-row groups so the band below can finish
        if (w == W - 1 && producer) {
            for (int g = (nlc - DW > 0 ? nlc - DW : 0); g < nlc; g++)
                st_relaxed64(hr_out + 32 * g + lane, pack_tagged(rout[(32 * g + lane) & (VR - 1)], a.tag));
        }
    }
}

}  // namespace nwb
