// nw_batch.cuh -- many independent short pairs (BASELINE config 3): one warp aligns one pair.
//
// The reference has no batch path: benchmark.cpp:406 calls align once per pair, and gpu9 launches
// trows+tcols-1 kernels of <= min(trows,tcols) blocks for each (3 launches of <= 2 blocks for 256 x 256).
// Here a pair of up to 32*R rows is ONE band: a warp sweeps it with the lanes one column apart (K = 1, the
// fill / drain of the systolic array costs 31 of m+31 steps), nothing touches HBM but the letters (read
// once) and the 4-byte score.  Warps take pairs from an atomic ticket, so ragged batches balance themselves.
#pragma once
#include "nw_engine.cuh"
#include "nw_sweep.cuh"

namespace nwb {


constexpr unsigned kPastEnd = 0x100u;      // "no letter here" while a byte letter is in flight: distinct from every byte, so that a real
                                           // letter equal to S (the internal zero-row code) is still reported as outside the alphabet

template <int R, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) nw_batch_kernel(const BatchArgs a)
{
    constexpr int K = 1;
    using SC = Sched<R, K>;
    constexpr int By = SC::By, PD = SC::PD, XR = SC::XR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned sp_tab[kSpWords];
    stage_sprime(sp_tab, a.sprime, a.S, smem_raw);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    WarpSmem<R, K> sm(smem_raw + (size_t)w * SC::warp_smem_bytes(a.S), a.S);
    const unsigned ZOFF = (unsigned)a.S * SC::LSTRIDE;
    ChunkIO io;
    io.prof_lane = sm.prof + lane * 4 * SC::WPL;
    io.rin_chunk = nullptr; io.rin_next = nullptr; io.rout_chunk = nullptr; io.rmid_chunk = nullptr; io.map_out = nullptr; io.org0 = 0;
    io.dirs_lane = nullptr; io.negg = 0;

    for (;;) {
        unsigned long long p = 0;
        if (lane == 0) p = a.first + atomicAdd(a.ticket, 1ull);
        p = __shfl_sync(kFull, p, 0);
        if (p >= a.npairs) break;
        const int n = (int)a.lenY[p], m = (int)a.lenX[p];
        if (n == 0 || m == 0) { if (lane == 0) a.scores[p] = (n + m) * a.gap; continue; }
        if (n > By) { if (lane == 0) a.scores[p] = kBatchTooTall; continue; }
        const uint8_t* y = a.letters + a.offY[p];
        const uint8_t* x = a.letters + a.offX[p];
        const int pad = By - n;                                    // rows are aligned to the bottom of the band
        __syncwarp();
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane * R - pad + r;
            if (i >= 0 && i < n && (unsigned)__ldg(y + i) >= (unsigned)a.S) *a.err = 1;
        }
        build_profile<R, K>(sm, sp_tab, a.S, y, (long long)lane * R - pad, n, lane, nullptr);
        for (int c = -32 + lane; c < 0; c += 32) sm.put_letter(c, ZOFF);
        for (int g = 0; g < PD; g++) {
            const int c = 32 * g + lane;
            unsigned v = c < m ? (unsigned)__ldg(x + c) : kPastEnd;
            if (v >= (unsigned)a.S) { if (v != kPastEnd) *a.err = 1; v = (unsigned)a.S; }      // a real letter must be < S
            sm.put_letter(c, v * SC::LSTRIDE);
        }
        __syncwarp();
        Lane<R, 0> st;
#pragma unroll
        for (int r = 0; r < R; r++) st.h[r] = 0;
        st.dprev = 0; st.up_next = 0; st.oprev = 0; st.oup_next = 0; st.o[0] = 0;
        const int nlc = SC::nlc(m);
        for (int lc = 0; lc < nlc; lc++) {
            const int cp = 32 * (lc + PD) + lane;
            unsigned pf_x = (cp < m) ? (unsigned)__ldg(x + cp) : kPastEnd;           // checked and scaled when it lands
            io.xs_lane = sm.xs + ((32 * lc - K * lane) & (XR - 1));
            sweep_chunk<R, K, 0, false>(st, lane, io, nullptr);
            __syncwarp();
            if (pf_x >= (unsigned)a.S) { if (pf_x != kPastEnd) *a.err = 1; pf_x = (unsigned)a.S; }
            sm.put_letter(cp, pf_x * SC::LSTRIDE);
            __syncwarp();
        }
        // lane 31's last row is row n of the matrix, frozen at column m: un-shift H = P + (n+m)*gap
        if (lane == 31) a.scores[p] = st.h[R - 1] + (n + m) * a.gap;
    }
}


// ---------------------------------------------------------------------------------------------- transcripts of a batch
// One warp per pair: the sweep of nw_batch_kernel with 2-bit move codes for EVERY cell kept in shared memory (MODE 2 of
// sweep_chunk: the reference's choice, nwtrace1_plain.cpp:29-100), then the walk from (n, m) back to (0, 0) over those codes --
// what nw_walk_kernel does for one band of a long pair, here for a whole short pair.  Emits the move list in backward path order
// (codes 0 '=', 1 'X', 2 'I', 3 'D'); run-length encoding + hash are O(path) byte work done by the host.

template <int R>
__global__ void __launch_bounds__(32) nw_batch_trace_kernel(const BatchTraceArgs a)
{
    constexpr int K = 1;
    using SC = Sched<R, K>;
    constexpr int By = SC::By, PD = SC::PD, XR = SC::XR, DB = R / 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned sp_tab[kSpWords];
    stage_sprime(sp_tab, a.sprime, a.S, smem_raw);
    const int lane = threadIdx.x & 31;
    WarpSmem<R, K> sm(smem_raw, a.S);
    unsigned char* dirs = smem_raw + SC::warp_smem_bytes(a.S);
    const unsigned ZOFF = (unsigned)a.S * SC::LSTRIDE;
    ChunkIO io;
    io.prof_lane = sm.prof + lane * 4 * SC::WPL;
    io.rin_chunk = nullptr; io.rin_next = nullptr; io.rout_chunk = nullptr; io.rmid_chunk = nullptr; io.map_out = nullptr; io.org0 = 0;
    io.negg = a.negg;

    for (unsigned long long p = a.first + blockIdx.x; p < a.npairs; p += gridDim.x) {
        const unsigned long long q = p - a.first;
        const int n = (int)a.lenY[p], m = (int)a.lenX[p];
        const int nlc = SC::nlc(m);
        if (n == 0 || m == 0 || n > By || nlc > a.chunks_cap) { if (lane == 0) a.cnt[q] = -1; continue; }
        const uint8_t* y = a.letters + a.offY[p];
        const uint8_t* x = a.letters + a.offX[p];
        const int pad = By - n;                                    // rows are aligned to the bottom of the band
        __syncwarp();
        unsigned yl[R], yoff[R];
        build_profile<R, K>(sm, sp_tab, a.S, y, (long long)lane * R - pad, n, lane, yl);
#pragma unroll
        for (int r = 0; r < R; r++) yoff[r] = yl[r] * SC::LSTRIDE;
        for (int c = -32 + lane; c < 0; c += 32) sm.put_letter(c, ZOFF);
        for (int g = 0; g < PD; g++) {
            const int c = 32 * g + lane;
            unsigned v = c < m ? (unsigned)__ldg(x + c) : (unsigned)a.S;
            if (v > (unsigned)a.S) v = (unsigned)a.S;
            sm.put_letter(c, v * SC::LSTRIDE);
        }
        __syncwarp();
        Lane<R, 2> st;
#pragma unroll
        for (int r = 0; r < R; r++) st.h[r] = 0;
        st.dprev = 0; st.up_next = 0; st.oprev = 0; st.oup_next = 0; st.o[0] = 0;
        for (int lc = 0; lc < nlc; lc++) {
            const int cp = 32 * (lc + PD) + lane;
            unsigned pf_x = (cp < m) ? (unsigned)__ldg(x + cp) : (unsigned)a.S;
            io.xs_lane = sm.xs + ((32 * lc - K * lane) & (XR - 1));
            io.dirs_lane = dirs + ((size_t)(32 * lc) * 32 + lane) * DB;
            sweep_chunk<R, K, 2, false>(st, lane, io, yoff);
            __syncwarp();
            if (pf_x > (unsigned)a.S) pf_x = (unsigned)a.S;
            sm.put_letter(cp, pf_x * SC::LSTRIDE);
            __syncwarp();
        }
        // ---- walk (uniform across the warp; lane 0 stores): cell (row, j) of the band was computed by lane row / R at step j-1 + K*lane
        unsigned char* out = a.moves + a.moff[q];
        int j = m, row = By - 1, cnt = 0;
        while (j > 0 && row >= pad) {
            const int ln = row / R, r = row % R;
            const int step = (j - 1) + K * ln;
            unsigned codes;
            if constexpr (DB == 1) codes = dirs[step * 32 + ln];
            else if constexpr (DB == 2) codes = reinterpret_cast<const unsigned short*>(dirs)[step * 32 + ln];
            else codes = reinterpret_cast<const unsigned*>(dirs)[step * 32 + ln];
            const unsigned code = (codes >> (2 * r)) & 3u;
            if (lane == 0) out[cnt] = (unsigned char)code;
            cnt++;
            if (code < 2u) { row--; j--; }
            else if (code == 2u) row--;
            else j--;
        }
        if (j == 0) {                              // column 0: straight up (nwtrace1_plain.cpp:57-63 with j == 0)
            const int k = row - pad + 1;
            for (int t = lane; t < k; t += 32) out[cnt + t] = 2;
            cnt += k > 0 ? k : 0;
        } else {                                   // matrix row 0: the rest of the path runs left along it
            for (int t = lane; t < j; t += 32) out[cnt + t] = 3;
            cnt += j;
        }
        if (lane == 0) a.cnt[q] = cnt;
    }
}

}  // namespace nwb
