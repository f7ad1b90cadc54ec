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

// TMA = true (A/B switch NWB200_BATCH_TMA=1, north_star: "sequences ... staged into shared memory via TMA"): the column letters
// reach shared memory as BULK COPIES of the TMA engine (cp.async.bulk global -> shared, 128 columns per copy, two landing buffers per
// warp, completion on an mbarrier each) instead of one prefetching LDG per lane and chunk; a lane then reads its letter with an
// LDS.  Needs a 16-byte aligned sequence start (pairs that are not take the LDG path).  Measured: profiles/r2w_* (DESIGN.md).
constexpr int kTmaBlock = 128;                                   // columns per bulk copy
__host__ __device__ constexpr size_t batch_tma_warp_bytes() { return 2 * kTmaBlock + 16; }      // two landing buffers + two mbarriers

template <int R, int WARPS, bool TMA = false>
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
    // TMA: landing buffers and mbarriers of this warp behind the per-warp regions
    const unsigned land_s = TMA ? (unsigned)__cvta_generic_to_shared(smem_raw + (((size_t)WARPS * SC::warp_smem_bytes(a.S) + 15) & ~(size_t)15) + (size_t)w * batch_tma_warp_bytes()) : 0u;
    const unsigned mbar_s = land_s + 2 * kTmaBlock;
    unsigned phase_bits = 0;                          // parity of the next completion of each landing buffer's mbarrier
    if constexpr (TMA) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s + 8u) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    // lane 0: bulk copy of block kb (columns 128 kb ..) of x into its landing buffer; bytes rounded up to 16 (the pool has slack)
    auto tma_issue = [&](const uint8_t* x, int m, int kb) {
        if constexpr (TMA) {
            const int c0 = kTmaBlock * kb;
            if (lane == 0 && c0 < m) {
                const unsigned bytes = (unsigned)((min(kTmaBlock, m - c0) + 15) & ~15);
                const unsigned dst = land_s + (unsigned)(kb & 1) * kTmaBlock, mb = mbar_s + 8u * (unsigned)(kb & 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(x + c0), "r"(bytes), "r"(mb) : "memory");
            }
        }
    };
    // all lanes: wait until block kb has landed
    auto tma_wait = [&](int m, int kb) {
        if constexpr (TMA) {
            if (kTmaBlock * kb < m) {
                const unsigned mb = mbar_s + 8u * (unsigned)(kb & 1), par = (phase_bits >> (kb & 1)) & 1u;
                unsigned done = 0, polls = 0;
                while (!done) {
                    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                                 : "=r"(done) : "r"(mb), "r"(par) : "memory");
                    if (++polls > (1u << 22)) { g_wait_timeout = 1; break; }
                }
                phase_bits ^= 1u << (kb & 1);
            }
        }
    };
    auto tma_letter = [&](int c) -> unsigned {      // letter of column c from its landing buffer
        unsigned v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(land_s + (unsigned)(c & (2 * kTmaBlock - 1))));
        return v;
    };
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
        const bool bulk = TMA && ((reinterpret_cast<unsigned long long>(x) & 15ull) == 0ull);      // (warp-uniform)
        if (bulk) { tma_issue(x, m, 0); tma_issue(x, m, 1); }      // land under the profile build
        build_profile<R, K>(sm, sp_tab, a.S, y, (long long)lane * R - pad, n, lane, nullptr);
        for (int c = -32 + lane; c < 0; c += 32) sm.put_letter(c, ZOFF);
        if (bulk) tma_wait(m, 0);
        for (int g = 0; g < PD; g++) {
            const int c = 32 * g + lane;
            unsigned v = c < m ? (bulk ? tma_letter(c) : (unsigned)__ldg(x + c)) : kPastEnd;
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
            unsigned pf_x;
            if (bulk) {
                const int g = lc + PD;                                             // group of 32 columns fetched now: block g / 4
                if ((g & 3) == 0) tma_wait(m, g >> 2);
                pf_x = (cp < m) ? tma_letter(cp) : kPastEnd;
                if ((g & 3) == 3) { __syncwarp(); tma_issue(x, m, (g >> 2) + 2); }     // every lane has read this block: its buffer takes the block after next
            } else {
                pf_x = (cp < m) ? (unsigned)__ldg(x + cp) : kPastEnd;             // checked and scaled when it lands
            }
            io.xs_lane = sm.xs + ((32 * lc - K * lane) & (XR - 1));
            sweep_chunk<R, K, 0, false>(st, lane, io, nullptr);
            __syncwarp();
            if (pf_x >= (unsigned)a.S) { if (pf_x != kPastEnd) *a.err = 1; pf_x = (unsigned)a.S; }
            sm.put_letter(cp, pf_x * SC::LSTRIDE);
            __syncwarp();
        }
        if (bulk) {
            // copies that were issued for blocks the sweep never read (the last groups lie past the end of x) must land before the
            // buffers are reused: wait for every block that was requested
            const int last_read = (nlc - 1 + PD) >> 2, last_issued = ((nlc - 1 + PD) >> 2) + 1 + (((nlc - 1 + PD) & 3) == 3 ? 1 : 0);
            for (int kb = last_read + 1; kb <= last_issued; kb++) tma_wait(m, kb);
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
