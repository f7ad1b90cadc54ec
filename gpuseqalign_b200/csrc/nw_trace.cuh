// nw_trace.cuh -- traceback on the GPU from the sparse representation the fill leaves in HBM (band header
// rows + register snapshots): the reference's nwtrace2_sparse scheme (NwTrace2_Sparse,
// nwtrace2_sparse.cpp:102-257: recompute a tile from its headers, walk it, hop to the next tile), re-designed
// so that nothing but a short pointer chase is sequential:
//
//   pass A  nw_map_kernel    every band, in parallel, at full issue rate: re-sweep the band from its header row
//                            with an ORIGIN label per cell (the column at which the traceback path through that
//                            cell leaves the band through its top row).  The bottom row's labels are the band's
//                            entry -> exit map.
//   pass B  nw_hop_kernel    one thread: entry[nb-1] = m, entry[b-1] = map[b][entry[b]] -- nb dependent loads.
//                            Now every band knows where the path enters it.
//   pass C  nw_walk_kernel   every band, in parallel: resume the sweep from the nearest snapshot left of the entry,
//                            write 2-bit move codes for the window into shared memory, walk them (the only
//                            cell-by-cell sequential part, <= By + window steps per band), emit one byte per move.
//   pass D  nw_pack_kernel   concatenate the per-band move lists (backward path order) into one dense array.
//
// The move choice is the reference's (nwtrace1_plain.cpp:29-100): neighbour scores, strict '<', diag > up > left.
// Run-length encoding + djb2 hash of the transcript (nwtrace1_plain.cpp:81-128) is O(path) byte work done by the
// host on the dense move list.
#pragma once
#include "nw_sweep.cuh"

namespace nwb {

struct TraceArgs {
    const uint8_t* y;
    const uint8_t* x;
    int n, m;
    const uint8_t* sprime;
    int S;
    int negg;                        // -gap
    const unsigned long long* HR;    // header rows of the fill (low 32 bits = P)
    long long ldr;
    const int* snap;                 // snapshots of the fill
    int nsnap, snap_chunks;
    int* map;                        // map[b*ldr + kPadL + c]: exit column (matrix j) on the top row for entry (bottom row, column c)
    int nb, pad;
    int* entry;                      // [nb]   entry column (matrix j) of the path on the bottom row of band b
    int* exitj;                      // [nb]   column at which the walker left band b (consistency check)
    long long* off;                  // [nb+1] start of band b's region in ops (backward path order: band nb-1 first)
    int* cnt;                        // [nb]   moves emitted by band b
    unsigned char* ops;              // per-band move lists, codes 0 '=', 1 'X', 2 'I', 3 'D'
    unsigned char* dense;            // packed backward move list
    long long* total;                // [4]: total moves, consistency flag, the fill's score element (tag | P), the engine's wait-timeout flag
                                     // -- header of the dense list, so that ONE device-to-host copy brings everything a pair needs
    const unsigned long long* score_elem;
    int map_half;                    // 1: map rows 2b (upper half of band b) and 2b+1 (lower half); 0: one map row per band
    // ---- segmented maps (nw_map_kernel): a band's map is computed in nseg independent column segments, every segment resuming
    // the sweep from a snapshot of the fill.  A path that leaves a segment through its LEFT cut carries the negative label
    // -(1 + lane*(R+2) + slot) of the cut cell (slot r < R: the lane's row r, slot R: the cell above its first row one column
    // to the left, slot R+1: the same cell of the cut column); cut[((b*nseg + j)*32 + lane)*(R+2) + slot] holds the label that
    // cell got in segment j, its own left neighbour: the chase follows these until it meets a top-row column.
    int* cut;
    int nseg;
    // ---- corridor maps (long pairs): pass A computes only the segments of a band within corr_d columns of the straight line from
    // (n, m) to the origin -- corr_w segments per band instead of nseg -- and pass B checks every lookup against that range.  A path
    // that leaves the corridor sets *miss; the full pass A and pass B, enqueued right behind (phase 1), then run -- they return at
    // once otherwise.  The walkers check the result either way (exitj against entry).
    int corr_w;                      // segments per band of this pass's corridor; 0: every segment
    int corr_d;                      // half width of the corridor in columns
    int phase;                       // 0: first pass; > 0: a wider corridor / the full pass, which runs only when the pass before it missed
    const int2* corr_tab;            // corridor passes: (first, last) segment of every band, tabulated by the host (the formula has three 64-bit
                                     // divisions -- a third of the one-warp pointer chase when it ran per band on the device)
    int* miss;                       // [0] the pass before this one missed, [1] passes that missed so far (nwb200_trace_info)
    // A cascade: a narrow corridor, a wide one, everything -- each pass enqueued behind the one before it and returning at once unless
    // miss[0] is set (the traceback of two unrelated or of two similar 200 000-letter sequences stays within 512 columns of the line).
};

// Segments [lo, hi] of band b that the corridor pass computes (every segment without a corridor).  Column c of a band's bottom row
// comes out of chunk (c + lag) / 32, i.e. of segment (c + lag) / (32 snap_chunks).
__host__ __device__ __forceinline__ void corridor_range(const TraceArgs& a, int b, int By, int lag, int& lo, int& hi)
{
    if (a.corr_w <= 0) { lo = 0; hi = a.nseg - 1; return; }
#ifdef __CUDA_ARCH__
    if (a.corr_tab != nullptr) { const int2 t = __ldg(a.corr_tab + b); lo = t.x; hi = t.y; return; }
#endif
    long long rb = (long long)(b + 1) * By - a.pad, rt = (long long)b * By - a.pad;      // matrix rows of the band's last row and of the row above its first
    if (rb > a.n) rb = a.n;
    if (rt < 0) rt = 0;
    const long long per = 32LL * a.snap_chunks;
    long long xlo = (long long)a.m * rt / a.n - a.corr_d, xhi = (long long)a.m * rb / a.n + a.corr_d;
    if (xlo < 0) xlo = 0;
    long long l = (xlo + lag) / per, h = (xhi + lag) / per;
    if (h > a.nseg - 1) h = a.nseg - 1;
    if (l > h) l = h;
    lo = (int)l; hi = (int)h;
}

// Loads a plain (already complete) header-row group into the top-row ring.
template <int R, int K>
__device__ __forceinline__ void load_top_group(const WarpSmem<R, K>& sm, const unsigned long long* hr_in, int g, int m, int lane)
{
    const int c = 32 * g + lane;
    sm.rin[c & (Sched<R, K>::VR - 1)] = (hr_in != nullptr && c < m) ? (int)(unsigned)__ldg(hr_in + c) : 0;
}
template <int R, int K>
__device__ __forceinline__ void load_letter_group(WarpSmem<R, K>& sm, const uint8_t* __restrict__ x, int g, int m, int lane, unsigned zoff)
{
    const int c = 32 * g + lane;
    sm.put_letter(c, (c >= 0 && c < m) ? (unsigned)__ldg(x + c) * Sched<R, K>::LSTRIDE : zoff);
}

// ---------------------------------------------------------------------------------------------- pass A
// Units = (band, column segment).  Bands are independent once the fill is done, and so are the segments of a band: the fill's
// snapshots hold the register state at every segment boundary.  One warp per band (782 warps for a 200k pair) left the machine
// at one warp per SM sub-partition, i.e. at a third of its issue rate (28 ms for the maps against 12.7 ms for the fill); the
// segments give as many units as the throughput regime needs.
template <int R, int K, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) nw_map_kernel(const TraceArgs a)
{
    using SC = Sched<R, K>;
    constexpr int By = SC::By, LAG = SC::LAG, PD = SC::PD, VR = SC::VR, XR = SC::XR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned sp_tab[kSpWords];
    stage_sprime(sp_tab, a.sprime, a.S, smem_raw);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    WarpSmem<R, K> sm(smem_raw + (size_t)w * SC::warp_smem_bytes(a.S), a.S);
    const int m = a.m, nlc = SC::nlc(m);
    const unsigned ZOFF = (unsigned)a.S * SC::LSTRIDE;
    const long long nwarps = (long long)gridDim.x * WARPS;
    const int nseg = a.nseg;
    if (a.phase > 0 && __ldcg(a.miss) == 0) return;               // the pass before found the whole path
    const bool corr = a.corr_w > 0;
    const int per_band = corr ? a.corr_w : nseg;
    const long long nunits = (long long)(a.nb - 1) * per_band;   // band 0 needs no map: the walker of band 0 ends the path itself
    for (long long u = (long long)blockIdx.x * WARPS + w; u < nunits; u += nwarps) {
        const int b = 1 + (int)(u / per_band);
        int seg = (int)(u % per_band);
        if (corr) {
            int lo, hi;
            corridor_range(a, b, By, LAG, lo, hi);
            seg += lo;
            if (seg > hi) continue;
        }
        const int lc0 = seg * a.snap_chunks;
        int lc1 = lc0 + a.snap_chunks;
        if (seg == nseg - 1 || lc1 > nlc) lc1 = nlc;
        const long long prow0 = (long long)b * By + (long long)lane * R;
        __syncwarp();
        build_profile<R, K>(sm, sp_tab, a.S, a.y, prow0 - a.pad, a.n, lane, nullptr);
        const unsigned long long* hr_in = a.HR + (long long)b * a.ldr + kPadL;
        for (int g = lc0 - 2 * K; g < lc0 + PD; g++) load_letter_group<R, K>(sm, a.x, g, m, lane, ZOFF);
        for (int i = lane; i < VR; i += 32) sm.rin[i] = 0;
        __syncwarp();
        for (int g = lc0; g < lc0 + PD; g++) load_top_group<R, K>(sm, hr_in, g, m, lane);
        __syncwarp();
        Lane<R, 1> st;
        if (lc0 > 0) {                                // resume from the fill's snapshot; the cells of the cut get fresh (negative) labels
            const int* sp = a.snap + (((long long)b * a.nsnap + (seg - 1)) * 32 + lane) * SC::SNAP_INTS;
            const int id0 = -(1 + lane * (R + 2));
#pragma unroll
            for (int r = 0; r < R; r++) { st.h[r] = sp[r]; st.o[r] = id0 - r; }
            st.dprev = sp[R];
            // lane 0's next input from above is the top row at the segment's first column: taken from the header row, not from the
            // snapshot (the snapshot at the END of a column block of the cross-GPU fill saw a zero there: that column belongs to the
            // next block, whose owner holds its header-row element)
            st.up_next = (lane == 0) ? sm.rin[(32 * lc0) & (VR - 1)] : sp[R + 1];
            // lane 0's upper neighbours are cells of the band's top row: their label is their own column
            st.oprev = (lane == 0) ? 32 * lc0 : id0 - R;
            st.oup_next = (lane == 0) ? 32 * lc0 + 1 : id0 - (R + 1);
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) { st.h[r] = 0; st.o[r] = 0; }
            st.dprev = 0; st.oprev = 0;
            st.up_next = (lane == 0) ? sm.rin[0] : 0;
            st.oup_next = (lane == 0) ? 1 : 0;
        }
        ChunkIO io;
        io.prof_lane = sm.prof + lane * 4 * SC::WPL;
        io.rout_chunk = nullptr; io.rmid_chunk = nullptr; io.dirs_lane = nullptr; io.negg = a.negg;
        int* map_row = a.map + (long long)b * a.ldr + kPadL;
        for (int lc = lc0; lc < lc1; lc++) {
            const int cp = 32 * (lc + PD) + lane;
            const int pf_top = (cp < m) ? (int)(unsigned)__ldg(hr_in + cp) : 0;
            const unsigned pf_x = (cp < m) ? (unsigned)__ldg(a.x + cp) : (unsigned)a.S;      // scaled when it lands
            io.xs_lane = sm.xs + ((32 * lc - K * lane) & (XR - 1));
            io.rin_chunk = sm.rin + ((32 * lc) & (VR - 1));
            io.rin_next = sm.rin + ((32 * lc + 32) & (VR - 1));
            io.map_out = map_row + (32 * lc - LAG);
            io.org0 = 32 * lc + 1;
            sweep_chunk<R, K, 1>(st, lane, io, nullptr);
            __syncwarp();
            sm.rin[cp & (VR - 1)] = pf_top;
            sm.put_letter(cp, pf_x * SC::LSTRIDE);
            __syncwarp();
        }
        if (seg + 1 < nseg) {                         // the labels of this segment's right cut, for the segment to the right
            int* cp = a.cut + (((long long)b * nseg + seg) * 32 + lane) * (R + 2);
#pragma unroll
            for (int r = 0; r < R; r++) cp[r] = st.o[r];
            cp[R] = st.oprev;
            cp[R + 1] = st.oup_next;
        }
    }
}

// ---------------------------------------------------------------------------------------------- pass B
// One warp.  Also lays out the per-band regions of the move buffer: band b can emit at most
// By + (entry[b] - entry[b-1]) moves (+ slack), band 0 at most By + entry[0].
// The chase itself is a chain of dependent loads (one or two per band), each an L2 round trip of ~0.4 us: 100 us for the 256
// half-band maps of a 16k pair.  The path runs roughly along the line to the origin, so the warp copies, kHopAhead lookups ahead,
// the 1024 map entries around the column that line predicts into shared memory (cp.async, one group per lookup): the dependent
// load then is a shared-memory read.  A wrong guess costs the L2 round trip, never the result.
constexpr int kHopAhead = 10, kHopWin = 256;
__global__ void __launch_bounds__(32) nw_hop_kernel(const TraceArgs a, int By, int cut_lag, int cut_slots)      // cut_lag = 31*K, cut_slots = R+2
{
    __shared__ __align__(16) int win[kHopAhead][kHopWin];
    __shared__ int wbase[kHopAhead];
    // segmented maps: the cut labels of the segment LEFT of the predicted crossing travel with the window (a path that crosses a cut --
    // at 512-row bands nearly every band's does -- otherwise pays a dependent L2 load per crossing: 0.61 ms of the 200k^2 traceback)
    constexpr int kCutMax = 32 * 18;                          // labels per (band, segment): 32 lanes x (R + 2), R <= 16
    __shared__ __align__(16) int cutwin[kHopAhead][kCutMax];
    __shared__ int cutseg[kHopAhead];
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    if (a.phase > 0 && __ldcg(a.miss) == 0) return;           // the pass before found the whole path
    const bool corr = a.corr_w > 0;
    const int nunits = a.map_half ? 2 * a.nb : a.nb;         // unit u = map row u; the units of band 0 are never looked up
    const int first = a.map_half ? 2 : 1;
    const int wmax = ((a.m + 31) / 32) * 32 - kHopWin;       // last window start that stays inside a map row
    const bool pre = wmax >= 0;
    // copies the window of unit u that the line from (unit u_now, column j_now) to the origin predicts; ALWAYS commits one group
    auto issue = [&](int u, int j_now, int u_now) {
        if (pre && u >= first) {
            int c0 = (int)((float)j_now * __fdividef((float)(u + 1), (float)(u_now + 1))) - kHopWin / 2;      // a guess: float is plenty
            c0 = c0 < 0 ? 0 : (c0 > wmax ? wmax : c0);
            c0 &= ~3;
            const int slot = u % kHopAhead;
            if (lane == 0) wbase[slot] = c0;
            const int* src = a.map + (long long)u * a.ldr + kPadL + c0;
#pragma unroll
            for (int i = 0; i < kHopWin / 128; i++) {
                const int e = (i * 32 + lane) * 4;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(&win[slot][e])), "l"(src + e));
            }
            if (a.cut != nullptr && !a.map_half) {
                int sgp = (c0 + kHopWin / 2 + cut_lag) / (32 * a.snap_chunks) - 1;      // the segment left of the predicted crossing
                if (sgp > a.nseg - 2) sgp = a.nseg - 2;
                if (lane == 0) cutseg[slot] = sgp;
                if (sgp >= 0) {
                    const int nlab = 32 * cut_slots;                                     // (a multiple of 4: cut_slots is even)
                    const int* cs = a.cut + ((long long)u * a.nseg + sgp) * nlab;
                    for (int e = 4 * lane; e < nlab; e += 128)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(&cutwin[slot][e])), "l"(cs + e));
                }
            }
        } else if (lane == 0 && u >= 0) {
            cutseg[u % kHopAhead] = -1;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // map[u][col] once the window of unit u has landed
    auto lookup = [&](int u, int col) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kHopAhead - 1) : "memory");
        __syncwarp();
        int v;
        const int slot = u % kHopAhead;
        const int rel = col - wbase[slot];
        if (pre && rel >= 0 && rel < kHopWin) v = win[slot][rel];
        else v = __ldcg(a.map + (long long)u * a.ldr + kPadL + col);
        __syncwarp();
        return v;
    };
    int j = a.m;
    long long off = 0;
    int u = nunits - 1;                                       // next unit to look up
    for (int k = 0; k < kHopAhead; k++) issue(u - k, j, u);
    for (int b = a.nb - 1; b >= 0; b--) {
        if (lane == 0) a.entry[b] = j;
        int jn = 0;
        if (b > 0) {
            if (a.map_half) {                         // through the lower half to the band's middle row, then through the upper half
                int jm = 0;
                if (j > 0) jm = lookup(2 * b + 1, j - 1);
                issue(2 * b + 1 - kHopAhead, jm, 2 * b);
                if (jm > 0) jn = lookup(2 * b, jm - 1);
                issue(2 * b - kHopAhead, jn, 2 * b - 1);
            } else {
                if (j > 0) {
                    int sg = ((j - 1) + cut_lag) / (32 * a.snap_chunks);
                    if (sg > a.nseg - 1) sg = a.nseg - 1;
                    int lo = 0, hi = a.nseg - 1;
                    if (corr) corridor_range(a, b, By, cut_lag, lo, hi);
                    bool out = sg < lo || sg > hi;                      // the path is outside the corridor: the full pass takes over
                    if (!out) {
                        jn = lookup(b, j - 1);
                        // a negative label: the path left the segment that computed it through its left cut -- follow the cut cell's
                        // label in the segment(s) to the left until it is a top-row column (a dependent load each)
                        while (jn < 0 && sg > 0) {
                            sg--;
                            if (sg < lo) { out = true; break; }
                            const int slot = b % kHopAhead;
                            if (pre && cutseg[slot] == sg) jn = cutwin[slot][-jn - 1];          // (the window's group has landed: lookup waited for it)
                            else jn = __ldcg(a.cut + ((long long)b * a.nseg + sg) * 32 * cut_slots + (-jn - 1));
                        }
                    }
                    if (out) {
                        asm volatile("cp.async.wait_all;" ::: "memory");
                        if (lane == 0) { a.miss[0] = 1; a.miss[1] += 1; }
                        return;
                    }
                    if (jn < 0) jn = 0;               // cannot happen (segment 0 has no cut)
                    if (jn > j) jn = j;               // cannot happen with this pair's headers (the walkers' check then fails)
                }
                issue(b - kHopAhead, jn, b - 1);
            }
        }
        if (lane == 0) a.off[b] = off;
        off += (long long)By + (j - jn) + 4;
        j = jn;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (lane == 0) { a.off[a.nb] = off; a.miss[0] = 0; if (a.phase == 0) a.miss[1] = 0; }
}

// ---------------------------------------------------------------------------------------------- pass C
template <int R, int K>
__global__ void __launch_bounds__(32) nw_walk_kernel(const TraceArgs a, const int minw_chunks, const int seg_chunks)
{
    using SC = Sched<R, K>;
    constexpr int By = SC::By, LAG = SC::LAG, PD = SC::PD, VR = SC::VR, XR = SC::XR;
    constexpr int DB = R / 4;                                  // bytes of move codes per lane and step
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned sp_tab[kSpWords];
    stage_sprime(sp_tab, a.sprime, a.S, smem_raw);
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x;
    WarpSmem<R, K> sm(smem_raw, a.S);
    unsigned char* dirs = smem_raw + SC::warp_smem_bytes(a.S);
    const int m = a.m;
    const unsigned ZOFF = (unsigned)a.S * SC::LSTRIDE;
    const long long prow0 = (long long)b * By + (long long)lane * R;
    unsigned yl[R], yoff[R];
    build_profile<R, K>(sm, sp_tab, a.S, a.y, prow0 - a.pad, a.n, lane, yl);
#pragma unroll
    for (int r = 0; r < R; r++) yoff[r] = yl[r] * SC::LSTRIDE;
    const unsigned long long* hr_in = (b > 0) ? a.HR + (long long)b * a.ldr + kPadL : nullptr;
    const int rmin = (b == 0) ? a.pad : 0;        // first real row of the band (band 0 starts with the padding rows)
    unsigned char* out = a.ops + a.off[b];
    int j = a.entry[b];
    int row = By - 1;
    int cnt = 0;
    if (j < 0 || j > m) {                         // a map entry that is no column (headers that are not this pair's): reported by the pack kernel
        if (lane == 0) { a.cnt[b] = 0; a.exitj[b] = -1; }
        return;
    }
    int guard = m / 32 + By + 16;                 // every pass below consumes at least one column or one row

    for (;;) {
        if (--guard < 0) { j = -1; break; }
        if (j == 0) {                             // column 0: the path goes straight up (nwtrace1_plain.cpp:57-63 with j == 0)
            const int k = row - rmin + 1;
            for (int t = lane; t < k; t += 32) out[cnt + t] = 2;
            cnt += k > 0 ? k : 0;
            break;
        }
        if (row < rmin) {                         // left the band through its top row
            if (b == 0) {                         // matrix row 0: the rest of the path runs left along it
                for (int t = lane; t < j; t += 32) out[cnt + t] = 3;
                cnt += j;
                j = 0;
            }
            break;
        }
        // ---- recompute the window [32*lc0 .. j-1] of the band with move codes
        const int lc_hi = ((j - 1) + LAG) / 32;
        int lc0 = 0;
        if (lc_hi - minw_chunks >= a.snap_chunks) {
            int k = (lc_hi - minw_chunks) / a.snap_chunks;
            if (k > a.nsnap) k = a.nsnap;
            lc0 = k * a.snap_chunks;
        }
        if (lc_hi - lc0 + 1 > seg_chunks) lc0 = lc_hi - seg_chunks + 1;    // cannot happen with valid snapshots; keeps smem in bounds
        Lane<R, 2> st;
        st.oprev = 0; st.oup_next = 0; st.o[0] = 0;
        __syncwarp();
        for (int g = lc0 - 2; g < lc0 + PD; g++) load_letter_group<R, K>(sm, a.x, g, m, lane, ZOFF);
        for (int i = lane; i < VR; i += 32) sm.rin[i] = 0;
        __syncwarp();
        for (int g = lc0; g < lc0 + PD; g++) load_top_group<R, K>(sm, hr_in, g, m, lane);
        __syncwarp();
        if (lc0 > 0) {
            const int* sp = a.snap + (((long long)b * a.nsnap + (lc0 / a.snap_chunks - 1)) * 32 + lane) * SC::SNAP_INTS;
#pragma unroll
            for (int r = 0; r < R; r++) st.h[r] = sp[r];
            st.dprev = sp[R];
            st.up_next = (lane == 0) ? sm.rin[(32 * lc0) & (VR - 1)] : sp[R + 1];      // (see nw_map_kernel)
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) st.h[r] = 0;
            st.dprev = 0;
            st.up_next = (lane == 0) ? sm.rin[0] : 0;
        }
        ChunkIO io;
        io.prof_lane = sm.prof + lane * 4 * SC::WPL;
        io.rout_chunk = nullptr; io.rmid_chunk = nullptr; io.map_out = nullptr; io.org0 = 0; io.negg = a.negg;
        for (int lc = lc0; lc <= lc_hi; lc++) {
            const int cp = 32 * (lc + PD) + lane;
            const int pf_top = (hr_in != nullptr && cp < m) ? (int)(unsigned)__ldg(hr_in + cp) : 0;
            const unsigned pf_x = (cp < m) ? (unsigned)__ldg(a.x + cp) : (unsigned)a.S;      // scaled when it lands
            io.xs_lane = sm.xs + ((32 * lc - K * lane) & (XR - 1));
            io.rin_chunk = sm.rin + ((32 * lc) & (VR - 1));
            io.rin_next = sm.rin + ((32 * lc + 32) & (VR - 1));
            io.dirs_lane = dirs + ((size_t)(32 * (lc - lc0)) * 32 + lane) * DB;
            sweep_chunk<R, K, 2>(st, lane, io, yoff);
            __syncwarp();
            sm.rin[cp & (VR - 1)] = pf_top;
            sm.put_letter(cp, pf_x * SC::LSTRIDE);
            __syncwarp();
        }
        // ---- walk the window (uniform across the warp; lane 0 stores)
        const int cmin = 32 * lc0;                // columns >= cmin are complete in this window (all columns if lc0 == 0)
        const int sbase = 32 * lc0;
        while (j > 0 && row >= rmin && (j - 1) >= cmin) {
            const int ln = row / R, r = row % R;
            const int step = (j - 1) + K * ln - sbase;
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
        __syncwarp();
    }
    if (lane == 0) { a.cnt[b] = cnt; a.exitj[b] = j; }
}

// ---------------------------------------------------------------------------------------------- pass D
// One CTA: exclusive scan of the per-band counts in backward path order, then a coalesced gather.
__global__ void __launch_bounds__(1024) nw_pack_kernel(const TraceArgs a)
{
    __shared__ long long s_base[1024];
    __shared__ long long s_carry;
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    if (tid == 0) { s_carry = 0; s_bad = 0; }
    __syncthreads();
    for (int k0 = 0; k0 < a.nb; k0 += 1024) {
        const int k = k0 + tid;                          // k-th band in backward order
        const int b = a.nb - 1 - k;
        const long long c = (k < a.nb) ? a.cnt[b] : 0;
        if (k < a.nb && b > 0 && a.exitj[b] != a.entry[b - 1]) s_bad = 1;     // the walker must agree with the map
        s_base[tid] = c;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            long long v = (tid >= d) ? s_base[tid - d] : 0;
            __syncthreads();
            s_base[tid] += v;
            __syncthreads();
        }
        const long long incl = s_base[tid];
        __syncthreads();
        s_base[tid] = s_carry + incl - c;                // exclusive base of band k in the dense list
        __syncthreads();
        if (tid == 1023) s_carry += incl;
        // one warp per band, coalesced
        const int lane = tid & 31, w = tid >> 5;
        for (int kk = w; kk < 1024 && k0 + kk < a.nb; kk += 32) {
            const int bb = a.nb - 1 - (k0 + kk);
            const unsigned char* src = a.ops + a.off[bb];
            const long long base = s_base[kk];
            const int cc = a.cnt[bb];
            for (int t = lane; t < cc; t += 32) a.dense[base + t] = src[t];
        }
        __syncthreads();
    }
    if (tid == 0) {
        a.total[0] = s_carry; a.total[1] = s_bad;
        a.total[2] = (long long)ld_relaxed64(a.score_elem);
        a.total[3] = (long long)g_wait_timeout;
    }
}


// ---------------------------------------------------------------------------------------------- headers of a cross-GPU fill
// The rank that walks the path pulls what the traceback reads from the other ranks' buffers (mapped with CUDA IPC, loads over NVLink):
// for band b the header row above it and its snapshots, in the columns of the traceback's corridor (+ slack for the windows of the
// map segments and walkers) -- or in all columns (`full`: after a corridor miss).  Column c of every row lives on rank
// (c / wc) % world; rows and snapshots have the layout of the whole matrix on every rank.  One CTA per band.
struct GatherArgs {
    unsigned long long* HR;                  // this rank's header rows
    int* snap;                               // this rank's snapshots
    const unsigned long long* peer_HR[16];   // the other ranks' (nullptr: this rank)
    const int* peer_snap[16];
    int rank, world;
    long long wc;                            // columns per block
    int n, m, nb, pad, By, Bx;
    long long ldr;
    int nsnap, snap_ints;                    // snapshots per band, ints per snapshot (32 lanes x SNAP_INTS)
    long long d, slack;                      // corridor half width (0: everything), slack on either side
    int full;
};

__global__ void __launch_bounds__(256) nw_gather_kernel(const GatherArgs a)
{
    const int b = blockIdx.x;                // 0 .. nb-1: the band whose top row (b >= 1) and snapshots are pulled; nb: the score element's row
    if (b > a.nb) return;
    long long c_lo = 0, c_hi = a.m;          // columns [c_lo, c_hi)
    if (b == a.nb) { c_lo = a.m - 1; c_hi = a.m; }
    else if (!a.full && a.d > 0) {
        long long rb = (long long)(b + 1) * a.By - a.pad, rt = (long long)b * a.By - a.pad;
        if (rb > a.n) rb = a.n;
        if (rt < 0) rt = 0;
        c_lo = (long long)a.m * rt / a.n - a.d - a.slack;
        c_hi = (long long)a.m * rb / a.n + a.d + a.slack;
        if (c_lo < 0) c_lo = 0;
        if (c_hi > a.m) c_hi = a.m;
    }
    // ---- header row b
    unsigned long long* dst = a.HR + (long long)b * a.ldr + kPadL;
    for (long long c = c_lo + threadIdx.x; b >= 1 && c < c_hi; c += blockDim.x) {
        const int owner = (int)((c / a.wc) % a.world);
        const unsigned long long* src = a.peer_HR[owner];
        if (src != nullptr) dst[c] = src[(long long)b * a.ldr + kPadL + c];
    }
    if (b == a.nb) return;
    // ---- snapshots of band b whose boundary column (k + 1) * Bx lies in the range (snapshot k belongs to the block of column k * Bx)
    long long k_lo = c_lo / a.Bx - 1, k_hi = c_hi / a.Bx + 1;
    if (k_lo < 0) k_lo = 0;
    if (k_hi > a.nsnap) k_hi = a.nsnap;
    for (long long k = k_lo; k < k_hi; k++) {
        const int owner = (int)(((k * a.Bx) / a.wc) % a.world);
        const int* src = a.peer_snap[owner];
        if (src == nullptr) continue;
        const long long off = ((long long)b * a.nsnap + k) * a.snap_ints;
        const int4* s4 = reinterpret_cast<const int4*>(src + off);
        int4* d4 = reinterpret_cast<int4*>(a.snap + off);
        for (int i = threadIdx.x; i < a.snap_ints / 4; i += blockDim.x) d4[i] = s4[i];
    }
}

// ---------------------------------------------------------------------------------------------- score matrix slabs
// Re-sweeps bands b0 .. b0+nbands-1 from their header rows and writes EVERY cell (shifted value P) to a row-major
// slab: slab[(b-b0)*By + local row][kPadL + c].  Used by the score hash (NwHash1_Plain / NwHash2_Sparse semantics,
// nwtrace1_plain.cpp:133-154: the fold itself is sequential and done by the host as the slabs stream back) and by
// the header export.  Bands are independent once the fill is done, so this runs at full issue rate.
struct DumpArgs {
    const uint8_t* y;
    const uint8_t* x;
    int n, m;
    const uint8_t* sprime;
    int S;
    const unsigned long long* HR;
    long long ldr;
    int pad;
    int b0, nbands;
    int* slab;
    long long ld;          // slab row pitch in ints (>= kPadL + 32*nlc)
};

template <int R, int K, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) nw_dump_kernel(const DumpArgs a)
{
    using SC = Sched<R, K>;
    constexpr int By = SC::By, PD = SC::PD, VR = SC::VR, XR = SC::XR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned sp_tab[kSpWords];
    stage_sprime(sp_tab, a.sprime, a.S, smem_raw);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    WarpSmem<R, K> sm(smem_raw + (size_t)w * SC::warp_smem_bytes(a.S), a.S);
    const int m = a.m, nlc = SC::nlc(m);
    const unsigned ZOFF = (unsigned)a.S * SC::LSTRIDE;
    const int nwarps = gridDim.x * WARPS;
    for (int bi = blockIdx.x * WARPS + w; bi < a.nbands; bi += nwarps) {
        const int b = a.b0 + bi;
        const long long prow0 = (long long)b * By + (long long)lane * R;
        build_profile<R, K>(sm, sp_tab, a.S, a.y, prow0 - a.pad, a.n, lane, nullptr);
        const unsigned long long* hr_in = (b > 0) ? a.HR + (long long)b * a.ldr + kPadL : nullptr;
        for (int g = -2; g < PD; g++) load_letter_group<R, K>(sm, a.x, g, m, lane, ZOFF);
        for (int i = lane; i < VR; i += 32) sm.rin[i] = 0;
        __syncwarp();
        for (int g = 0; g < PD; g++) load_top_group<R, K>(sm, hr_in, g, m, lane);
        __syncwarp();
        Lane<R, 3> st;
#pragma unroll
        for (int r = 0; r < R; r++) st.h[r] = 0;
        st.dprev = 0; st.oprev = 0; st.oup_next = 0; st.o[0] = 0;
        st.up_next = (lane == 0) ? sm.rin[0] : 0;
        ChunkIO io;
        io.prof_lane = sm.prof + lane * 4 * SC::WPL;
        io.rout_chunk = nullptr; io.rmid_chunk = nullptr; io.map_out = nullptr; io.org0 = 0; io.dirs_lane = nullptr; io.negg = 0;
        io.dump_ld = a.ld;
        int* slab_lane = a.slab + ((long long)bi * By + (long long)lane * R) * a.ld + kPadL - K * lane;
        for (int lc = 0; lc < nlc; lc++) {
            const int cp = 32 * (lc + PD) + lane;
            const int pf_top = (hr_in != nullptr && cp < m) ? (int)(unsigned)__ldg(hr_in + cp) : 0;
            const unsigned pf_x = (cp < m) ? (unsigned)__ldg(a.x + cp) : (unsigned)a.S;      // scaled when it lands
            io.xs_lane = sm.xs + ((32 * lc - K * lane) & (XR - 1));
            io.rin_chunk = sm.rin + ((32 * lc) & (VR - 1));
            io.rin_next = sm.rin + ((32 * lc + 32) & (VR - 1));
            io.dump_lane = slab_lane + 32 * lc;
            sweep_chunk<R, K, 3>(st, lane, io, nullptr);
            __syncwarp();
            sm.rin[cp & (VR - 1)] = pf_top;
            sm.put_letter(cp, pf_x * SC::LSTRIDE);
            __syncwarp();
        }
    }
}


// ---------------------------------------------------------------------------------------------- header export
// Reference layout (producer nwalign_gpu9_mlsp_diagdiagdiag.cu:321-359, consumer
// nwtrace2_sparse.cpp:48-67), tiles of By x Bx aligned to the TOP-LEFT corner like the reference's:
//   hrow[(iT*tcols + jT)*(1+Bx) + k] = H[iT*By][jT*Bx + k]      hcol[(iT*tcols + jT)*(1+By) + k] = H[iT*By + k][jT*Bx]
// with H = P + (i+j)*gap.  Entries outside the real matrix are written as 0 (nothing consumes them).
__global__ void nw_export_hrow_kernel(const unsigned long long* __restrict__ HR, long long ldr, int n, int m, int By, int Bx,
                                      int trows, int tcols, int gap, int* __restrict__ hrow)
{
    const long long total = (long long)trows * tcols * (1 + Bx);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long t = e / (1 + Bx); const int k = (int)(e % (1 + Bx));
        const int iT = (int)(t / tcols), jT = (int)(t % tcols);
        const long long i = (long long)iT * By, j = (long long)jT * Bx + k;
        int v = 0;
        if (i <= n && j <= m) {
            const int P = (i == 0 || j == 0) ? 0 : (int)(unsigned)HR[(long long)iT * ldr + kPadL + (j - 1)];
            v = P + (int)((i + j) * gap);
        }
        hrow[e] = v;
    }
}
// rows [r0, r0+nrows) of the matrix (1-based i = r0+1 ..) are present in `slab` (row-major, pitch ld, column c at kPadL + c)
__global__ void nw_export_hcol_kernel(const int* __restrict__ slab, long long ld, long long r0, long long nrows, int n, int m, int By, int Bx,
                                      int trows, int tcols, int gap, int* __restrict__ hcol, int first)
{
    const long long total = (long long)trows * tcols * (1 + By);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long t = e / (1 + By); const int k = (int)(e % (1 + By));
        const int iT = (int)(t / tcols), jT = (int)(t % tcols);
        const long long i = (long long)iT * By + k, j = (long long)jT * Bx;
        if (i == 0 || j == 0 || i > n || j > m) { if (first) hcol[e] = (i <= n && j <= m) ? (int)((i + j) * gap) : 0; continue; }
        const long long lr = (i - 1) - r0;
        if (lr < 0 || lr >= nrows) continue;
        hcol[e] = slab[lr * ld + kPadL + (j - 1)] + (int)((i + j) * gap);
    }
}

}  // namespace nwb
