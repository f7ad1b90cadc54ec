// nw_fill.cuh -- score-matrix fill for ONE pair: the replacement of Nw_Gpu9_KernelA/KernelB
// (reference nwalign_gpu9_mlsp_diagdiagdiag.cu:15-63,69-360) and of the per-diagonal CUDA graph that
// drives them (:555-684).
//
//   * One persistent launch.  Warps are independent dataflow agents: a warp takes a band from an atomic ticket
//     (so the warp that owns band b-1 is always running or done when band b starts) and sweeps it over all
//     columns (nw_sweep.cuh).  The only coupling between bands is the bottom row of band b-1, which is the top
//     row of band b -- and IS the tile header row the traceback needs anyway.  It is streamed through L2 as
//     64-bit (epoch-tag | value) elements, one per step from lane 31: data and ready flag arrive in one
//     naturally atomic store, the consumer prefetches two 32-column groups ahead and never fences.  There is
//     no launch per diagonal, no __syncthreads, no cooperative launch, no global barrier.
//   * For the traceback the warp also drops a SNAPSHOT of its register state every snap_chunks*32 steps
//     (R+2 ints per lane, at a chunk boundary: nothing is checked per step).  A snapshot is the engine's
//     "header column": the walker resumes the sweep from it to recompute a window of the band
//     (the reference keeps rectangular header columns instead, nwalign_gpu9...cu:296-300,340-359).
//   * Column blocks (cross-GPU wavefront for one very long pair): the columns are dealt to the ranks in blocks of wc
//     (block g belongs to rank g % world).  A (block, band) unit starts from the right border column of the block to its
//     left, which the owner of that block PUSHES into this rank's receive buffer with plain peer stores over NVLink
//     followed by a system-scope release flag (st.release.sys); it pushes its own right border on to the next rank.
//     Tickets run block-major, so one persistent launch per GPU pipelines all of its blocks: the wavefront skew is
//     paid once, not once per block.  With world == 1 the receive buffer is the GPU's own (the same code path).
#pragma once
#include "nw_sweep.cuh"

namespace nwb {

struct FillArgs {
    const uint8_t* y;        // lenY letters
    const uint8_t* x;        // ALL lenX letters
    int n, m;                // lenY, lenX
    const uint8_t* sprime;   // S*S bytes: s'[y*S+x]
    int S;                   // alphabet size (<= kMaxLetters)
    unsigned long long* HR;  // header rows of column block q: HR[q*hr_stride + b*ldr + kPadL + c] = (tag << 32 | P[top row of band b][c0+c+1]), b = 1..nb
    long long ldr;           // >= kPadL + 32*nlc(block width) + 32
    long long hr_stride;     // elements between the header blocks of two column blocks
    int* snap;               // snapshots: snap[((b*nsnap + k)*32 + lane)*SNAP_INTS + i] after chunk (k+1)*snap_chunks-1 (nullable; single block only)
    int nsnap;               // snapshots per band
    int snap_chunks;         // chunks between snapshots
    // ---- column blocks (cross-GPU wavefront): this rank owns blocks rank, rank+world, ... of width wc
    int wc;                  // block width in columns (multiple of 32); single GPU, single block: wc >= m
    int nq;                  // blocks owned by this rank
    int rank, world;
    int* recv;               // recv[q*recv_stride + 1 + padded row] = P of the column left of block q (recv[q*recv_stride] = row above), written by the left neighbour
    unsigned* recv_flag;     // recv_flag[q*nb + b] == tag when band b of that column is complete
    int* peer_recv;          // the RIGHT neighbour's recv / recv_flag (peer memory mapped over NVLink; own buffers when world == 1)
    unsigned* peer_flag;
    long long recv_stride;
    int* lastcol;            // nullable: lastcol[1 + padded row] = P[row][m] (last column of the matrix), written by the owner of the last block
    unsigned long long timeout_ns;   // give up waiting for a neighbour after this long (0 = never)
    int* err;                // set to 1 on timeout
    unsigned tag;            // epoch tag of this run's header rows (never 0; local to this GPU)
    unsigned xtag;           // epoch tag of the cross-GPU border flags (the same on every rank)
    int* ticket;             // (block, band) ticket counter
    int nb;                  // number of bands
    int pad;                 // padding rows above row 1 in band 0 (nb*By - n)
    unsigned long long* dbg; // developer aid: [nb][4] globaltimer stamps (start, prologue done, end) + poll count; nullable
    int dbg_mode;            // developer aid: 1 = consumers do not wait (timing experiment, wrong results)
    int slack;               // groups of head start a consumer gives its producer before it starts
    int pd;                  // header-row groups prefetched ahead (2: the K == 2 look-ahead reads the first element of the next group)
    int* map;                // nullable (single block only): origin maps for the traceback, computed IN THIS LAUNCH by extra
                             // warps that follow the fill one band behind (map[b*ldr + kPadL + c], see nw_trace.cuh pass A)
    int negg;                // -gap (map units)
    unsigned long long* MID; // nullable: MID[b*ldr + kPadL + c] = (tag << 32 | P[middle row of band b][c+1]) -- the bottom row of lane 15,
                             // published like HR so that the map units can work on HALF bands (two rows per lane: the step of a
                             // map unit then costs what a fill step costs and the maps finish right behind the fill)
    int map_half;            // 1: map[(2b+h)*ldr + ...] is the map of half h (0 upper, 1 lower) of band b; 0: map[b*ldr + ...] whole bands
    int grouped;             // 1 (single block): the W warps of a CTA take W CONSECUTIVE bands and hand the header row from warp to
                             // warp through shared memory (3 of 4 hand-offs no longer cost an L2 round trip); CTAs take tickets
    int map_inline;          // 0: map units shadow the fill units on otherwise idle SM sub-partitions (few bands);
                             // 1: every band is swept ONCE with origin labels and publishes its header row itself (many bands)
};

__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Origin map of one band or half band (traceback pass A, see nw_trace.cuh) computed inside the fill launch: the warp
// follows the fill, consuming the same tagged rows the fill publishes.  `hr_in` = the row above the unit's rows (nullptr:
// row 0 of the matrix), `prow_base` = padded index of its first row, `map_row` = where lane 31's labels go,
// `b_out` = band whose header row / snapshots this unit publishes itself (publish only: the unit then IS the band's fill).
template <int R, int K>
__device__ __forceinline__ void map_unit(const FillArgs& a, unsigned char* warp_smem, const unsigned* sp_tab, const unsigned long long* hr_in,
                                         const long long prow_base, int* map_row, const int b_out, const int lane, const bool publish)
{
    using SC = Sched<R, K>;
    constexpr int LAG = SC::LAG, VR = SC::VR, XR = SC::XR;
    WarpSmem<R, K> sm(warp_smem, a.S);
    const int b = b_out;
    const int PD = a.pd;
    const int m = a.m, nlc = SC::nlc(m);
    const unsigned ZOFF = (unsigned)a.S * SC::LSTRIDE;
    const long long prow0 = prow_base + (long long)lane * R;
    build_profile<R, K>(sm, sp_tab, a.S, a.y, prow0 - a.pad, a.n, lane, nullptr);
    for (int c = -64 + lane; c < 0; c += 32) sm.put_letter(c, ZOFF);
    for (int g = 0; g < PD; g++) {
        const int c = 32 * g + lane;
        sm.put_letter(c, c < m ? (unsigned)__ldg(a.x + c) * SC::LSTRIDE : ZOFF);
    }
    for (int i = lane; i < VR; i += 32) sm.rin[i] = 0;
    __syncwarp();
    const bool consumer = hr_in != nullptr;
    unsigned long long* hr_out = a.HR + (long long)(b + 1) * a.ldr + kPadL;
    if (consumer) {
        int c = 32 * (PD - 1 + a.slack) + 31;
        if (c > m - 1) c = m - 1;
        (void)wait_tagged(hr_in + c, ld_relaxed64(hr_in + c), a.tag);
        for (int g = 0; g < PD; g++) {
            const int cc = 32 * g + lane;
            if (cc < m) sm.rin[cc & (VR - 1)] = wait_tagged(hr_in + cc, ld_relaxed64(hr_in + cc), a.tag);
        }
    }
    __syncwarp();
    Lane<R, 1> st;
#pragma unroll
    for (int r = 0; r < R; r++) { st.h[r] = 0; st.o[r] = 0; }
    st.dprev = 0; st.oprev = 0;
    st.up_next = (lane == 0) ? sm.rin[0] : 0;
    st.oup_next = (lane == 0) ? 1 : 0;
    ChunkIO io;
    io.prof_lane = sm.prof + lane * 4 * SC::WPL;
    io.rout_chunk = nullptr; io.rmid_chunk = nullptr; io.dirs_lane = nullptr; io.negg = a.negg; io.dump_lane = nullptr; io.dump_ld = 0;
    int snap_left = a.snap_chunks, snap_k = 0;
    for (int lc = 0; lc < nlc; lc++) {
        const int cp = 32 * (lc + PD) + lane;
        unsigned long long pf_hr = 0;
        const bool want_hr = consumer && cp < m;
        if (want_hr) pf_hr = ld_relaxed64(hr_in + cp);
        const unsigned pf_x = (cp < m) ? (unsigned)__ldg(a.x + cp) : (unsigned)a.S;
        io.xs_lane = sm.xs + ((32 * lc - K * lane) & (XR - 1));
        io.rin_chunk = sm.rin + ((32 * lc) & (VR - 1));
        io.rin_next = sm.rin + ((32 * lc + 32) & (VR - 1));
        io.map_out = map_row + (32 * lc - LAG);
        io.org0 = 32 * lc + 1;
        io.rout_chunk = publish ? sm.rout + (lc & 1) * 32 : nullptr;
        sweep_chunk<R, K, 1>(st, lane, io, nullptr);
        __syncwarp();
        if (publish && lc >= SC::GL) {
            const int v = (lane >= SC::SH) ? sm.rout[(lc & 1) * 32 + lane - SC::SH] : sm.rout[((lc + 1) & 1) * 32 + 32 - SC::SH + lane];
            st_relaxed64(hr_out + 32 * (lc - SC::GL) + lane, pack_tagged(v, a.tag));
        }
        sm.rin[cp & (VR - 1)] = want_hr ? wait_tagged(hr_in + cp, pf_hr, a.tag) : 0;
        sm.put_letter(cp, pf_x * SC::LSTRIDE);
        if (--snap_left == 0) {
            snap_left = a.snap_chunks;
            const int k = snap_k++;
            if (publish && a.snap != nullptr && k < a.nsnap) {
                int* sp = a.snap + (((long long)b * a.nsnap + k) * 32 + lane) * SC::SNAP_INTS;
#pragma unroll
                for (int r = 0; r < R; r += 4)
                    st_cs4(reinterpret_cast<int4*>(sp + r), make_int4(st.h[r], st.h[r + 1], st.h[r + 2], st.h[r + 3]));
                st_cs4(reinterpret_cast<int4*>(sp + R), make_int4(st.dprev, st.up_next, 0, 0));
            }
        }
        __syncwarp();
    }
    if (publish && lane < SC::SH)
        st_relaxed64(hr_out + 32 * (nlc - SC::GL) + lane, pack_tagged(sm.rout[((nlc - 1) & 1) * 32 + 32 - SC::SH + lane], a.tag));
    __syncwarp();
}

// Shared-memory hand-off of the header row between the warps of one CTA (grouped mode).
struct Handoff {
    volatile int* ready_in;    // groups the warp above has put into THIS warp's top-row ring (nullptr: top row comes from HR in HBM)
    volatile int* cons_out;    // this warp's chunk counter, read by the warp above for back-pressure
    int* next_rin;             // the top-row ring of the warp below (nullptr: nobody below in this CTA)
    volatile int* ready_out;   // groups this warp has put there
    volatile int* cons_in;     // the chunk counter of the warp below
};

__device__ __forceinline__ void spin_until_ge(volatile int* p, int need)
{
    unsigned polls = 0;
    while (*p < need) {
        __nanosleep(20);
        if (++polls > (1u << 26)) { g_wait_timeout = 1; break; }
    }
}

// Fill of band b of column block q (one warp).
template <int R, int K>
__device__ __forceinline__ void fill_unit(const FillArgs& a, unsigned char* warp_smem, const unsigned* sp_tab, const int t, const int lane,
                                          const bool half_map, const Handoff hand)
{
    using SC = Sched<R, K>;
    constexpr int By = SC::By, VR = SC::VR, XR = SC::XR;
    WarpSmem<R, K> sm(warp_smem, a.S);
    const int PD = a.pd;
    const unsigned ZOFF = (unsigned)a.S * SC::LSTRIDE;      // profile offset of the all-zero row
    const int nblocks = (a.m + a.wc - 1) / a.wc;             // column blocks of the whole matrix
    const bool smem_in = hand.ready_in != nullptr, smem_out = hand.next_rin != nullptr;
    const int q = t / a.nb, b = t - q * a.nb;             // tickets run block-major: (q, b-1) is always taken before (q, b)
    const int gb = q * a.world + a.rank;                  // global column block
    const long long c0 = (long long)gb * a.wc;            // its first column
    const int m = (int)((a.m - c0 < a.wc) ? a.m - c0 : a.wc);
    const int nlc = SC::nlc(m);
    const uint8_t* xb = a.x + c0;
    const bool has_left = gb > 0, has_right = gb + 1 < nblocks;
    if (a.dbg && lane == 0 && q == 0) a.dbg[4 * b + 0] = globaltimer_ns();
    unsigned spins = 0;

    const long long prow0 = (long long)b * By + (long long)lane * R;      // padded row index of this lane's first row
    build_profile<R, K>(sm, sp_tab, a.S, a.y, prow0 - a.pad, a.n, lane, nullptr);
    // letter ring: columns -64..-1 use the zero row, groups 0..PD-1 are loaded now
    for (int c = -64 + lane; c < 0; c += 32) sm.put_letter(c, ZOFF);
    for (int g = 0; g < PD; g++) {
        const int c = 32 * g + lane;
        sm.put_letter(c, c < m ? (unsigned)__ldg(xb + c) * SC::LSTRIDE : ZOFF);
    }
    if (!smem_in) for (int i = lane; i < VR; i += 32) sm.rin[i] = 0;      // grouped mode: cleared before the warp above could write into it

    Lane<R, 0> st;
    st.oprev = 0; st.oup_next = 0; st.o[0] = 0;
    if (has_left) {
        // the column left of this block arrives from the left neighbour (peer stores over NVLink + system-scope flag)
        // (band b-1's flag as well: this lane 0 reads the last row of band b-1 as its diagonal input)
        const unsigned* fl = a.recv_flag + (long long)q * a.nb + b;
        const unsigned long long t0 = a.timeout_ns ? globaltimer_ns() : 0ull;
        while (ld_acquire_sys_u32(fl) != a.xtag || (b > 0 && ld_acquire_sys_u32(fl - 1) != a.xtag)) {
            __nanosleep(200);
            if (a.timeout_ns && globaltimer_ns() - t0 > a.timeout_ns) { if (lane == 0) atomicExch(a.err, 1); break; }
        }
        const int* lp = a.recv + (long long)q * a.recv_stride + 1 + prow0;
#pragma unroll
        for (int r = 0; r < R; r++) st.h[r] = ld_volatile(lp + r);
        st.dprev = ld_volatile(lp - 1);
    } else {
#pragma unroll
        for (int r = 0; r < R; r++) st.h[r] = 0;
        st.dprev = 0;
    }
    const bool consumer = (b > 0);                 // band 0 has row 0 (P = 0) above it
    const unsigned long long* hr_in = a.HR + (long long)q * a.hr_stride + (long long)b * a.ldr + kPadL;
    unsigned long long* hr_out = a.HR + (long long)q * a.hr_stride + (long long)(b + 1) * a.ldr + kPadL;
    unsigned long long* mid_out = (half_map && b > 0) ? a.MID + (long long)b * a.ldr + kPadL : nullptr;
    constexpr int GLM = (15 * K + 31) / 32, SHM = 32 * GLM - 15 * K;
    __syncwarp();
    // ---- prologue: the first PD groups of the row above
    const int gcap = (m + 31) / 32;                  // groups that hold real columns
    if (smem_in) {
        spin_until_ge(hand.ready_in, PD < gcap ? PD : gcap);
        __syncwarp();
    } else if (consumer) {
        {   // head start for the producer: the consumer's prefetches then only touch lines that are complete
            int c = 32 * (PD - 1 + a.slack) + 31;
            if (c > m - 1) c = m - 1;
            (void)wait_tagged(hr_in + c, ld_relaxed64(hr_in + c), a.tag);
        }
        for (int g = 0; g < PD; g++) {
            const int c = 32 * g + lane;
            if (c < m) sm.rin[c & (VR - 1)] = wait_tagged(hr_in + c, ld_relaxed64(hr_in + c), a.tag);
        }
    }
    __syncwarp();
    if (a.dbg && lane == 0 && q == 0) a.dbg[4 * b + 1] = globaltimer_ns();
    st.up_next = (lane == 0) ? sm.rin[0] : st.dprev;

    ChunkIO io;
    io.prof_lane = sm.prof + lane * 4 * SC::WPL;
    io.map_out = nullptr; io.org0 = 0; io.dirs_lane = nullptr; io.negg = 0; io.dump_lane = nullptr; io.dump_ld = 0; io.rmid_chunk = nullptr;
    int snap_left = a.snap_chunks, snap_k = 0;
    for (int lc = 0; lc < nlc; lc++) {
        // ---- issue the prefetches of chunk lc + PD
        const int cp = 32 * (lc + PD) + lane;
        unsigned long long pf_hr = 0;
        const bool want_hr = consumer && !smem_in && cp < m;
        if (smem_in) {                                  // the warp above writes straight into this warp's ring: wait for groups lc and lc+1
            spin_until_ge(hand.ready_in, lc + 2 < gcap ? lc + 2 : gcap);
            if (lane == 0) *hand.cons_out = lc;
            __syncwarp();
        }
        if (want_hr) pf_hr = ld_relaxed64(hr_in + cp);
        const unsigned pf_x = (cp < m) ? (unsigned)__ldg(xb + cp) : (unsigned)a.S;      // scaled when it lands: nothing waits on the load here
        // ---- the chunk itself
        io.xs_lane = sm.xs + ((32 * lc - K * lane) & (XR - 1));
        io.rin_chunk = sm.rin + ((32 * lc) & (VR - 1));
        io.rin_next = sm.rin + ((32 * lc + 32) & (VR - 1));
        io.rout_chunk = sm.rout + (lc & 1) * 32;
        io.rmid_chunk = mid_out != nullptr ? sm.rmid + (lc & 1) * 32 : nullptr;
        sweep_chunk<R, K, 0>(st, lane, io, nullptr);
        __syncwarp();
        // ---- publish the group of the bottom row that this chunk completed: ONE coalesced 256-byte store
        if (lc >= SC::GL) {
            const int v = (lane >= SC::SH) ? sm.rout[(lc & 1) * 32 + lane - SC::SH] : sm.rout[((lc + 1) & 1) * 32 + 32 - SC::SH + lane];
            st_relaxed64(hr_out + 32 * (lc - SC::GL) + lane, pack_tagged(v, a.tag));
            if (smem_out) {                             // ... and straight into the ring of the warp below (same CTA): no L2 round trip
                const int gi = lc - SC::GL;
                if (gi >= VR / 32 - 1) spin_until_ge(hand.cons_in, gi - (VR / 32 - 1));      // never overwrite a group it still reads
                hand.next_rin[(32 * gi + lane) & (VR - 1)] = v;
                __threadfence_block();
                __syncwarp();
                if (lane == 0) *hand.ready_out = gi + 1;
            }
        }
        // ---- and the group of the middle row (lane 15 is 15*K columns behind lane 0)
        if (mid_out != nullptr && lc >= GLM) {
            const int v = (lane >= SHM) ? sm.rmid[(lc & 1) * 32 + lane - SHM] : sm.rmid[((lc + 1) & 1) * 32 + 32 - SHM + lane];
            st_relaxed64(mid_out + 32 * (lc - GLM) + lane, pack_tagged(v, a.tag));
        }
        // ---- land the prefetches
        if (want_hr) {
            if (a.dbg_mode == 1) sm.rin[cp & (VR - 1)] = (int)(unsigned)pf_hr;
            else sm.rin[cp & (VR - 1)] = wait_tagged_count(hr_in + cp, pf_hr, a.tag, spins);
        }
        else if (consumer && !smem_in) sm.rin[cp & (VR - 1)] = 0;
        sm.put_letter(cp, pf_x * SC::LSTRIDE);
        // ---- snapshot of the register state for the traceback
        if (--snap_left == 0) {
            snap_left = a.snap_chunks;
            const int k = snap_k++;
            if (a.snap != nullptr && k < a.nsnap) {
                int* sp = a.snap + (((long long)b * a.nsnap + k) * 32 + lane) * SC::SNAP_INTS;
#pragma unroll
                for (int r = 0; r < R; r += 4)
                    st_cs4(reinterpret_cast<int4*>(sp + r), make_int4(st.h[r], st.h[r + 1], st.h[r + 2], st.h[r + 3]));
                st_cs4(reinterpret_cast<int4*>(sp + R), make_int4(st.dprev, st.up_next, 0, 0));
            }
        }
        __syncwarp();
    }
    // ---- the first SH elements of the next group were produced by the last chunk (they hold the last real column)
    if (smem_out && nlc - SC::GL >= VR / 32 - 1) spin_until_ge(hand.cons_in, nlc - SC::GL - (VR / 32 - 1));
    if (lane < SC::SH) {
        const int v = sm.rout[((nlc - 1) & 1) * 32 + 32 - SC::SH + lane];
        st_relaxed64(hr_out + 32 * (nlc - SC::GL) + lane, pack_tagged(v, a.tag));
        if (smem_out) hand.next_rin[(32 * (nlc - SC::GL) + lane) & (VR - 1)] = v;
    }
    if (mid_out != nullptr && lane < SHM)
        st_relaxed64(mid_out + 32 * (nlc - GLM) + lane, pack_tagged(sm.rmid[((nlc - 1) & 1) * 32 + 32 - SHM + lane], a.tag));
    if (a.dbg && lane == 0 && q == 0) { a.dbg[4 * b + 2] = globaltimer_ns(); a.dbg[4 * b + 3] = spins; }
    // ---- every row is frozen at its last-column value by now
    if (has_right) {
        // push the right border column into the right neighbour's receive buffer; its round index is q, or q+1 when we
        // are the last rank (the next block then belongs to rank 0's next round)
        const int qn = (a.rank + 1 == a.world) ? q + 1 : q;
        int* lp = a.peer_recv + (long long)qn * a.recv_stride + 1 + prow0;
#pragma unroll
        for (int r = 0; r < R; r++) lp[r] = st.h[r];
        if (b == 0 && lane == 0) lp[-1] = 0;                       // row above the matrix: P = 0
        __syncwarp();
        __threadfence_system();
        if (lane == 0) st_release_sys_u32(a.peer_flag + (long long)qn * a.nb + b, a.xtag);
    } else if (a.lastcol != nullptr) {
        int* lp = a.lastcol + 1 + prow0;
#pragma unroll
        for (int r = 0; r < R; r++) lp[r] = st.h[r];
    }
    if (smem_out) { __threadfence_block(); __syncwarp(); if (lane == 0) *hand.ready_out = 0x7fffffff; }
    __syncwarp();
}

template <int R, int K, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) nw_fill_kernel(const FillArgs a)
{
    using SC = Sched<R, K>;
    constexpr int By = SC::By, LAG = SC::LAG, VR = SC::VR, XR = SC::XR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned sp_tab[kSpWords];
    stage_sprime(sp_tab, a.sprime, a.S);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    WarpSmem<R, K> sm(smem_raw + (size_t)w * SC::warp_smem_bytes(a.S), a.S);
    const int PD = a.pd;
    const bool with_map = a.map != nullptr;
    const bool inline_map = with_map && a.map_inline != 0;
    const bool half_map = with_map && !inline_map && a.map_half != 0;
    // units: fill units only | band 0 fill + (fill, map) per further band | band 0 fill + (fill, upper map, lower map) per further band
    const int nunits = inline_map ? a.nb : (with_map ? (half_map ? 3 * a.nb - 2 : 2 * a.nb - 1) : a.nb * a.nq);
    const unsigned ZOFF = (unsigned)a.S * SC::LSTRIDE;      // profile offset of the all-zero row
    const int nblocks = (a.m + a.wc - 1) / a.wc;             // column blocks of the whole matrix
    unsigned char* warp_smem = smem_raw + (size_t)w * SC::warp_smem_bytes(a.S);

    if (a.grouped) {
        // ---- grouped mode: CTA-level tickets.  Fill group k = bands k*WARPS .. k*WARPS+WARPS-1 (warp w takes band k*WARPS+w and
        // feeds the warp below through shared memory); map CTA j = map units j*WARPS .. (one warp each, inputs from HBM).
        // Order: F0, then for k >= 1: F(k) followed by the map CTAs of group k-1, finally the map CTAs of the last group.
        __shared__ int s_ticket;
        __shared__ volatile int s_ready[WARPS];
        __shared__ volatile int s_cons[WARPS];
        const int nG = (a.nb + WARPS - 1) / WARPS;
        const int mper = with_map ? (half_map ? 2 : 1) : 0;           // map CTAs per fill group
        const int ncta_units = nG * (1 + mper);
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1);
            if (threadIdx.x < WARPS) { s_ready[threadIdx.x] = 0; s_cons[threadIdx.x] = 0; }
            for (int i = lane; i < VR; i += 32) sm.rin[i] = 0;
            __syncthreads();
            const int T = s_ticket;
            if (T >= ncta_units) break;
            // decode
            int fill_k = -1, map_j = -1;
            if (T == 0) fill_k = 0;
            else if (mper == 0) fill_k = T;
            else {
                const int j = T - 1, per = 1 + mper;
                const int k = j / per + 1, r = j % per;
                if (k < nG) { if (r == 0) fill_k = k; else map_j = mper * (k - 1) + (r - 1); }
                else map_j = mper * (nG - 1) + (j - per * (nG - 1));
            }
            if (fill_k >= 0) {
                const int b = fill_k * WARPS + w;
                if (b < a.nb) {
                    Handoff hand;
                    hand.ready_in = (w > 0) ? &s_ready[w - 1] : nullptr;
                    hand.cons_out = &s_cons[w];
                    const bool below = (w + 1 < WARPS) && (b + 1 < a.nb);
                    hand.next_rin = below ? WarpSmem<R, K>(warp_smem + SC::warp_smem_bytes(a.S), a.S).rin : nullptr;
                    hand.ready_out = &s_ready[w];
                    hand.cons_in = below ? &s_cons[w + 1] : nullptr;
                    fill_unit<R, K>(a, warp_smem, sp_tab, b, lane, half_map, hand);
                }
            } else {
                const int u = map_j * WARPS + w;                    // map unit: band u / 2, half u % 2 (half maps) or band u
                const int bb = half_map ? (u >> 1) : u;
                if (bb >= 1 && bb < a.nb) {
                    if (half_map) {
                        const int hh = u & 1;
                        map_unit<R / 2, K>(a, warp_smem, sp_tab, (hh ? a.MID : a.HR) + (long long)bb * a.ldr + kPadL, (long long)bb * By + hh * (By / 2),
                                           a.map + (long long)u * a.ldr + kPadL, bb, lane, false);
                    } else {
                        map_unit<R, K>(a, warp_smem, sp_tab, a.HR + (long long)bb * a.ldr + kPadL, (long long)bb * By,
                                       a.map + (long long)bb * a.ldr + kPadL, bb, lane, false);
                    }
                }
            }
        }
        return;
    }
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(a.ticket, 1);
        t = __shfl_sync(kFull, t, 0);
        if (t >= nunits) break;
        if (inline_map) {                                     // one sweep per band: values, origin labels, header row, snapshots
            map_unit<R, K>(a, warp_smem, sp_tab, t > 0 ? a.HR + (long long)t * a.ldr + kPadL : nullptr, (long long)t * By,
                           a.map + (long long)t * a.ldr + kPadL, t, lane, true);
            continue;
        }
        if (half_map && t > 0) {
            // ticket 1 + 3(b-1) + k: k = 0 fill of band b, k = 1 map of its upper half (input: header row b, from fill b-1),
            // k = 2 map of its lower half (input: the middle row that fill b publishes)
            const int bb = 1 + (t - 1) / 3, k = (t - 1) % 3;
            if (k == 1) {
                map_unit<R / 2, K>(a, warp_smem, sp_tab, a.HR + (long long)bb * a.ldr + kPadL, (long long)bb * By,
                               a.map + (long long)(2 * bb) * a.ldr + kPadL, bb, lane, false);
                continue;
            }
            if (k == 2) {
                map_unit<R / 2, K>(a, warp_smem, sp_tab, a.MID + (long long)bb * a.ldr + kPadL, (long long)bb * By + By / 2,
                               a.map + (long long)(2 * bb + 1) * a.ldr + kPadL, bb, lane, false);
                continue;
            }
            t = bb;
        } else if (with_map && t > 0) {
            if ((t & 1) == 0) {                               // ticket 2k: origin map of band k (its input, header row k, comes from fill unit k-1)
                const int bb = t >> 1;
                map_unit<R, K>(a, warp_smem, sp_tab, a.HR + (long long)bb * a.ldr + kPadL, (long long)bb * By,
                               a.map + (long long)bb * a.ldr + kPadL, bb, lane, false);
                continue;
            }
            t = (t + 1) >> 1;                                 // tickets 0, 1, 3, 5, ... are the fill units of bands 0, 1, 2, 3, ...
        }
        fill_unit<R, K>(a, warp_smem, sp_tab, t, lane, half_map, Handoff{nullptr, nullptr, nullptr, nullptr, nullptr});
        __syncwarp();
    }
}

}  // namespace nwb
