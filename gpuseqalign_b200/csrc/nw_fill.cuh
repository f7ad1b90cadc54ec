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
    const int2* order;       // nullable: ticket t -> (q, b).  Column blocks: units in the order of the wavefront (start time g*U + b*L), so that
                             // the window of tickets in flight follows the diagonal of active units instead of holding whole blocks whose
                             // lower bands cannot start yet (block-major tickets, block 2 048: 1 unit in 8 of the window was runnable)
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

// Shared-memory hand-off of the header row between the warps of one CTA (grouped mode): the top-row ring of the warp below is
// written quad by quad from inside the step loop of the warp above (nw_sweep.cuh, HAND), so a warp follows the one above it at
// the distance of the lane pipeline plus one quad.  The only other coupling is a chunk counter for back-pressure.
struct Handoff {
    unsigned rin_s;            // shared-space address of THIS warp's top-row ring (grouped mode; 0: the plain top-row buffer is used)
    bool self_fed;             // first warp of the CTA: nobody above in this CTA, the warp puts the header row it fetches from HBM into its own ring
    unsigned cons_out_s;       // shared-space address of the count of chunks this warp has completed, read by the warp above
    int* next_rin_g;           // GENERIC pointer (mapa.u64) to the ring of the warp below: the next warp of this CTA, warp 0 of the next
                               // CTA of the thread-block cluster, or a sink ring nobody reads (last band of a cluster unit)
    const int* cons_in_g;      // generic pointer to the chunk counter of the warp below (sink: a word that holds INT_MAX)
};
// All warps of a grouped CTA run the SAME instance of the chunk code (ring in, ring out): three differently specialised copies of
// the unrolled chunk next to the map units' copy overflowed the instruction cache level that SMs share (no_inst 4 % -> 25 % of the
// fill's stall samples as soon as map CTAs ran on neighbouring SMs, ncu r1o).

__device__ __forceinline__ int ld_counter(const int* p)     // a counter in (possibly another CTA's) shared memory
{
    int v;
    asm volatile("ld.relaxed.cluster.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int spin_until_ge(const int* p, int need)
{
    unsigned polls = 0;
    int v;
#pragma unroll 1
    while ((v = ld_counter(p)) < need) {
        if (++polls > (1u << 22)) { g_wait_timeout = 1; break; }
    }
    return v;
}

// Fill of band b of column block q (one warp).  HAND: bit 0 = the top row arrives through the hand-off ring of the warp above,
// bit 1 = the bottom row also goes into the ring of the warp below (grouped mode, see Handoff).
//
// The chunk loop is written with RUNNING values (ring positions, global pointers, the column the prefetches fetch) and with
// shared-space byte addresses: written as base + f(lc) through generic pointers, ptxas rebuilt the 64-bit row products and the
// shared-window bases in every iteration (an S2R and a multiply chain per chunk: 570 clk between two chunks, ncu r1m).
template <int R, int K, int HAND>
__device__ __forceinline__ void fill_unit(const FillArgs& a, unsigned char* warp_smem, const unsigned* sp_tab, const int t, const int lane,
                                          const bool half_map, const Handoff hand)
{
    using SC = Sched<R, K>;
    constexpr int By = SC::By, VR = SC::VR, XR = SC::XR;
    constexpr bool HIN = (HAND & 1) != 0, HOUT = (HAND & 2) != 0;
    static_assert(VR == XR, "the letter ring and the top-row ring advance together");
    WarpSmem<R, K> sm(warp_smem, a.S);
    const int PD = a.pd;
    const unsigned ZOFF = (unsigned)a.S * SC::LSTRIDE;      // profile offset of the all-zero row
    const int nblocks = (a.m + a.wc - 1) / a.wc;             // column blocks of the whole matrix
    constexpr int L4 = SC::L4, HB0 = SC::LAG - SC::L4;     // HB0: ring position of quad 0 of chunk 0
    int q, b;                                             // (q, b-1) and the unit left of (q, b) are always taken before (q, b)
    if (a.order != nullptr) { const int2 o = __ldg(a.order + t); q = o.x; b = o.y; }
    else { q = t / a.nb; b = t - q * a.nb; }              // block-major
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
    if constexpr (!HIN) for (int i = lane; i < VR; i += 32) sm.rin[i] = 0;      // HIN: the CTA set the ring to "empty" before the warp above could write into it

    Lane<R, 0> st;
    st.oprev = 0; st.oup_next = 0; st.o[0] = 0;
    if (has_left) {
        // the column left of this block arrives from the left neighbour (peer stores over NVLink + system-scope flag)
        // (band b-1's flag as well: this lane 0 reads the last row of band b-1 as its diagonal input)
        const unsigned* fl = a.recv_flag + (long long)q * a.nb + b;
        const unsigned long long t0 = a.timeout_ns ? globaltimer_ns() : 0ull;
        while (ld_acquire_sys_u32(fl) != a.xtag || (b > 0 && ld_acquire_sys_u32(fl - 1) != a.xtag)) {
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
    const bool from_hbm = !HIN || hand.self_fed;   // this warp fetches the row above from HR in HBM (or it is row 0 of the matrix)
    const bool consumer = (b > 0) && from_hbm;     // band 0 has row 0 (P = 0) above it
    const unsigned long long* hr_in = a.HR + (long long)q * a.hr_stride + (long long)b * a.ldr + kPadL;
    unsigned long long* hr_out = a.HR + (long long)q * a.hr_stride + (long long)(b + 1) * a.ldr + kPadL;
    const bool with_mid = half_map && b > 0;
    constexpr int GLM = (15 * K + 31) / 32, SHM = 32 * GLM - 15 * K;
    __syncwarp();
    // ---- prologue: the start of the row above
    if (consumer) {
        {   // head start for the producer: the consumer's prefetches then only touch lines that are complete
            int c = 32 * (PD - 1 + a.slack) + 31;
            if (c > m - 1) c = m - 1;
            (void)wait_tagged(hr_in + c, ld_relaxed64(hr_in + c), a.tag);
        }
        for (int g = 0; g < PD; g++) {
            const int c = 32 * g + lane;
            if (c < m) sm.rin[(c + (HIN ? SC::LAG : 0)) & (VR - 1)] = wait_tagged(hr_in + c, ld_relaxed64(hr_in + c), a.tag);     // ring: column c sits at c + LAG
        }
    }
    __syncwarp();
    if constexpr (HIN) {                             // quad 0 of chunk 0: columns -L4 .. 3-L4
        const int4 v = ring_take(hand.rin_s + 4u * (unsigned)(HB0 & (VR - 1)));
        st.cq[0] = v.x; st.cq[1] = v.y; st.cq[2] = v.z; st.cq[3] = v.w;
    } else {
        st.cq[0] = 0; st.cq[1] = 0; st.cq[2] = 0; st.cq[3] = 0;
    }
    __syncwarp();
    if (a.dbg && lane == 0 && q == 0) a.dbg[4 * b + 1] = globaltimer_ns();
    st.up_next = (lane == 0) ? (HIN ? st.cq[L4] : sm.rin[0]) : st.dprev;

    // ---- running values of the chunk loop
    const unsigned xs_s = (unsigned)__cvta_generic_to_shared(sm.xs), rin_s = (unsigned)__cvta_generic_to_shared(sm.rin);
    const unsigned rout_s = (unsigned)__cvta_generic_to_shared(sm.rout), rmid_s = (unsigned)__cvta_generic_to_shared(sm.rmid);
    ChunkIO io;
    io.prof_lane = nullptr; io.xs_lane = nullptr; io.rin_chunk = nullptr; io.rin_next = nullptr; io.rout_chunk = nullptr; io.rmid_chunk = nullptr;
    io.map_out = nullptr; io.org0 = 0; io.dirs_lane = nullptr; io.negg = 0; io.dump_lane = nullptr; io.dump_ld = 0;
    io.prof_s = (unsigned)__cvta_generic_to_shared(sm.prof) + (unsigned)(lane * 4 * SC::WPL);
    io.hin_s = 0; io.hout_p = nullptr; io.hout_on = false;
    io.kconst_s = (unsigned)__cvta_generic_to_shared(sm.kconst);
    if constexpr (HAND != 0) {
        if (lane < 4) sm.kconst[lane] = 1 << (8 * lane);
        if (lane == 4) sm.kconst[4] = -1;
        sm.kconst[8 + lane] = (lane == 31) ? -1 : 0;
        __syncwarp();
    }
    unsigned xpos = (unsigned)(-K * lane) & (XR - 1);      // ring position of this lane's letter of step 0
    unsigned grp = 0;                                      // (32*lc) & (VR-1)
    unsigned tog = 0;                                      // 128 * (lc & 1): which half of the bottom-row / middle-row staging the chunk fills
    int cp = 32 * PD + lane;                               // column whose letter / top-row element this iteration prefetches
    const uint8_t* x_p = xb + cp;
    const unsigned long long* hr_in_p = hr_in + cp;
    unsigned long long* hr_out_p = hr_out + lane - 32 * SC::GL;            // where the group that chunk lc completes goes
    unsigned long long* mid_out_p = a.MID + (long long)b * a.ldr + kPadL + lane - 32 * GLM;
    // a lane's element of a completed group sits in the half the chunk just filled (lanes >= SH) or in the other one
    const unsigned po = (lane >= SC::SH) ? 4u * (unsigned)(lane - SC::SH) : 128u + 4u * (unsigned)(32 - SC::SH + lane);
    const unsigned pm = (lane >= SHM) ? 4u * (unsigned)(lane - SHM) : 128u + 4u * (unsigned)(32 - SHM + lane);
    constexpr int LCF = (HB0 + 3) / 32;                    // the first quad the warp below reads (columns -L4 ..) is the last quad of chunk LCF
    int cons_seen = 0;
    // (a column block of the cross-GPU wavefront that keeps snapshots starts at a multiple of the snapshot spacing: its snapshots carry
    //  the indices they have in the whole matrix)
    int snap_left = a.snap_chunks, snap_k = (int)(c0 / (32LL * a.snap_chunks));
    int lc = 0;
    for (int left = nlc; left > 0; left--, lc++) {        // (a running count: written as lc < nlc, ptxas recomputes nlc from m every iteration)
        // ---- issue the prefetches of chunk lc + PD
        unsigned long long pf_hr = 0;
        const bool want_hr = consumer && cp < m;
        if constexpr (HOUT) {                           // never overwrite a quad the warp below still has to read
            if (cons_seen < lc - 8) cons_seen = spin_until_ge(hand.cons_in_g, lc - 8);
            else cons_seen = ld_counter(hand.cons_in_g);
        }
        if (want_hr) pf_hr = ld_relaxed64(hr_in_p);
        const unsigned pf_x = (cp < m) ? (unsigned)__ldg(x_p) : (unsigned)a.S;      // scaled when it lands: nothing waits on the load here
        // ---- the chunk itself
        io.xs_s = xs_s + 2u * xpos;
        io.rin_s = rin_s + 4u * grp;
        io.rin_next_s = rin_s + 4u * ((grp + 32u) & (VR - 1));
        io.rout_s = rout_s + tog;
        io.rmid_s = with_mid ? rmid_s + tog : 0u;
        // hand-off rings: quads 1..8 of this chunk sit in ONE 32-column group of the ring (HB0 + 4 is a multiple of 32); the warp above
        // writes every quad up to column 32*nlc + 3 (the tail past its last chunk repeats the frozen last value), so there is no
        // end-of-row case in the step loop
        if constexpr (HIN) io.hin_s = hand.rin_s + 4u * ((grp + (unsigned)(HB0 + 4)) & (VR - 1));
        if constexpr (HOUT) { io.hout_p = hand.next_rin_g + grp; io.hout_on = lc > LCF; }
        sweep_chunk<R, K, 0, true, HAND | 4>(st, lane, io, nullptr);
        if constexpr (HIN) { if (lane == 0) sts_volatile1(hand.cons_out_s, lc + 1); }
        if constexpr (HOUT) {
            if (lc == LCF && lane == 31) {
                const int4 v = lds_volatile4(rout_s + tog + 4u * 28u);
                stg_quad(io.hout_p + 28, v.x, v.y, v.z, v.w);
            }
        }
        __syncwarp();
        // ---- publish the group of the bottom row that this chunk completed: ONE coalesced 256-byte store
        //      (and the group of the middle row: lane 15 is 15*K columns behind lane 0)
        int pv = 0, pmv = 0;
        if (lc >= SC::GL) pv = lds_volatile1(rout_s + (po ^ tog));
        if (with_mid && lc >= GLM) pmv = lds_volatile1(rmid_s + (pm ^ tog));
        // ---- land the prefetches
        const unsigned gput = (grp + 32u * (unsigned)PD) & (VR - 1);       // ring position of column cp - lane
        if (from_hbm) {                                  // (a self-fed ring is re-written group by group: reading a group hands its slot back)
            int v = 0;
            if (want_hr) v = (a.dbg_mode == 1) ? (int)(unsigned)pf_hr : wait_tagged_count(hr_in_p, pf_hr, a.tag, spins);
            if (HIN) sts_volatile1(rin_s + 4u * ((gput + (unsigned)(lane + SC::LAG)) & (VR - 1)), v);
            else if (consumer) sts_volatile1(rin_s + 4u * (gput + lane), v);
        }
        {
            const unsigned off16 = pf_x * SC::LSTRIDE;
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(xs_s + 2u * (gput + lane)), "h"((unsigned short)off16));
            if (gput == 0) asm volatile("st.shared.u16 [%0], %1;" ::"r"(xs_s + 2u * (XR + lane)), "h"((unsigned short)off16));   // mirror (XM = 32)
        }
        if (lc >= SC::GL) st_tagged(hr_out_p, pv, a.tag);
        if (with_mid && lc >= GLM) st_tagged(mid_out_p, pmv, a.tag);
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
        xpos = (xpos + 32u) & (XR - 1); grp = (grp + 32u) & (VR - 1); tog ^= 128u;
        cp += 32; x_p += 32; hr_in_p += 32; hr_out_p += 32; mid_out_p += 32;
    }
    // ---- the first SH elements of the next group were produced by the last chunk (they hold the last real column)
    tog ^= 128u;                                     // the half the last chunk filled
    if (lane < SC::SH) st_relaxed64(hr_out_p, pack_tagged(lds_volatile1(rout_s + tog + 4u * (unsigned)(32 - SC::SH + lane)), a.tag));
    if constexpr (HOUT) {
        // the warp below reads the top row up to column 32*nlc + 3: repeat the frozen last value (P[bottom row][m]) over the three
        // groups that follow this warp's last quad
        const int v = __shfl_sync(kFull, st.h[R - 1], 31);
        if (cons_seen < nlc - 5) cons_seen = spin_until_ge(hand.cons_in_g, nlc - 5);
        if (lane < 24) stg_quad(hand.next_rin_g + ((32 * nlc + 4 * lane) & (VR - 1)), v, v, v, v);
    }
    if (with_mid && lane < SHM) st_relaxed64(mid_out_p, pack_tagged(lds_volatile1(rmid_s + tog + 4u * (unsigned)(32 - SHM + lane)), a.tag));
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
    __syncwarp();
}


template <int R, int K, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) nw_fill_kernel(const FillArgs a)
{
    using SC = Sched<R, K>;
    constexpr int By = SC::By, LAG = SC::LAG, VR = SC::VR, XR = SC::XR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned sp_tab[kSpWords];
    stage_sprime(sp_tab, a.sprime, a.S, smem_raw);
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
        // ---- grouped mode: CLUSTER-level tickets (a cluster of CL CTAs; CL = 1 without a cluster launch).  Fill unit k = bands
        // k*CL*WARPS .. (k+1)*CL*WARPS-1: CTA r of the cluster takes the four bands (k*CL + r)*WARPS + w, warp w feeds the warp below
        // through its ring in shared memory, the last warp of CTA r feeds warp 0 of CTA r+1 through distributed shared memory
        // (the same ring, st.shared::cluster).  Map unit j = CL*WARPS map units (one warp each, inputs from HBM).
        // Order: F0, then for k >= 1: F(k) followed by the map units of fill unit k-1, finally the map units of the last fill unit.
        __shared__ int s_ticket;
        __shared__ int s_cons[WARPS + 1];       // [WARPS]: INT_MAX, the "consumer" of a warp that feeds nobody
        const unsigned CL = cluster_size(), crank = cluster_rank();
        const int per_unit = (int)CL * WARPS;
        const int nG = (a.nb + per_unit - 1) / per_unit;
        const int mper = with_map ? (half_map ? 2 : 1) : 0;           // map units per fill unit
        const int ncl_units = nG * (1 + mper);
        for (;;) {
            if (CL > 1) cluster_sync_all(); else __syncthreads();      // everybody is done with the previous unit (and with my shared memory)
            if (crank == 0 && threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1);
            if (threadIdx.x <= WARPS) s_cons[threadIdx.x] = (threadIdx.x < WARPS) ? 0 : 0x7fffffff;
            if (CL > 1) cluster_sync_all(); else __syncthreads();
            const int T = (crank == 0) ? s_ticket : ldc_volatile1(map_to_cta((unsigned)__cvta_generic_to_shared(&s_ticket), 0));
            if (T >= ncl_units) {
                if (CL > 1) cluster_sync_all();      // CTA 0 must not exit while the others still read its ticket
                break;
            }
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
            // a ring that another warp feeds starts out empty (-1 in every slot); every other top-row ring is plain data
            for (int i = lane; i < VR; i += 32) sm.rin[i] = (fill_k >= 0 && (w > 0 || crank > 0)) ? -1 : 0;
            if (CL > 1) cluster_sync_all(); else __syncthreads();
            if (fill_k >= 0) {
                const int b = (fill_k * (int)CL + (int)crank) * WARPS + w;
                if (b < a.nb) {
                    Handoff hand;
                    hand.rin_s = (unsigned)__cvta_generic_to_shared(sm.rin);
                    hand.self_fed = (w == 0 && crank == 0);
                    hand.cons_out_s = (unsigned)__cvta_generic_to_shared(&s_cons[w]);
                    const bool last_of_unit = (w + 1 == WARPS) && (crank + 1 == CL);
                    const bool below = !last_of_unit && (b + 1 < a.nb);
                    if (!below) {
                        hand.next_rin_g = reinterpret_cast<int*>(smem_raw + (size_t)WARPS * SC::warp_smem_bytes(a.S));      // the sink
                        hand.cons_in_g = &s_cons[WARPS];
                    } else if (w + 1 < WARPS) {
                        hand.next_rin_g = WarpSmem<R, K>(warp_smem + SC::warp_smem_bytes(a.S), a.S).rin;
                        hand.cons_in_g = &s_cons[w + 1];
                    } else {                            // warp 0 of the next CTA of the cluster (same layout: my own warp 0's addresses, mapped)
                        hand.next_rin_g = map_generic_to_cta(WarpSmem<R, K>(smem_raw, a.S).rin, crank + 1);
                        hand.cons_in_g = map_generic_to_cta(&s_cons[0], crank + 1);
                    }
                    fill_unit<R, K, 3>(a, warp_smem, sp_tab, b, lane, half_map, hand);
                }
            } else {
                const int u = (map_j * (int)CL + (int)crank) * WARPS + w;   // map unit: band u / 2, half u % 2 (half maps) or band u
                const int bb = half_map ? (u >> 1) : u;
                if (bb >= 1 && bb < a.nb) {
                    if (a.dbg && lane == 0) a.dbg[4 * (a.nb + u) + 0] = globaltimer_ns();
                    if (half_map) {
                        const int hh = u & 1;
                        map_unit<R / 2, K>(a, warp_smem, sp_tab, (hh ? a.MID : a.HR) + (long long)bb * a.ldr + kPadL, (long long)bb * By + hh * (By / 2),
                                           a.map + (long long)u * a.ldr + kPadL, bb, lane, false);
                    } else {
                        map_unit<R, K>(a, warp_smem, sp_tab, a.HR + (long long)bb * a.ldr + kPadL, (long long)bb * By,
                                       a.map + (long long)bb * a.ldr + kPadL, bb, lane, false);
                    }
                    if (a.dbg && lane == 0) a.dbg[4 * (a.nb + u) + 2] = globaltimer_ns();
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
        fill_unit<R, K, 0>(a, warp_smem, sp_tab, t, lane, half_map, Handoff{0u, false, 0u, nullptr, nullptr});
        __syncwarp();
    }
}

}  // namespace nwb
