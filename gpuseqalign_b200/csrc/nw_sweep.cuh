// nw_sweep.cuh -- the band sweep: the one device routine every kernel of the engine is built on.
//
// It replaces the tile computation of the reference (Nw_Gpu9_KernelB, nwalign_gpu9_mlsp_diagdiagdiag.cu:69-360,
// recurrence at :292-294) and its host-side twin used by the traceback (NwTrace2_AlignTile,
// nwtrace2_sparse.cpp:40-96).  B200-first design, not a translation:
//
//   * The matrix is cut into horizontal BANDS of By = 32*R rows.  One WARP sweeps one band left to right
//     over all columns in a single pass: lane l owns rows l*R .. l*R+R-1 of the band and is at column
//     c = s - K*l at step s, so the warp is a 32-stage systolic array whose lanes are K columns apart.
//     All state lives in registers (R "left" values + one diagonal value per lane); the value crossing a lane
//     boundary moves with ONE rotate-shuffle per step (K = 2: issued one step early, off the critical path).
//   * Arithmetic is done in shifted coordinates P[i][j] = H[i][j] - (i+j)*gap (the
//     reference's half-way form is nwalign_gpu1_ml_diag.cu:65-70):
//         P[i][j] = max3(P[i-1][j-1] + s', P[i-1][j], P[i][j-1]),   s' = max(subst - 2*gap, 0)  (a byte)
//     One cell = one IDP.4A (adds byte r of a packed per-lane profile word to the diagonal value) + one
//     VIMNMX3 (DPX 3-way max): two issue slots on two pipes.  Row 0 and column 0 of P are all zero, so
//     there is no header-init kernel (reference Nw_Gpu9_KernelA).  P is monotone along rows and columns;
//     therefore a column whose profile row is all zero (columns outside the sequence) FREEZES every row at
//     its last value and a row whose profile bytes are zero fed with zeros stays zero: padding needs no
//     bounds checks anywhere in the inner loop.  Rows are aligned to the BOTTOM of the matrix (the padding
//     rows sit above row 1 of band 0), so the last band's bottom row is row n and the score is the last
//     value lane 31 holds.
//   * MODE selects what a cell leaves behind:
//       0  score       nothing (the band's bottom row streams out as the next band's top row)
//       1  origin      every cell carries the column at which its traceback path leaves the band through the
//                      top row; the bottom row's labels form the band's entry->exit map
//       2  directions  a 2-bit move code per cell (0 '=', 1 'X', 2 'I' up, 3 'D' left) written to shared
//                      memory for the walker.  The choice follows nwtrace1_plain.cpp:29-100 /
//                      nwtrace2_sparse.cpp:149-180 exactly: compare the NEIGHBOUR SCORES diag, up, left with
//                      strict '<', preference diag > up > left.  In shifted coordinates H_diag vs H_up vs
//                      H_left is (P_diag - gap) vs P_up vs P_left.
//       3  dump        every cell value P is written to a row-major slab in HBM (score hash, header export)
#pragma once
#include "nw_common.cuh"

namespace nwb {

constexpr int kPadL = 64;     // elements of padding in front of every header row (columns -64..-1 are scratch)

// Per-warp block of "constants" in shared memory (fill, grouped mode): the four byte selectors of IDP.4A, -1, and a per-lane mask
// that is -1 in lane 31.  They are LOADED at the top of every chunk: behind each wait loop of the unrolled chunk ptxas otherwise
// re-materialises them (4 IMAD.MOV, and S2R + LOP3 + ISETP for the lane-31 predicate, per loop), which a lone warp pays in full.
constexpr int kConstInts = 48;      // [0..3] selectors, [4] -1, [8 + lane] last-lane mask

template <int R, int K>
struct Sched {
    static_assert(R == 2 || R == 4 || R == 8 || R == 16, "rows per lane");
    static_assert(K == 1 || K == 2, "lane skew");
    static constexpr int WPL = (R + 3) / 4;            // profile words per lane and letter (R == 2 uses half a word)
    static constexpr int By = 32 * R;                  // rows per band
    static constexpr int LAG = 31 * K;                 // columns lane 31 is behind lane 0
    static constexpr int PD = 2;                       // top-row / letter groups prefetched ahead
    static constexpr int VR = 256;                     // ints in the top-row ring (8 groups of 32 columns)
    static constexpr int XR = 256;                     // entries in the letter ring (8 groups)
    static constexpr int XM = 32;                      // mirror entries behind the letter ring
    static constexpr int LSTRIDE = 128 * WPL;          // bytes between the profile rows of two letters
    static constexpr int SNAP_INTS = R + 4;            // per-lane snapshot: h[R], dprev, up_next, 2 pad
    static constexpr int GL = (LAG + 31) / 32;         // chunks until a 32-column group of the bottom row is complete
    static constexpr int SH = 32 * GL - LAG;           // element of a group that the first step of a chunk produces
    static constexpr int L4 = LAG & 3;                 // hand-off ring (grouped fill): column c sits at position (c + LAG) & (VR-1), so that the
                                                       // four values lane 31 produces in steps 4k..4k+3 form one aligned quad
    static_assert(kPadL >= LAG + 1, "header row padding");
    __host__ __device__ static constexpr int nlc(int m) { return (m + LAG + 31) / 32; }      // 32-step chunks per band
    __host__ __device__ static constexpr size_t prof_bytes(int S) { return (size_t)(S + 1) * LSTRIDE; }
    __host__ __device__ static constexpr size_t warp_smem_bytes(int S)
    {
        return (prof_bytes(S) + (size_t)VR * 4 + 128 * 4 + kConstInts * 4 + (size_t)(XR + XM) * 2 + 15) & ~(size_t)15;
    }
};

// Register state of one lane.
template <int R, int MODE>
struct Lane {
    int h[R];        // h[r] = P[row r][c-1]: the left neighbour of the cell row r computes next
    int dprev;       // P[row0-1][c-1]: the diagonal neighbour of row 0 (= the previous step's `up`)
    int up_next;     // K == 2: the shuffled upper neighbour for the next step
    int o[MODE == 1 ? R : 1];   // MODE 1: origin labels travelling with h / dprev / up_next
    int oprev;
    int oup_next;
    int cq[4];       // grouped fill: the last quad of the top row read from the hand-off ring (its tail belongs to the next chunk)
};

// Warp-private shared memory.
template <int R, int K>
struct WarpSmem {
    using SC = Sched<R, K>;
    unsigned char* prof;     // [(S+1)][32 lanes][WPL] words: bytes s'(y[row r], letter)
    int* rin;                // [VR] top-row ring: P[top][c] at (c & (VR-1))
    int* rout;               // [64] bottom-row staging: chunk lc, step s at ((lc & 1) * 32 + s)
    int* rmid;               // [64] the same for the band's MIDDLE row (bottom row of lane 15), see nw_fill.cuh
    int* kconst;             // [kConstInts] see above
    unsigned short* xs;      // [XR + XM] letter ring: profile byte offset (letter * LSTRIDE) of column c at (c & (XR-1))
    __device__ __forceinline__ WarpSmem(unsigned char* base, int S)
    {
        prof = base;
        rin = reinterpret_cast<int*>(base + SC::prof_bytes(S));
        rout = rin + SC::VR;
        rmid = rout + 64;
        kconst = rmid + 64;
        xs = reinterpret_cast<unsigned short*>(kconst + kConstInts);
    }
    __device__ __forceinline__ void put_letter(int c, unsigned off16)
    {
        const int p = c & (SC::XR - 1);
        xs[p] = (unsigned short)off16;
        if (p < SC::XM) xs[p + SC::XR] = (unsigned short)off16;
    }
};

// CTA-wide copy of the byte table s'[y][x] in shared memory: rows of 17 words (68 bytes: conflict-free when the lanes
// of a warp read the same column group of 32 different rows), row S is all zero (padding rows).
constexpr int kSpPitch = 17;                              // words per row
constexpr int kSpWords = (kMaxLetters + 1) * kSpPitch;    // static shared memory of every kernel that builds profiles
// The S*S byte table arrives with ONE bulk copy of the TMA engine (cp.async.bulk global -> shared, completion on an mbarrier)
// into `landing` -- any 16-byte aligned shared memory of >= S*S + 15 bytes that the caller does not use yet (every kernel passes
// the start of its dynamic shared memory) -- and is then spread into the padded rows.  (The letters themselves are consumed 32
// bytes per chunk and warp, at arbitrary alignment: below the granularity of a bulk copy; they stay on the LDG path.)
// `pitch` (words per row of sp_tab) defaults to kSpPitch; the packed batch kernel asks for 16-byte aligned rows.
__device__ __forceinline__ void stage_sprime(unsigned* sp_tab, const uint8_t* __restrict__ sprime, int S, unsigned char* landing, int words = kSpWords,
                                             int pitch = kSpPitch)
{
    __shared__ __align__(8) unsigned long long s_mbar;
    const unsigned bytes = ((unsigned)(S * S) + 15u) & ~15u;           // the device buffer is allocated with slack (nwb200_set_scoring)
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(&s_mbar), dst = (unsigned)__cvta_generic_to_shared(landing);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int i = threadIdx.x; i < words; i += blockDim.x) sp_tab[i] = 0u;        // `words` < kSpWords: a table for a smaller alphabet
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(sprime), "r"(bytes), "r"(mbar) : "memory");
    }
    unsigned done = 0, polls = 0;
    while (!done) {                                                      // phase 0 of the one-shot barrier
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(mbar) : "memory");
        if (++polls > (1u << 22)) { g_wait_timeout = 1; break; }
    }
    unsigned char* t = reinterpret_cast<unsigned char*>(sp_tab);
    for (int i = threadIdx.x; i < S * S; i += blockDim.x) t[(i / S) * (pitch * 4) + (i % S)] = landing[i];
    __syncthreads();
}

// Per-lane packed profile for the band's rows: prof[letter][lane][q] = bytes s'(y[row 4q+0..3], letter); row S of prof is
// the all-zero row used outside the sequence.  Four table rows (one per matrix row) are read a word (4 letters) at a
// time and transposed with byte permutes: 16 instructions per 4 letters and 4 rows.
// yrow0 = 0-based index into y of this lane's first row (negative / >= n: padding row, zero bytes).
// SHL: bit q set = the bytes of word q are stored doubled (the packed batch kernel keeps 2*s' for its second pair; s' <= 127 there).
template <int R, int K, int SHL = 0>
__device__ __forceinline__ void build_profile(const WarpSmem<R, K>& sm, const unsigned* __restrict__ sp_tab, int S,
                                              const uint8_t* __restrict__ y, long long yrow0, long long n, int lane, unsigned* yl_out)
{
    using SC = Sched<R, K>;
    constexpr int WPL = SC::WPL;
    unsigned yl[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const long long i = yrow0 + r;
        yl[r] = (i >= 0 && i < n) ? (unsigned)__ldg(y + i) : 0xffu;
        if (yl_out) yl_out[r] = yl[r];
    }
    unsigned* pl = reinterpret_cast<unsigned*>(sm.prof) + lane * WPL;
#pragma unroll
    for (int q = 0; q < WPL; q++) {
        const unsigned* row[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const unsigned yy = (q * 4 + r < R) ? yl[(q * 4 + r < R) ? q * 4 + r : 0] : 0xffu;
            row[r] = sp_tab + (yy != 0xffu && yy < (unsigned)S ? yy : (unsigned)S) * kSpPitch;
        }
        for (int g4 = 0; 4 * g4 < S; g4++) {
            const unsigned w0 = row[0][g4], w1 = row[1][g4], w2 = row[2][g4], w3 = row[3][g4];
            const unsigned t0 = __byte_perm(w0, w1, 0x5140), t1 = __byte_perm(w2, w3, 0x5140);
            const unsigned t2 = __byte_perm(w0, w1, 0x7362), t3 = __byte_perm(w2, w3, 0x7362);
            const unsigned o0 = __byte_perm(t0, t1, 0x5410), o1 = __byte_perm(t0, t1, 0x7632);
            const unsigned o2 = __byte_perm(t2, t3, 0x5410), o3 = __byte_perm(t2, t3, 0x7632);
            const int xl = 4 * g4;
            pl[(size_t)(xl + 0) * 32 * WPL + q] = o0 << ((SHL >> q) & 1);
            if (xl + 1 < S) pl[(size_t)(xl + 1) * 32 * WPL + q] = o1 << ((SHL >> q) & 1);
            if (xl + 2 < S) pl[(size_t)(xl + 2) * 32 * WPL + q] = o2 << ((SHL >> q) & 1);
            if (xl + 3 < S) pl[(size_t)(xl + 3) * 32 * WPL + q] = o3 << ((SHL >> q) & 1);
        }
        pl[(size_t)S * 32 * WPL + q] = 0u;
    }
}

// What one chunk reads and writes besides the lane state.
struct ChunkIO {
    const unsigned short* xs_lane;   // &xs[(32*lc - K*lane) & (XR-1)]: this lane's letter offsets, index s
    const unsigned char* prof_lane;  // &prof[0][lane][0] as bytes
    const int* rin_chunk;            // &rin[(32*lc) & (VR-1)]: top row for lane 0, index s
    const int* rin_next;             // &rin[(32*lc + 32) & (VR-1)]: first element of the next group (K == 2 look-ahead)
    int* rout_chunk;                 // MODE 0: &rout[(lc & 1) * 32] (lane 31 stores element s), nullptr = keep nothing
    int* rmid_chunk;                 // MODE 0: &rmid[(lc & 1) * 32] (lane 15 stores element s), nullptr = keep nothing
    int* map_out;                    // MODE 1: &map[32*lc - LAG] (lane 31 stores element s)
    int org0;                        // MODE 1: label of the cell above lane 0 at step 0 of this chunk (= 32*lc + 1)
    unsigned char* dirs_lane;        // MODE 2: &dirs[(32*(lc-lc0))*32 + lane], one byte (R=4) / two (R=8) / four (R=16) per step
    int negg;                        // -gap
    int* dump_lane;                  // MODE 3: &dump[(lane*R)*dump_ld + kPadL + 32*lc - K*lane]: cell (row r, step s) at [r*dump_ld + s]
    long long dump_ld;
    // ---- grouped fill (HAND != 0): the header row crosses the warps of a CTA through a ring of quads in shared memory.
    // A quad is written by ONE 16-byte store and is its own ready flag: P >= 0 everywhere, the last word of an empty slot is -1.
    unsigned hin_s;                  // HAND & 1: shared-space byte address of quad 1 of this chunk in this warp's ring (quad j at + 16*(j-1))
    int* hout_p;                     // HAND & 2: generic pointer (possibly into another CTA of the cluster) to the ring of the warp below at position (32*lc) & (VR-1)
    bool hout_on;                    // HAND & 2: this chunk's quads are read by the warp below (false in the first chunk(s): columns < 0)
    // ---- HAND & 4: the fill's chunk loop hands over shared-space byte addresses (it keeps them as running values: a chunk loop
    // written with pointers into the shared window makes ptxas rebuild window bases and 64-bit products in every iteration)
    unsigned xs_s, prof_s;           // = xs_lane, prof_lane
    unsigned rin_s, rin_next_s;      // = rin_chunk, rin_next
    unsigned rout_s, rmid_s;         // = rout_chunk, rmid_chunk (0 = keep nothing)
    unsigned kconst_s;               // the warp's constant block (HAND & 3 only)
};

__device__ __forceinline__ int4 lds_volatile4(unsigned addr)
{
    int4 v;
    asm volatile("ld.volatile.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_volatile4(unsigned addr, int a, int b, int c, int d)
{
    asm volatile("st.volatile.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d));
}
// the same through the cluster window: `addr` comes from mapa (the ring / counter of a warp in ANOTHER CTA of the thread-block
// cluster, or in this one -- a CTA's own shared memory is part of the window)
// A quad into a ring that may live in ANOTHER CTA of the thread-block cluster: a plain store through a GENERIC pointer (mapa.u64).
// (st.shared::cluster with a 32-bit window address is lowered to the same generic store, but ptxas then rebuilds the 64-bit
// address -- S2R SR_SWINHI + two moves -- in front of every single store.)
__device__ __forceinline__ void stg_quad(int* p, int a, int b, int c, int d)
{
    asm volatile("st.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ int* map_generic_to_cta(int* p, unsigned cta_rank)
{
    unsigned long long r;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(r) : "l"((unsigned long long)p), "r"(cta_rank));
    return reinterpret_cast<int*>(r);
}
__device__ __forceinline__ int ldc_volatile1(unsigned addr)
{
    int v;
    asm volatile("ld.volatile.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned map_to_cta(unsigned addr, unsigned cta_rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ unsigned cluster_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_size()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void sts_volatile1(unsigned addr, int a)
{
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(addr), "r"(a));
}
__device__ __forceinline__ int lds_volatile1(unsigned addr)
{
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// Slow path of a ring read: the quad at `addr` was still empty.  Polls (warp-uniform, bounded like every wait of the engine).
__device__ __forceinline__ int4 ring_wait(unsigned addr)
{
    int4 v;
    unsigned polls = 0;
#pragma unroll 1
    do {
        v = lds_volatile4(addr);
        if (++polls > (1u << 22)) { g_wait_timeout = 1; v = make_int4(0, 0, 0, 0); }
    } while (__any_sync(kFull, (v.x | v.y | v.z | v.w) < 0));
    return v;
}
// A quad is its own ready flag: an EMPTY slot holds -1 in ALL FOUR words and P >= 0 everywhere, so a quad is complete exactly when
// none of its words is negative.  (The producer's 16-byte store -- possibly into another CTA's shared memory -- is not promised to
// be single-copy atomic: checking one word only would let a torn store pass with stale words of the previous lap.)
// complete read of one quad: wait for the producer, then hand the slot back as empty
__device__ __forceinline__ int4 ring_take(unsigned addr)
{
    int4 v = lds_volatile4(addr);
    if (__any_sync(kFull, (v.x | v.y | v.z | v.w) < 0)) v = ring_wait(addr);
    sts_volatile4(addr, -1, -1, -1, -1);
    return v;
}
constexpr int kRingQG = 2;      // quads per readiness check of a hand-off ring

// One 32-step chunk of one warp.
template <int R, int K, int MODE, bool TOP = true, int HAND = 0>
__device__ __forceinline__ void sweep_chunk(Lane<R, MODE>& st, const int lane, const ChunkIO& io, const unsigned* __restrict__ yoff)
{
    using SC = Sched<R, K>;
    constexpr int WPL = SC::WPL;
    const int src_lane = (lane + 31) & 31;
    const bool last = (lane == 31);
    // grouped fill: selectors, -1 and the last-lane mask come from shared memory (see kConstInts)
    constexpr bool KC = (HAND & 3) != 0;
    int4 selq = make_int4(1, 1 << 8, 1 << 16, 1 << 24);
    int minus1 = -1, lastmask = 0;
    if constexpr (KC) {
        selq = lds_volatile4(io.kconst_s);
        minus1 = lds_volatile1(io.kconst_s + 16u);
        lastmask = lds_volatile1(io.kconst_s + 32u + 4u * (unsigned)lane);
    }
    // A single warp can issue a shared-memory / shuffle instruction only every ~5-6 clk (measured: LDS.32 20 thread-ops/clk/SM
    // at one warp per SM sub-partition, profiles/microbench_r1.jsonl), and that -- not the DPX chain -- bounds the step of a
    // lone warp.  So the loads are widened: letter offsets two steps per LDS.32 (K == 2: the lane's ring position is even),
    // the top row four steps per LDS.128, and lane 31's bottom row / map row leave four (two) steps per store.
    // software pipeline: letter offsets >= 3 steps ahead, profile words 2 steps ahead, top row >= 2 steps ahead
    unsigned xo[32 + 4];
    unsigned pw[32 + 2][WPL];
    int rv[32 + 8];
    int outv[32];      // lane 31's bottom-row values of this chunk (registers; stored four at a time)
    int outo[32];      // MODE 1: their origin labels
    constexpr bool SA = (HAND & 4) != 0;
    static_assert(HAND == 0 || SA, "the hand-off rings are used by the fill, which passes shared-space addresses");
    auto load_pw = [&](int s) {
        if constexpr (SA) {
            const unsigned ad = io.prof_s + xo[s];
            if constexpr (WPL == 1) asm("ld.shared.u32 %0, [%1];" : "=r"(pw[s][0]) : "r"(ad));
            else if constexpr (WPL == 2) asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(pw[s][0]), "=r"(pw[s][1]) : "r"(ad));
            else asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(pw[s][0]), "=r"(pw[s][1]), "=r"(pw[s][2]), "=r"(pw[s][3]) : "r"(ad));
            return;
        }
        const unsigned char* pp = io.prof_lane + xo[s];
        if constexpr (WPL == 1) pw[s][0] = *reinterpret_cast<const unsigned*>(pp);
        else if constexpr (WPL == 2) { uint2 v = *reinterpret_cast<const uint2*>(pp); pw[s][0] = v.x; pw[s][1] = v.y; }
        else { uint4 v = *reinterpret_cast<const uint4*>(pp); pw[s][0] = v.x; pw[s][1] = v.y; pw[s][2] = v.z; pw[s][3] = v.w; }
    };
    auto load_xo = [&](int s) {      // fills xo[s] (and xo[s+1] when the ring position allows a paired load)
        if constexpr (K == 2) {
            if ((s & 1) == 0 && s < 32) {
                unsigned v;
                if constexpr (SA) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(io.xs_s + 2u * s));
                else v = *reinterpret_cast<const unsigned*>(io.xs_lane + s);
                xo[s] = v & 0xffffu; xo[s + 1] = v >> 16;
            }
        } else {
            if (s < 32) {
                if constexpr (SA) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(xo[s]) : "r"(io.xs_s + 2u * s));
                else xo[s] = io.xs_lane[s];
            }
        }
    };
    // rvq[s]: top-row element s of this chunk's group (element 32 = first element of the next group)
    auto load_rv4 = [&](int q) {     // elements 4q .. 4q+3
        if constexpr (TOP) {
            if (q < 8) {
                int4 v;
                if constexpr (SA) v = lds_volatile4(io.rin_s + 16u * q);
                else v = *reinterpret_cast<const int4*>(io.rin_chunk + 4 * q);
                rv[4 * q] = v.x; rv[4 * q + 1] = v.y; rv[4 * q + 2] = v.z; rv[4 * q + 3] = v.w;
            } else {
                if constexpr (SA) rv[32] = lds_volatile1(io.rin_next_s);
                else rv[32] = io.rin_next[0];
            }
        }
    };
    constexpr int A = (K == 2) ? 1 : 0;          // the shuffle issued at step s carries top-row element s + A
    constexpr bool HIN = (HAND & 1) != 0, HOUT = (HAND & 2) != 0;
    static_assert(!HIN || TOP, "a hand-off ring feeds the top row");
    // HIN: quad j of the chunk = top-row elements 4j-L4 .. 4j-L4+3, first needed at step 4j-3 (both skews).  Quads are taken in
    // groups of QG: the loads are issued one step before the first quad is needed and land (wait for the producer while a quad
    // of the group is still empty, then hand the slots back) right before use: the warp follows the warp above at a distance of
    // one group.  Quad 0 is the previous chunk's quad 8, carried in registers.  (A check per quad costs more than it saves: every
    // wait loop inside the unrolled chunk makes ptxas rematerialise addresses and constants behind it.)
    constexpr int L4 = SC::L4, QG = kRingQG;
    int4 qq[QG];
    if constexpr (HIN) {
#pragma unroll
        for (int e = 0; e < 4 - L4; e++) rv[e] = st.cq[e + L4];
    }
#pragma unroll
    for (int s = 0; s < 4; s++) load_xo(s);
    if constexpr (!HIN) { load_rv4(0); load_rv4(1); }
#pragma unroll
    for (int s = 0; s < 2; s++) load_pw(s);
#pragma unroll
    for (int s = 0; s < 32; s++) {
        if constexpr (K == 2) { if ((s & 1) == 0) load_xo(s + 4); } else { load_xo(s + 3); }
        if constexpr (HIN) {
            if ((s & (4 * QG - 1)) == 0) {
#pragma unroll
                for (int q = 0; q < QG; q++) qq[q] = lds_volatile4(io.hin_s + 16u * (s / 4 + q));
            } else if ((s & (4 * QG - 1)) == 1) {
                const int j0 = (s + 3) / 4;                        // first quad of the group
                const unsigned a0 = io.hin_s + 16u * (j0 - 1);
                // EVERY quad of the group is checked and handed back (not only the last one): stores into another CTA's shared
                // memory are not promised to arrive in program order
                // ... and ALL FOUR words of a quad: a torn 16-byte store must not pass with stale words (see ring_take)
                int any_w = qq[0].x | qq[0].y | qq[0].z | qq[0].w;
#pragma unroll
                for (int q = 1; q < QG; q++) any_w |= qq[q].x | qq[q].y | qq[q].z | qq[q].w;
                if (__any_sync(kFull, any_w < 0)) {
#pragma unroll
                    for (int q = 0; q < QG; q++) qq[q] = ring_wait(a0 + 16u * q);
                }
#pragma unroll
                for (int q = 0; q < QG; q++) sts_volatile4(a0 + 16u * q, minus1, minus1, minus1, minus1);
#pragma unroll
                for (int q = 0; q < QG; q++) {
                    const int e0 = 4 * (j0 + q) - L4;
                    rv[e0] = qq[q].x; rv[e0 + 1] = qq[q].y; rv[e0 + 2] = qq[q].z; rv[e0 + 3] = qq[q].w;
                }
                if (j0 + QG - 1 == 8) { st.cq[0] = qq[QG - 1].x; st.cq[1] = qq[QG - 1].y; st.cq[2] = qq[QG - 1].z; st.cq[3] = qq[QG - 1].w; }
            }
        } else {
            if ((s & 3) == 0) load_rv4(s / 4 + 2);
        }
        if (s + 2 < 32) load_pw(s + 2);
        const int rvs = TOP ? rv[s + A] : 0;
        // Rotate-shuffle: lanes 0..30 hand their bottom row to the lane below; lane 31 hands lane 0 its next
        // input from the row above the band, so the result feeds the first VIMNMX3 directly.
        int up, oup = 0;
        if constexpr (K == 1) {
            up = __shfl_sync(kFull, last ? rvs : st.h[R - 1], src_lane);
            if constexpr (MODE == 1) oup = __shfl_sync(kFull, last ? io.org0 + s : st.o[R - 1], src_lane);
        } else {
            up = st.up_next;
            if constexpr (KC) st.up_next = __shfl_sync(kFull, (st.h[R - 1] & ~lastmask) | (rvs & lastmask), src_lane);
            else st.up_next = __shfl_sync(kFull, last ? rvs : st.h[R - 1], src_lane);     // consumed at step s+1
            if constexpr (MODE == 1) {
                oup = st.oup_next;
                st.oup_next = __shfl_sync(kFull, last ? io.org0 + s + 1 : st.o[R - 1], src_lane);
            }
        }
        int diag = st.dprev, odiag = st.oprev;
        st.dprev = up;
        if constexpr (MODE == 1) st.oprev = oup;
        unsigned codes = 0;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int left = st.h[r];
            const unsigned selr = (r & 3) == 0 ? (unsigned)selq.x : (r & 3) == 1 ? (unsigned)selq.y : (r & 3) == 2 ? (unsigned)selq.z : (unsigned)selq.w;
            const int t = add_byte(pw[s][r >> 2], KC ? selr : (1u << (8 * (r & 3))), diag);
            const int nv = max3(t, up, left);
            if constexpr (MODE == 3) io.dump_lane[(long long)r * io.dump_ld + s] = nv;
            if constexpr (MODE == 1 || MODE == 2) {
                const int cd = diag + io.negg;            // H_diag - H_up/H_left offset: P_diag - gap
                const bool p1 = cd < up;                  // nwtrace1_plain.cpp:57: max < up  -> move up
                const int b1 = max(cd, up);
                const bool p2 = b1 < left;                // nwtrace1_plain.cpp:65: max < left -> move left
                if constexpr (MODE == 1) {
                    // the label coming down the column (oup) is the late input: everything else is selected first, so only
                    // one SEL per row sits on the chain that runs down the lane's rows
                    const int oleft = st.o[r];
                    const int early = p2 ? oleft : odiag;
                    const bool take_up = p1 && !p2;
                    const int on = take_up ? oup : early;
                    odiag = oleft; oup = on; st.o[r] = on;
                } else {
                    const unsigned eqx = (xo[s] == yoff[r]) ? 0u : 1u;
                    const unsigned code = p2 ? 3u : (p1 ? 2u : eqx);
                    codes |= code << (2 * r);
                }
            }
            diag = left; up = nv; st.h[r] = nv;
        }
        if constexpr (MODE == 0 || MODE == 1) {
            outv[s] = st.h[R - 1];
            if constexpr (SA) {
                if ((s & 3) == 3 && io.rout_s != 0 && last) sts_volatile4(io.rout_s + 4u * (s - 3), outv[s - 3], outv[s - 2], outv[s - 1], outv[s]);
                if ((s & 3) == 3 && io.rmid_s != 0 && lane == 15) sts_volatile4(io.rmid_s + 4u * (s - 3), outv[s - 3], outv[s - 2], outv[s - 1], outv[s]);
            } else {
                if ((s & 3) == 3 && io.rout_chunk != nullptr && last)
                    *reinterpret_cast<int4*>(io.rout_chunk + s - 3) = make_int4(outv[s - 3], outv[s - 2], outv[s - 1], outv[s]);
                if constexpr (MODE == 0) {
                    if ((s & 3) == 3 && io.rmid_chunk != nullptr && lane == 15)
                        *reinterpret_cast<int4*>(io.rmid_chunk + s - 3) = make_int4(outv[s - 3], outv[s - 2], outv[s - 1], outv[s]);
                }
            }
            if constexpr (HOUT) {      // ... and straight into the ring of the warp below: one quad per four steps
                if ((s & 3) == 3 && last && io.hout_on) stg_quad(io.hout_p + (s - 3), outv[s - 3], outv[s - 2], outv[s - 1], outv[s]);
            }
        }
        if constexpr (MODE == 1) {
            outo[s] = st.o[R - 1];
            // map_out + s is 8-byte aligned for even s (the row starts at an even element: kPadL, 32*lc and LAG are even for K == 2)
            if constexpr (K == 2) { if ((s & 1) == 1 && last) *reinterpret_cast<int2*>(io.map_out + s - 1) = make_int2(outo[s - 1], outo[s]); }
            else { if (last) io.map_out[s] = outo[s]; }
        } else if constexpr (MODE == 2) {
            if constexpr (R == 4) io.dirs_lane[s * 32] = (unsigned char)codes;
            else if constexpr (R == 8) reinterpret_cast<unsigned short*>(io.dirs_lane)[s * 32] = (unsigned short)codes;
            else reinterpret_cast<unsigned*>(io.dirs_lane)[s * 32] = codes;
        }
    }
}

}  // namespace nwb
