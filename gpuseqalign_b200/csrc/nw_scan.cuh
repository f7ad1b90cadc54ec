// nw_scan.cuh -- score of a matrix with FEW ROWS and VERY MANY COLUMNS (BASELINE config 4: 2 048 x 4 194 304) by the
// row-parallel prefix-max formulation (the reference has no counterpart -- its tile wavefront has at
// most min(trows, tcols) = 16 blocks in flight for this shape, nwalign_gpu9_mlsp_diagdiagdiag.cu:589).
//
// In shifted coordinates (P = H - (i+j)*gap, nw_sweep.cuh) a whole row is an elementwise step and a prefix maximum:
//     A[j]    = max(P[i-1][j-1] + s'(y_i, x_j), P[i-1][j])              (no dependence inside the row)
//     P[i][j] = max(A[1..j])                                            (max is associative and exact on integers)
// so the dependence along the long dimension disappears: the columns are cut into chunks of kScanT*kScanC columns, one
// CTA owns a chunk for ALL rows and keeps the previous row in registers (kScanC columns per thread).  Per row a thread
// runs the usual IDP.4A + VIMNMX3 cell over its columns (local prefix), the CTA combines the strip maxima with a
// shuffle scan, and the only thing that crosses a chunk boundary is ONE int per row -- the running maximum P[i][last column
// of the chunk], which is also the diagonal input of the next chunk's first column for row i+1.  It travels as a tagged
// 64-bit element (value | epoch) through a per-boundary array in HBM; across GPUs the same store goes into the right
// neighbour's array mapped over NVLink (CUDA IPC), so the chunks of all GPUs form one pipeline that is
// rows + chunks * (a few rows) deep instead of rows * columns.  CTAs take chunks from a ticket, so a GPU may own more
// chunks than fit on its SMs.
#pragma once
#include "nw_common.cuh"

namespace nwb {

constexpr int kScanT = 256;                 // threads per CTA
constexpr int kScanC = 16;                  // columns per thread
constexpr int kScanW = kScanT * kScanC;     // columns per chunk
constexpr int kScanAhead = 3;               // rows the carry of the left chunk is requested ahead

struct ScanArgs {
    const uint8_t* y;                 // n row letters
    const uint8_t* x;                 // m column letters
    int n;
    long long m;
    const uint8_t* sprime;
    int S;
    int chunk0;                       // first chunk of this rank (global chunk index)
    int nchunks;                      // chunks of this rank
    int total_chunks;                 // chunks of the whole matrix
    unsigned long long* carry;        // carry[(1 + local chunk) * n + (i-1)] = (tag << 32 | P[i][last column of that chunk]); slot 0 = from the left rank
    unsigned long long* peer_carry0;  // the right rank's slot 0 (peer memory), nullptr when this rank owns the last chunk
    unsigned tag;                     // epoch tag, the same on every rank
    int* ticket;
    unsigned long long* score;        // (tag << 32 | P[n][m]) written by the owner of the last column
    unsigned long long timeout_ns;
    int* err;
};

__global__ void __launch_bounds__(kScanT) nw_scan_kernel(const ScanArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // prof[letter][column of the chunk] bytes s'(letter, x[column]); row S is all zero
    unsigned char* prof = smem_raw;
    __shared__ int wtot[2][kScanT / 32];          // per-warp strip maxima, double buffered by row parity
    __shared__ int wlast[2][kScanT / 32];         // previous-row value of every warp's last column (halo of the next warp)
    __shared__ int cin[4];                        // carry of the left chunk for rows i .. i+3 (ring)
    __shared__ int s_chunk;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int n = a.n;

    for (;;) {
        if (tid == 0) s_chunk = atomicAdd(a.ticket, 1);
        __syncthreads();
        const int lc = s_chunk;                   // local chunk
        __syncthreads();
        if (lc >= a.nchunks) break;
        const int gc = a.chunk0 + lc;             // global chunk
        const long long c0 = (long long)gc * kScanW;
        const bool has_left = gc > 0;
        const unsigned long long* cin_g = a.carry + (long long)lc * n;                       // written by the chunk to the left
        unsigned long long* cout_g = (lc + 1 == a.nchunks && a.peer_carry0 != nullptr) ? a.peer_carry0 : a.carry + (long long)(lc + 1) * n;
        const bool has_right = gc + 1 < a.total_chunks;

        // ---- profile of this chunk's columns: prof[yl][c] = s'[yl][x[c0 + c]] (zero past the end of x)
        for (int c = tid; c < kScanW; c += kScanT) {
            const long long j = c0 + c;
            const int xl = (j < a.m) ? (int)__ldg(a.x + j) : -1;
            for (int yl = 0; yl < a.S; yl++) prof[yl * kScanW + c] = (xl >= 0 && xl < a.S) ? __ldg(a.sprime + yl * a.S + xl) : (unsigned char)0;
            prof[a.S * kScanW + c] = 0;
        }
        if (tid < 4) cin[tid] = 0;
        if (tid < kScanT / 32) { wtot[0][tid] = wtot[1][tid] = 0; wlast[0][tid] = wlast[1][tid] = 0; }
        __syncthreads();

        int prev[kScanC];                          // P[i-1][this thread's columns]
#pragma unroll
        for (int k = 0; k < kScanC; k++) prev[k] = 0;
        int carry_prev = 0;                        // P[i-1][c0 - 1]: carry of the left chunk for the previous row
        unsigned long long pend = 0;               // warp 0, lanes 0..3: the outstanding request of "their" row
        const unsigned long long t0 = a.timeout_ns ? globaltimer_ns() : 0ull;
        // requests for the first rows
        if (has_left && w == 0 && lane < 4 && lane < kScanAhead && 1 + lane <= n) pend = ld_relaxed64(cin_g + lane);

        for (int i = 1; i <= n; i++) {
            const int par = i & 1;
            // ---- warp 0: land the left chunk's carry for row i, request the one for row i + kScanAhead
            if (w == 0) {
                if (has_left) {
                    if (lane == ((i - 1) & 3)) {
                        unsigned long long v = pend;
                        while ((unsigned)(v >> 32) != a.tag) {      // no __nanosleep: it oversleeps by milliseconds now and then (nw_common.cuh)
                            v = ld_relaxed64(cin_g + (i - 1));
                            if (a.timeout_ns && globaltimer_ns() - t0 > a.timeout_ns) { atomicExch(a.err, 1); break; }
                        }
                        cin[(i - 1) & 3] = (int)(unsigned)v;
                    }
                    if (lane == ((i - 1 + kScanAhead) & 3) && i + kScanAhead <= n) pend = ld_relaxed64(cin_g + (i - 1 + kScanAhead));
                }
            }
            // ---- local prefix over this thread's columns
            const unsigned yl = (unsigned)__ldg(a.y + (i - 1));
            const uint4 sw = *reinterpret_cast<const uint4*>(prof + (yl < (unsigned)a.S ? yl : (unsigned)a.S) * kScanW + tid * kScanC);
            const unsigned swv[4] = {sw.x, sw.y, sw.z, sw.w};
            int diag = __shfl_up_sync(kFull, prev[kScanC - 1], 1);
            if (lane == 0) diag = (w == 0) ? carry_prev : wlast[par ^ 1][w - 1];
            int run = 0;                           // P >= 0: neutral start of the running maximum
            int cur[kScanC];
#pragma unroll
            for (int k = 0; k < kScanC; k++) {
                const int t = add_byte(swv[k >> 2], 1u << (8 * (k & 3)), diag);
                diag = prev[k];
                run = max3(t, prev[k], run);
                cur[k] = run;
            }
            // ---- strip maxima -> exclusive prefix maxima across the CTA
            int incl = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl = max(incl, v);
            }
            int excl = __shfl_up_sync(kFull, incl, 1);
            if (lane == 0) excl = 0;
            if (lane == 31) wtot[par][w] = incl;
            __syncthreads();                        // wtot[par], cin[(i-1)&3] visible
            int cw = 0;
#pragma unroll
            for (int q = 0; q < kScanT / 32; q++) { const int v = wtot[par][q]; if (q < w) cw = max(cw, v); }
            const int cleft = has_left ? cin[(i - 1) & 3] : 0;
            const int carry = max3(cleft, cw, excl);
#pragma unroll
            for (int k = 0; k < kScanC; k++) prev[k] = max(cur[k], carry);
            carry_prev = cleft;
            // ---- hand the row's value at the chunk's last column to the chunk on the right, and the halo to the next warp
            if (lane == 31) wlast[par][w] = prev[kScanC - 1];
            if (tid == kScanT - 1) {
                if (has_right) st_relaxed64(cout_g + (i - 1), pack_tagged(prev[kScanC - 1], a.tag));
            }
            // wlast[par] is read at the top of row i+1, before that row's barrier
            __syncthreads();
        }
        // ---- the score lives at column m of the last row
        {
            const long long j0 = c0 + (long long)tid * kScanC;          // first column (0-based) of this thread
            if (a.m - 1 >= j0 && a.m - 1 < j0 + kScanC) {
                const int k = (int)(a.m - 1 - j0);
                int v = 0;
#pragma unroll
                for (int q = 0; q < kScanC; q++) if (q == k) v = prev[q];
                *a.score = pack_tagged(v, a.tag);
            }
        }
        __syncthreads();
    }
}

}  // namespace nwb
