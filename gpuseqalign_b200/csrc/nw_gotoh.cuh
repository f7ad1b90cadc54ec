// nw_gotoh.cuh -- the variants the reference lists as future work (README.md:6-29: NW_AG, SW_LG, SW_AG; its --gapeCost option is
// parsed and unused, cmd_parser.cpp:143,213): affine gaps (Gotoh) and local alignment (Smith-Waterman) for batches of short pairs,
// scores only.  One warp per pair, the band of 32*R rows swept with the lanes one column apart like nw_batch.cuh.
//
// Conventions (the same as the CPU restatement the tests check against -- there is no reference implementation to compare with, so they
// are chosen to contain the one case the reference defines: gape == gapo is its linear-gap recurrence):
//   a gap of L residues costs gapo + (L - 1) * gape
//   E[i][j] = max(E[i][j-1] + gape, H[i][j-1] + gapo)     F[i][j] = max(F[i-1][j] + gape, H[i-1][j] + gapo)
//   H[i][j] = max(H[i-1][j-1] + subst[y_i][x_j], E[i][j], F[i][j])  [, 0 for local];  global borders gapo + (k-1) * gape, local 0
//
// The shifted coordinates of the linear kernels (one IDP + one VIMNMX3 per cell) do not carry over: E and F are recurrences of their
// own.  Per cell here: two VIADDMNMX (E, F), one IDP.4A on SIGNED profile bytes (diag + subst), one VIMNMX3 (.RELU for local:
// the floor at 0 is free) and one add (H + gapo, shared by the E of the cell to the right and the F of the cell below): 5 integer
// instructions, 3 of them DPX.  Rows are aligned to the TOP of the band (padding rows below the matrix never feed real cells); a lane
// works only while its column is inside the matrix (a divergent branch around the step: idle lanes cost nothing extra), so at the end
// every lane holds its rows at the last column.
#pragma once
#include "nw_engine.cuh"
#include "nw_sweep.cuh"

namespace nwb {

struct GotohArgs {
    BatchArgs b;              // letters, per-pair metadata, scores, tickets, error flag (sprime / gap unused)
    const int* subst;         // S x S, subst[y * S + x], every entry in [-128, 127]
    int gapo, gape;
};

constexpr int kGotohNeg = -(1 << 28);      // "minus infinity" that survives a few thousand additions of a gap cost

template <int R, bool LOCAL>
__global__ void __launch_bounds__(128) nw_gotoh_batch_kernel(const GotohArgs g)
{
    constexpr int By = 32 * R, WPL = R / 4, W = 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];          // [W warps][S letters][32 lanes][R bytes] signed profile
    __shared__ signed char s_sub[(kMaxLetters + 1) * (kMaxLetters + 1)];
    const BatchArgs& a = g.b;
    const int S = a.S;
    for (int i = threadIdx.x; i < S * S; i += blockDim.x) s_sub[i] = (signed char)__ldg(g.subst + i);
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned* prof = reinterpret_cast<unsigned*>(smem_raw + (size_t)w * (size_t)S * 32 * R);
    const int src_lane = (lane + 31) & 31;
    const int go = g.gapo, ge = g.gape;
    (void)W;

    for (;;) {
        unsigned long long p = 0;
        if (lane == 0) p = a.first + atomicAdd(a.ticket, 1ull);
        p = __shfl_sync(kFull, p, 0);
        if (p >= a.npairs) break;
        const int n = (int)a.lenY[p], m = (int)a.lenX[p];
        if (n == 0 || m == 0) {                        // a border cell
            const int k = n + m;
            if (lane == 0) a.scores[p] = (LOCAL || k == 0) ? 0 : go + (k - 1) * ge;
            continue;
        }
        if (n > By) { if (lane == 0) a.scores[p] = kBatchTooTall; continue; }
        const uint8_t* y = a.letters + a.offY[p];
        const uint8_t* x = a.letters + a.offX[p];
        __syncwarp();
        // ---- signed byte profile of this lane's rows (top-aligned): prof[letter][lane][q] = subst[y[row 4q..4q+3]][letter]
        {
            unsigned yl[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = lane * R + r;
                yl[r] = i < n ? (unsigned)__ldg(y + i) : 0xffu;
                if (i < n && yl[r] >= (unsigned)S) { *a.err = 1; yl[r] = 0xffu; }
            }
            for (int l = 0; l < S; l++) {
#pragma unroll
                for (int q = 0; q < WPL; q++) {
                    unsigned wd = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const unsigned yy = yl[4 * q + k];
                        // padding rows: -128 keeps a local H at its floor; global padding rows never reach a real cell
                        const int v = yy < (unsigned)S ? (int)s_sub[yy * S + l] : -128;
                        wd |= ((unsigned)v & 0xffu) << (8 * k);
                    }
                    prof[(l * 32 + lane) * WPL + q] = wd;
                }
            }
        }
        __syncwarp();
        // ---- borders: column 0 of this lane's rows, the cell above-left of its first row
        int h[R], e[R], hgo[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane * R + r + 1;
            h[r] = LOCAL ? 0 : go + (i - 1) * ge;
            e[r] = kGotohNeg;
            hgo[r] = h[r] + go;
        }
        int dprev = (LOCAL || lane == 0) ? 0 : go + (lane * R - 1) * ge;       // H[lane*R][0]
        int fb = kGotohNeg;                                                     // F of this lane's last row at its current column
        int best = 0;
        const int steps = m + 31;
        unsigned xl_next = (lane == 0) ? (unsigned)__ldg(x) : 0u;
        for (int s = 0; s < steps; s++) {
            const int c = s - lane;
            const bool act = c >= 0 && c < m;
            const unsigned xl = xl_next;
            {   // the letter of the next step's column
                const int cn = c + 1;
                xl_next = (cn >= 0 && cn < m) ? (unsigned)__ldg(x + cn) : 0u;
            }
            // the last row of the lane above at this column (all lanes take part in the shuffles)
            int uh = __shfl_sync(kFull, h[R - 1], src_lane);
            int uf = __shfl_sync(kFull, fb, src_lane);
            if (lane == 0) { uh = LOCAL ? 0 : go + c * ge; uf = kGotohNeg; }          // matrix row 0: H[0][c+1]
            if (act) {
                unsigned letter = xl;
                if (letter >= (unsigned)S) { *a.err = 1; letter = 0; }
                unsigned wv[WPL];
#pragma unroll
                for (int q = 0; q < WPL; q++) wv[q] = prof[(letter * 32 + lane) * WPL + q];
                int diag = dprev;
                dprev = uh;
                int uhgo = uh + go;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int E = __viaddmax_s32(e[r], ge, hgo[r]);
                    const int F = __viaddmax_s32(uf, ge, uhgo);
                    const int t = __dp4a((int)wv[r >> 2], 1 << (8 * (r & 3)), diag);
                    const int H = LOCAL ? __vimax3_s32_relu(t, E, F) : __vimax3_s32(t, E, F);
                    diag = h[r];
                    h[r] = H; e[r] = E; hgo[r] = H + go;
                    uf = F; uhgo = hgo[r];
                    if (LOCAL && (r & 1)) best = __vimax3_s32(best, H, h[r - 1]);
                }
                fb = uf;
            }
        }
        __syncwarp();
        if (LOCAL) {
            best = __reduce_max_sync(kFull, best);
            if (lane == 0) a.scores[p] = best;
        } else {
            // H[n][m]: row (n - 1) % R of lane (n - 1) / R, which stopped at the last column
            int v = 0;
#pragma unroll
            for (int r = 0; r < R; r++) if (r == (n - 1) % R) v = h[r];
            v = __shfl_sync(kFull, v, (n - 1) / R);
            if (lane == 0) a.scores[p] = v;
        }
    }
}

}  // namespace nwb
