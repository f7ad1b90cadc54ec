// nw_trace.cuh -- traceback on the GPU from the sparse tile headers (the reference's
// nwtrace2_sparse scheme, nwtrace2_sparse.cpp:102-257, re-designed for the device) and the
// export of the device-resident headers in the reference's own layout.
#pragma once
#include "nw_common.cuh"

namespace nwb {

// Device headers (P-space, plain row-major) -> reference layout (H-space, tile-major;
// SURVEY.md App. A-4, producer nwalign_gpu9_mlsp_diagdiagdiag.cu:321-359, consumer
// nwtrace2_sparse.cpp:48-67):
//   hrow[(iT*tcols + jT)*(1+Bx) + k] = H[iT*By][jT*Bx + k]
//   hcol[(iT*tcols + jT)*(1+By) + k] = H[iT*By + k][jT*Bx]
// with H[i][j] = P[i][j] + (i+j)*gap, P[0][*] = P[*][0] = 0,
// HR[b*ldr + c] = (tag << 32 | P[b*By][c+1]) (b >= 1), HC[q*ldc + i0] = P[i0+1][(q+1)*Bx].
// Entries that lie outside the real matrix (padding) are written as 0; nothing consumes them.
__global__ void nw_export_headers_kernel(const unsigned long long* __restrict__ HR, long long ldr, const int* __restrict__ HC, long long ldc,
                                         int n, int m, int By, int Bx, int trows, int tcols, int gap,
                                         int* __restrict__ hrow, int* __restrict__ hcol)
{
    const long long nrow = (long long)trows * tcols * (1 + Bx);
    const long long ncol = (long long)trows * tcols * (1 + By);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nrow; e += stride) {
        long long t = e / (1 + Bx); int k = (int)(e % (1 + Bx));
        int iT = (int)(t / tcols), jT = (int)(t % tcols);
        long long i = (long long)iT * By, j = (long long)jT * Bx + k;
        int v = 0;
        if (i <= n && j <= m) {
            int P = (i == 0 || j == 0) ? 0 : (int)(unsigned)HR[(long long)iT * ldr + (j - 1)];
            v = P + (int)((i + j) * gap);
        }
        hrow[e] = v;
    }
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ncol; e += stride) {
        long long t = e / (1 + By); int k = (int)(e % (1 + By));
        int iT = (int)(t / tcols), jT = (int)(t % tcols);
        long long i = (long long)iT * By + k, j = (long long)jT * Bx;
        int v = 0;
        if (i <= n && j <= m) {
            int P = (i == 0 || j == 0) ? 0 : HC[(long long)(jT - 1) * ldc + (i - 1)];
            v = P + (int)((i + j) * gap);
        }
        hcol[e] = v;
    }
}

}  // namespace nwb
