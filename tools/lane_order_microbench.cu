// Lone-warp micro-benchmark for the order in which a lane issues the cells of its R x 32 block of the band sweep (DESIGN.md,
// "next lever"): ORDER 0 = step by step, the R rows of a step chained through their upper neighbour (what nw_sweep.cuh does);
// ORDER 1 = in-lane wavefront, row r one step behind row r-1, so the R cells issued together are independent.
// Same arithmetic (IDP.4A + VIMNMX3 per cell), same results; one warp per SM, clock64 around 256 chunks of 32 steps.
// Build (on the GPU box): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/lane_order_microbench tools/lane_order_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int R, int ORDER>
__global__ void __launch_bounds__(32) lane_kernel(int* out, long long* cyc, unsigned w, int seed)
{
    int h[R], hp[R];
#pragma unroll
    for (int r = 0; r < R; r++) { h[r] = seed + r; hp[r] = seed; }
    int uprev = seed;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; it++) {
        int U[32];
#pragma unroll
        for (int s = 0; s < 32; s++) U[s] = (seed ^ (s * 37)) + it;        // the upper neighbour's row: known one step early (K = 2)
        if (ORDER == 0) {
#pragma unroll
            for (int s = 0; s < 32; s++) {
                int up = U[s], diag = s ? U[s - 1] : uprev;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int left = h[r];
                    const int t = (int)__dp4a(w, 1u << (8 * (r & 3)), (unsigned)diag);
                    const int nv = __vimax3_s32(t, up, left);
                    diag = left; up = nv; h[r] = nv;
                }
            }
        } else {
            // hp[r] = value of row r one step before h[r]; a diagonal is issued bottom row first, so that row r-1 still holds step s
#pragma unroll
            for (int d = 0; d < 32 + R - 1; d++) {
#pragma unroll
                for (int r = R - 1; r >= 0; r--) {
                    const int s = d - r;
                    if (s < 0 || s >= 32) continue;
                    const int up = r ? h[r - 1] : U[s];
                    const int diag = r ? hp[r - 1] : (s ? U[s - 1] : uprev);
                    const int left = h[r];
                    const int t = (int)__dp4a(w, 1u << (8 * (r & 3)), (unsigned)diag);
                    const int nv = __vimax3_s32(t, up, left);
                    hp[r] = left; h[r] = nv;
                }
            }
        }
        uprev = U[31];
    }
    long long t1 = clock64();
    int acc = 0;
#pragma unroll
    for (int r = 0; r < R; r++) acc ^= h[r];
    out[blockIdx.x * 32 + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int R, int ORDER>
static void run(int* d_out, long long* d_cyc)
{
    lane_kernel<R, ORDER><<<148, 32>>>(d_out, d_cyc, 0x01020304u, 7);
    cudaDeviceSynchronize();
    lane_kernel<R, ORDER><<<148, 32>>>(d_out, d_cyc, 0x01020304u, 7);
    cudaDeviceSynchronize();
    long long c; int o;
    cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
    cudaMemcpy(&o, d_out, sizeof(o), cudaMemcpyDeviceToHost);
    printf("{\"kind\":\"lane_order\",\"rows_per_lane\":%d,\"order\":\"%s\",\"clk_per_step\":%.2f,\"clk_per_cell\":%.2f,\"checksum\":%d}\n",
           R, ORDER ? "in-lane wavefront" : "step by step (rows chained)", (double)c / (256.0 * 32.0), (double)c / (256.0 * 32.0 * R), o);
}

int main()
{
    int* d_out; long long* d_cyc;
    cudaMalloc(&d_out, sizeof(int) * 148 * 32);
    cudaMalloc(&d_cyc, sizeof(long long) * 148);
    run<4, 0>(d_out, d_cyc); run<4, 1>(d_out, d_cyc);
    run<8, 0>(d_out, d_cyc); run<8, 1>(d_out, d_cyc);
    return 0;
}
