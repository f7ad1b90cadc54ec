// nwb200_capi.cu -- the C ABI of include/nwb200.h on top of the sm_100a kernels.
//
// Host side of the hot path: what NwAlign_Gpu9_Mlsp_DiagDiagDiag's host function does
// (reference nwalign_gpu9_mlsp_diagdiagdiag.cu:368-722) minus everything that does not need to
// be inside the timed region: buffers, stream and events live in the context and grow on demand
// (the reference cudaMallocs, captures and instantiates a CUDA graph inside every align call).
#include "nw_engine.cuh"
#include "nw_fill.cuh"
#include "nw_trace.cuh"
#include "nw_scan.cuh"
#include "nw_gotoh.cuh"

using namespace nwb;

#define NWB_VERSION "nwb200 0.1 (sm_100a)"

namespace {

// ---------------------------------------------------------------------------------------------
// geometry: the analogue of nwalign_gpu9_mlsp_diagdiagdiag.cu:384-431
int plan_geometry(nwb200_ctx* c, int n, int m, const nwb200_params* p)
{
    Geometry& g = c->g;
    int R = 0, W = 4, K = 2, Bx = (m <= 65536) ? 256 : 512;      // snapshot spacing: the walkers' windows shrink with it, the snapshot volume grows
    if (p) {
        if (p->rows_per_lane) R = p->rows_per_lane;
        if (p->warps_per_block) W = p->warps_per_block;
        if (p->tile_cols) Bx = p->tile_cols;
        if (p->reserved) K = p->reserved;
    }
    // one warp per SM sub-partition while that covers all rows; very long pairs take 16 rows per lane (2.7 instead of 3.4 instructions
    // per cell, half as many band hand-offs: 200k^2 fill 13.4 -> 11.6 ms, traceback 2.55 -> 2.8 ms; 300k^2 27.2 -> 25.2 / 4.7 -> 5.4 ms;
    // below ~152k rows the 16-row bands are few enough for the origin maps to ride in the fill launch, which is slow at R = 16)
    // (across GPUs -- plan_world > 1 -- the critical path through the first band over ALL columns decides, and a step is shorter at R = 8:
    //  200k^2 on 4 GPUs 15.2 ms per fill + traceback at R = 8 against 18.5 ms at R = 16)
    if (R == 0) R = ((long long)n <= 4LL * 32 * 4 * c->sm_count) ? 4 : ((n >= 180000 && c->plan_world <= 1) ? 16 : 8);
    if (!(p && p->rows_per_lane)) { if (const char* e = getenv("NWB200_ROWS_PER_LANE")) { const int v = atoi(e); if (v == 4 || v == 8 || v == 16) R = v; } }      // developer switch
    if (!(R == 4 || R == 8 || R == 16) || !(W == 1 || W == 4) || Bx < 32 || (Bx % 32) != 0 || Bx > (R == 16 ? 512 : 1024) || !(K == 1 || K == 2))
        return fail(c, NWB200_ERR_INVALID_VALUE, "unsupported tile parameters (rows_per_lane in {4,8,16}, warps_per_block in {1,4}, tile_cols multiple of 32 up to 1024, skew in {1,2})");
    const int By = R * 32;
    long long nb = ((long long)n + By - 1) / By;
    if (nb < 1) nb = 1;
    g.R = R; g.W = W; g.K = K; g.By = By; g.Bx = Bx; g.n = n; g.m = m;
    g.nb = (int)nb;
    g.pad = (int)(nb * By - n);
    g.tcols = (m + Bx - 1) / Bx; if (g.tcols < 1) g.tcols = 1;
    g.nlc = (m + 31 * K + 31) / 32;                          // Sched<R,K>::nlc(m)
    g.ldr = ((long long)kPadL + 32LL * g.nlc + 32 + 31) / 32 * 32;
    g.snap_chunks = Bx / 32;
    g.nsnap = g.nlc / g.snap_chunks;
    g.npad = nb * By;
    return NWB200_SUCCESS;
}

// half width (columns) of the traceback's corridor around the line to the origin; 0: no corridor (nwb200_capi_trace.inc, nw_trace.cuh)
long long trace_corridor_halfwidth(const Geometry& g)
{
    long long d = g.m / 32 > 4096 ? g.m / 32 : 4096;
    if (const char* e = getenv("NWB200_CORRIDOR")) d = atoll(e);
    return d > 0 ? d : 0;
}

template <int R, int K, int W>
int launch_fill_t(nwb200_ctx* c, const FillArgs& a, int grid, int cluster)
{
    size_t smem = Sched<R, K>::warp_smem_bytes(c->S) * W + (a.grouped ? sizeof(int) * Sched<R, K>::VR : 0);      // + the sink ring
    // Few units (the latency-bound single pair): every CTA gets an SM of its own -- the block scheduler does co-locate CTAs while
    // other SMs are idle (measured: fill warps that share their SM sub-partition with a map warp ran 1.5x slower) -- by asking
    // for more than half of an SM's shared memory.
    if (grid <= c->sm_count && smem < 120 * 1024) smem = 120 * 1024;
    // cudaFuncSetAttribute applies to the CURRENT device: cached per kernel instance (function-local static of the template) AND device
    static size_t attr_set_dev[64] = {};
    static bool nonportable_dev[64] = {};
    size_t& attr_set = attr_set_dev[c->device & 63];
    bool& nonportable = nonportable_dev[c->device & 63];
    if (smem > 48 * 1024 && smem > attr_set) {
        cudaError_t e = cudaFuncSetAttribute(nw_fill_kernel<R, K, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(c, NWB200_ERR_KERNEL_FAILURE, "cudaFuncSetAttribute(fill)", e);
        attr_set = smem;
    }
    cudaError_t e = cudaSuccess;
    if (cluster > 8 && !nonportable) {
        e = cudaFuncSetAttribute(nw_fill_kernel<R, K, W>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return fail(c, NWB200_ERR_KERNEL_FAILURE, "cudaFuncSetAttribute(non-portable cluster)", e);
        nonportable = true;
    }
    if (cluster > 1) {
        // thread-block clusters: the hand-off ring of the last warp of a CTA is the shared memory of the next CTA of its cluster
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(W * 32); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, nw_fill_kernel<R, K, W>, a);
        if (e != cudaSuccess && cluster > 8) {      // a non-portable cluster that cannot be placed: the portable size always can
            (void)cudaGetLastError();
            at[0].val.clusterDim.x = 8;
            e = cudaLaunchKernelEx(&cfg, nw_fill_kernel<R, K, W>, a);
        }
    } else {
        nw_fill_kernel<R, K, W><<<grid, W * 32, smem, c->stream>>>(a);
        e = cudaGetLastError();
    }
    c->launches++;
    if (e != cudaSuccess) return fail(c, NWB200_ERR_KERNEL_FAILURE, "fill kernel launch", e);
    {
        static cudaFuncAttributes at_dev[64];
        static bool have[64] = {};
        if (!have[c->device & 63]) { have[c->device & 63] = cudaFuncGetAttributes(&at_dev[c->device & 63], nw_fill_kernel<R, K, W>) == cudaSuccess; (void)cudaGetLastError(); }
        if (have[c->device & 63]) {
            const cudaFuncAttributes& at = at_dev[c->device & 63];
            c->fill_regs = at.numRegs; c->fill_smem_static = at.sharedSizeBytes; c->fill_local = at.localSizeBytes;
        }
        c->fill_threads = W * 32; c->fill_blocks = grid; c->fill_smem_dynamic = smem;
    }
    return NWB200_SUCCESS;
}

int launch_fill(nwb200_ctx* c, const FillArgs& a)
{
    const Geometry& g = c->g;
    // persistent grid: one CTA of W warps per SM while that covers every band (each warp alone on its SM
    // sub-partition: the single pair is latency bound), otherwise up to 16 warps per SM
    long long ctas_needed = ((a.map ? (a.map_inline ? (long long)g.nb : (a.map_half ? 3LL * g.nb - 2 : 2LL * g.nb - 1)) : (long long)g.nb * a.nq) + g.W - 1) / g.W;
    int cluster = 1;
    if (a.grouped) {
        // cluster size: as many CTAs as a chain of bands can use, at most the portable 8
        if (c->cluster_max > 1 && g.W == 4) cluster = g.nb > 32 ? 16 : (g.nb > 16 ? 8 : (g.nb > 8 ? 4 : (g.nb > 4 ? 2 : 1)));
        if (cluster > c->cluster_max) cluster = c->cluster_max;
        const long long per = (long long)g.W * cluster;
        ctas_needed = (((long long)g.nb + per - 1) / per) * (1 + (a.map ? (a.map_half ? 2 : 1) : 0)) * cluster;
    }
    int per_sm = 16 / g.W; if (per_sm < 1) per_sm = 1;
    long long grid = (long long)c->sm_count * per_sm;
    if (grid > ctas_needed) grid = ctas_needed;
    if (a.order != nullptr) {
        // column-block wavefront: one unit per band is runnable at any time (on ONE of the ranks); warps beyond that only hold tickets and
        // poll, on the schedulers of the warps that work.  Window = this rank's share of the bands + half as many again for the hand-overs.
        long long window = ((long long)g.nb + a.world - 1) / a.world;
        window += window / 2 + 32;
        if (const char* e = getenv("NWB200_WAVE_WARPS")) { const long long v = atoll(e); if (v > 0) window = v; }      // developer switch
        const long long cap = (window + g.W - 1) / g.W;
        if (grid > cap) grid = cap;
    }
    grid -= grid % cluster;
    if (grid < cluster) grid = cluster;
#define NWB_CASE(R_, K_, W_) if (g.R == R_ && g.K == K_ && g.W == W_) return launch_fill_t<R_, K_, W_>(c, a, (int)grid, cluster)
    NWB_CASE(4, 2, 4); NWB_CASE(4, 1, 4); NWB_CASE(8, 2, 4); NWB_CASE(8, 1, 4); NWB_CASE(16, 2, 4); NWB_CASE(16, 1, 4);
    NWB_CASE(4, 2, 1); NWB_CASE(8, 2, 1);
#undef NWB_CASE
    return fail(c, NWB200_ERR_INVALID_VALUE, "no kernel instance for these tile parameters");
}

}  // namespace

extern "C" {

const char* nwb200_version(void) { return NWB_VERSION; }

int nwb200_create(nwb200_ctx** out, int device)
{
    if (!out) return NWB200_ERR_INVALID_VALUE;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return NWB200_ERR_CUDA_GENERAL;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return NWB200_ERR_CUDA_GENERAL;
    if (prop.major < 10) return NWB200_ERR_CUDA_GENERAL;     // sm_100a code only: no other-arch / CPU fallback
    if (cudaSetDevice(device) != cudaSuccess) return NWB200_ERR_CUDA_GENERAL;
    nwb200_ctx* c = new (std::nothrow) nwb200_ctx();
    if (!c) return NWB200_ERR_MEMORY_ALLOCATION;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return NWB200_ERR_CUDA_GENERAL; }
    for (auto& e : c->ev) if (cudaEventCreate(&e) != cudaSuccess) { delete c; return NWB200_ERR_CUDA_GENERAL; }
    if (c->h_small.ensure(4096) != cudaSuccess) { delete c; return NWB200_ERR_MEMORY_ALLOCATION; }
    if (cudaGetSymbolAddress((void**)&c->d_timeout_flag, g_wait_timeout) != cudaSuccess) { delete c; return NWB200_ERR_CUDA_GENERAL; }
    *out = c;
    return NWB200_SUCCESS;
}

void nwb200_destroy(nwb200_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->wave_peer_base) { cudaIpcCloseMemHandle(c->wave_peer_base); c->wave_peer_base = nullptr; }
    for (int r = 0; r < 16; r++) {
        if (c->wave_peer_hr[r]) cudaIpcCloseMemHandle(c->wave_peer_hr[r]);
        if (c->wave_peer_snap[r]) cudaIpcCloseMemHandle(c->wave_peer_snap[r]);
    }
    for (DevBuf* b : {&c->d_sprime, &c->d_subst, &c->d_y, &c->d_HR, &c->d_snap, &c->d_lastcol, &c->d_sync, &c->d_order,
                      &c->d_map, &c->d_MID, &c->d_tmeta, &c->d_ops, &c->d_dense, &c->d_export, &c->d_HR2, &c->d_cut,
                      &c->d_bletters, &c->d_bmeta, &c->d_bscores, &c->d_bticket, &c->d_bmoves, &c->d_bmoff, &c->d_bcnt, &c->d_dbg, &c->d_wave, &c->d_corr})
        b->release();
    c->h_stage.release(); c->h_small.release(); c->h_trace.release(); c->h_batch.release(); c->h_export.release();
    c->h_bscores.release(); c->h_bmoves.release();
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (auto& e : c->slice_ev) if (e) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (auto& e : c->slice_done_ev) if (e) cudaEventDestroy(e);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int nwb200_set_scoring(nwb200_ctx* c, const int32_t* subst, int substsz, int gap)
{
    if (!c || !subst || substsz < 1 || substsz > kMaxLetters) return fail(c, NWB200_ERR_INVALID_VALUE, "bad substitution matrix size");
    cudaSetDevice(c->device);
    const int S = substsz;
    std::vector<uint8_t> sp((size_t)S * S);
    int mx = 0;
    for (int i = 0; i < S * S; i++) {
        long long v = (long long)subst[i] - 2LL * gap;       // s' = s - 2*gap
        if (v < 0) v = 0;                                     // exact: a negative s' can never win (P is monotone along rows)
        if (v > 255) return fail(c, NWB200_ERR_INVALID_VALUE, "subst - 2*gap exceeds the byte profile range (0..255)");
        sp[i] = (uint8_t)v;
        if (v > mx) mx = (int)v;
    }
    CU(c, c->d_sprime.ensure(sp.size()), NWB200_ERR_MEMORY_ALLOCATION, "alloc sprime");
    CU(c, c->d_subst.ensure(sizeof(int32_t) * S * S), NWB200_ERR_MEMORY_ALLOCATION, "alloc subst");
    CU(c, cudaMemcpyAsync(c->d_sprime.p, sp.data(), sp.size(), cudaMemcpyHostToDevice, c->stream), NWB200_ERR_MEMORY_TRANSFER, "copy sprime");
    CU(c, cudaMemcpyAsync(c->d_subst.p, subst, sizeof(int32_t) * S * S, cudaMemcpyHostToDevice, c->stream), NWB200_ERR_MEMORY_TRANSFER, "copy subst");
    CU(c, cudaStreamSynchronize(c->stream), NWB200_ERR_MEMORY_TRANSFER, "sync scoring");
    c->subst.assign(subst, subst + (size_t)S * S);
    c->S = S; c->gap = gap; c->max_sprime = mx; c->scoring_set = true;
    c->pair_resident = false; c->batch_resident = false;
    return NWB200_SUCCESS;
}

static int check_range(nwb200_ctx* c, long long n, long long m)
{
    if (n < 1 || m < 1 || n > 0x7ffffff0LL || m > 0x7ffffff0LL) return fail(c, NWB200_ERR_INVALID_VALUE, "sequence lengths must be in [1, 2^31)");
    long long mn = n < m ? n : m;
    long long g = c->gap < 0 ? -(long long)c->gap : c->gap;
    if (mn * c->max_sprime + (n + m) * g >= 0x7fffffffLL) return fail(c, NWB200_ERR_INVALID_VALUE, "scores would overflow int32");
    return NWB200_SUCCESS;
}

static size_t stage_off_x(long long n) { return ((size_t)n + 63) & ~(size_t)63; }

static int upload_common(nwb200_ctx* c, const uint8_t* stage_y, long long n, const uint8_t* stage_x, long long m, const nwb200_params* p)
{
    int rc = plan_geometry(c, (int)n, (int)m, p);
    if (rc) return rc;
    // y and x live in ONE device buffer (x at x_off) and are staged the same way: one H2D copy per pair
    c->x_off = stage_off_x(n);
    if (stage_x != stage_y + c->x_off) return fail(c, NWB200_ERR_INVALID_VALUE, "internal: staging layout");
    CU(c, c->d_y.ensure(c->x_off + (size_t)m + 64), NWB200_ERR_MEMORY_ALLOCATION, "alloc letters");
    CU(c, cudaEventRecord(c->ev[0], c->stream), NWB200_ERR_CUDA_GENERAL, "event");
    CU(c, cudaMemcpyAsync(c->d_y.p, stage_y, c->x_off + (size_t)m, cudaMemcpyHostToDevice, c->stream), NWB200_ERR_MEMORY_TRANSFER, "H2D letters");
    CU(c, cudaEventRecord(c->ev[1], c->stream), NWB200_ERR_CUDA_GENERAL, "event");
    c->pair_resident = true; c->headers_valid = false; c->fill_done = false; c->trace_done = false;
    return NWB200_SUCCESS;
}

int nwb200_upload_pair_u8(nwb200_ctx* c, const uint8_t* y, int64_t n, const uint8_t* x, int64_t m, const nwb200_params* p)
{
    if (!c || !y || !x) return fail(c, NWB200_ERR_INVALID_VALUE, "null argument");
    if (!c->scoring_set) return fail(c, NWB200_ERR_INVALID_VALUE, "nwb200_set_scoring has not been called");
    cudaSetDevice(c->device);
    int rc = check_range(c, n, m);
    if (rc) return rc;
    CU(c, cudaStreamSynchronize(c->stream), NWB200_ERR_CUDA_GENERAL, "sync before staging");
    CU(c, c->h_stage.ensure(stage_off_x(n) + (size_t)m + 64), NWB200_ERR_MEMORY_ALLOCATION, "alloc pinned staging");
    uint8_t* sy = c->h_stage.as<uint8_t>();
    uint8_t* sx = sy + stage_off_x(n);
    const int S = c->S;
    unsigned bad = 0;
    for (int64_t i = 0; i < n; i++) { uint8_t v = y[i]; bad |= (v >= S); sy[i] = v; }
    for (int64_t j = 0; j < m; j++) { uint8_t v = x[j]; bad |= (v >= S); sx[j] = v; }
    if (bad) return fail(c, NWB200_ERR_INVALID_VALUE, "letter index outside the substitution matrix");
    return upload_common(c, sy, n, sx, m, p);
}

static int upload_pair_i32(nwb200_ctx* c, const int32_t* seqY, int64_t adjrows, const int32_t* seqX, int64_t adjcols, const nwb200_params* p)
{
    if (!c || !seqY || !seqX) return fail(c, NWB200_ERR_INVALID_VALUE, "null argument");
    if (!c->scoring_set) return fail(c, NWB200_ERR_INVALID_VALUE, "nwb200_set_scoring has not been called");
    cudaSetDevice(c->device);
    const int64_t n = adjrows - 1, m = adjcols - 1;
    int rc = check_range(c, n, m);
    if (rc) return rc;
    CU(c, cudaStreamSynchronize(c->stream), NWB200_ERR_CUDA_GENERAL, "sync before staging");
    CU(c, c->h_stage.ensure(stage_off_x(n) + (size_t)m + 64), NWB200_ERR_MEMORY_ALLOCATION, "alloc pinned staging");
    uint8_t* sy = c->h_stage.as<uint8_t>();
    uint8_t* sx = sy + stage_off_x(n);
    const unsigned S = (unsigned)c->S;
    unsigned bad = 0;
    for (int64_t i = 0; i < n; i++) { unsigned v = (unsigned)seqY[i + 1]; bad |= (v >= S); sy[i] = (uint8_t)v; }   // element 0 is the dummy header
    for (int64_t j = 0; j < m; j++) { unsigned v = (unsigned)seqX[j + 1]; bad |= (v >= S); sx[j] = (uint8_t)v; }
    if (bad) return fail(c, NWB200_ERR_INVALID_VALUE, "letter index outside the substitution matrix");
    return upload_common(c, sy, n, sx, m, p);
}

int nwb200_upload_pair_i32(nwb200_ctx* c, const int32_t* seqY, int64_t adjrows, const int32_t* seqX, int64_t adjcols, const nwb200_params* p)
{
    return upload_pair_i32(c, seqY, adjrows, seqX, adjcols, p);
}

int nwb200_fill_resident(nwb200_ctx* c, int flags)
{
    if (!c || !c->pair_resident) return fail(c, NWB200_ERR_INVALID_VALUE, "no pair resident on the device");
    cudaSetDevice(c->device);
    const Geometry& g = c->g;
    const bool keep = (flags & (NWB200_KEEP_HEADERS | NWB200_WITH_TRACE)) != 0;
    {
        const size_t before = c->d_HR.cap;
        CU(c, c->d_HR.ensure(sizeof(unsigned long long) * (size_t)(g.nb + 1) * (size_t)g.ldr), NWB200_ERR_MEMORY_ALLOCATION, "alloc header rows");
        if (c->d_HR.cap != before)   // fresh memory may hold anything: clear the tag halves once
            CU(c, cudaMemsetAsync(c->d_HR.p, 0, c->d_HR.cap, c->stream), NWB200_ERR_CUDA_GENERAL, "memset header rows");
    }
    const size_t snap_ints = (size_t)g.nb * (size_t)(g.nsnap > 0 ? g.nsnap : 1) * 32 * (size_t)(g.R + 4);
    if (keep) CU(c, c->d_snap.ensure(sizeof(int) * snap_ints), NWB200_ERR_MEMORY_ALLOCATION, "alloc snapshots");
    CU(c, c->d_sync.ensure(sizeof(int) * 8), NWB200_ERR_MEMORY_ALLOCATION, "alloc ticket");
    CU(c, cudaEventRecord(c->ev[2], c->stream), NWB200_ERR_CUDA_GENERAL, "event");
    CU(c, cudaMemsetAsync(c->d_sync.p, 0, sizeof(int) * 8, c->stream), NWB200_ERR_CUDA_GENERAL, "memset ticket");
    FillArgs a = {};
    a.y = c->d_y.as<uint8_t>(); a.x = (c->d_y.as<uint8_t>() + c->x_off); a.n = g.n; a.m = g.m;
    a.sprime = c->d_sprime.as<uint8_t>(); a.S = c->S;
    a.HR = c->d_HR.as<unsigned long long>(); a.ldr = g.ldr;
    a.snap = (keep && g.nsnap > 0) ? c->d_snap.as<int>() : nullptr;
    a.nsnap = g.nsnap; a.snap_chunks = g.snap_chunks;
    a.hr_stride = 0; a.wc = (g.m + 31) / 32 * 32; a.nq = 1; a.rank = 0; a.world = 1;
    a.recv = nullptr; a.recv_flag = nullptr; a.peer_recv = nullptr; a.peer_flag = nullptr; a.recv_stride = 0;
    a.lastcol = nullptr; a.timeout_ns = 0; a.err = nullptr;
    c->epoch++;
    if (c->epoch == 0) c->epoch = 1;
    a.tag = c->epoch; a.xtag = 0; a.ticket = c->d_sync.as<int>(); a.order = nullptr;
    a.nb = g.nb; a.pad = g.pad;
    a.pd = 2; a.negg = -c->gap; a.map = nullptr;
    // the origin maps of the traceback ride along in the fill launch while every unit still gets an SM sub-partition of its own
    // Few bands: map units shadow the fill units on idle SM sub-partitions (one launch).  Many bands: the maps come from
    // nw_map_kernel in nwb200_trace_resident; sweeping every band once WITH origin labels (map_inline) measured no better
    // (200k x 200k: 45.6 ms vs 43.7 ms, profiles/), it stays available as a developer option.
    const bool shadow = 2LL * g.nb - 1 <= 4LL * c->sm_count;
    a.map_inline = (!shadow && c->inline_map) ? 1 : 0;
    // half-band map units (two rows per lane) keep pace with the fill units; they need one SM sub-partition each as well
    const bool half = shadow && !a.map_inline && c->half_map && 3LL * g.nb - 2 <= 4LL * c->sm_count;
    a.map_half = half ? 1 : 0; a.MID = nullptr;
    if (keep && c->fuse_map && g.nb > 1 && (shadow || a.map_inline)) {
        CU(c, c->d_map.ensure(sizeof(int) * (size_t)(half ? 2 : 1) * (size_t)g.nb * (size_t)g.ldr), NWB200_ERR_MEMORY_ALLOCATION, "alloc maps");
        a.map = c->d_map.as<int>();
        if (half) {
            const size_t before = c->d_MID.cap;
            CU(c, c->d_MID.ensure(sizeof(unsigned long long) * (size_t)g.nb * (size_t)g.ldr), NWB200_ERR_MEMORY_ALLOCATION, "alloc middle rows");
            if (c->d_MID.cap != before) CU(c, cudaMemsetAsync(c->d_MID.p, 0, c->d_MID.cap, c->stream), NWB200_ERR_CUDA_GENERAL, "memset middle rows");
            a.MID = c->d_MID.as<unsigned long long>();
        }
    }
    c->map_valid = a.map != nullptr; c->map_is_half = a.map != nullptr && half;
    a.grouped = (c->grouped && !a.map_inline) ? 1 : 0;
    a.dbg = nullptr; a.dbg_mode = c->dbg_mode & 0xff; a.slack = (c->dbg_mode >> 8) ? (c->dbg_mode >> 8) - 1 : 1;
    if (c->dbg_stamps) {
        CU(c, c->d_dbg.ensure(sizeof(unsigned long long) * 4 * 3 * (size_t)g.nb), NWB200_ERR_MEMORY_ALLOCATION, "alloc debug stamps");
        a.dbg = c->d_dbg.as<unsigned long long>();
    }
    int rc = launch_fill(c, a);
    if (rc) return rc;
    CU(c, cudaEventRecord(c->ev[3], c->stream), NWB200_ERR_CUDA_GENERAL, "event");
    c->fill_done = true; c->headers_valid = keep; c->trace_done = false;
    return NWB200_SUCCESS;
}

int nwb200_fetch_score(nwb200_ctx* c, int32_t* align_cost)
{
    if (!c || !align_cost || !c->fill_done) return fail(c, NWB200_ERR_INVALID_VALUE, "no fill has been run");
    cudaSetDevice(c->device);
    const Geometry& g = c->g;
    unsigned long long* hs = c->h_small.as<unsigned long long>();
    const bool with_moves = c->trace_done && c->moves_on_host;      // NWB200_WITH_TRACE: score element and wait flag arrive in the move list's header
    if (!with_moves) {
        // the score is the last real element of the bottom row of the last band
        const unsigned long long* src = c->d_HR.as<unsigned long long>() + (long long)g.nb * g.ldr + kPadL + (g.m - 1);
        CU(c, cudaEventRecord(c->ev[4], c->stream), NWB200_ERR_CUDA_GENERAL, "event");
        CU(c, cudaMemcpyAsync(hs, src, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream), NWB200_ERR_MEMORY_TRANSFER, "D2H score");
        CU(c, cudaMemcpyAsync(hs + 1, c->d_timeout_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream), NWB200_ERR_MEMORY_TRANSFER, "D2H wait flag");
        CU(c, cudaEventRecord(c->ev[5], c->stream), NWB200_ERR_CUDA_GENERAL, "event");
    }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return fail(c, NWB200_ERR_KERNEL_FAILURE, "fill kernel execution", e);
    if (with_moves) {
        const long long* hd = c->h_trace.as<long long>();
        hs[0] = (unsigned long long)hd[2];
        hs[1] = (unsigned long long)hd[3];
    }
    if (*reinterpret_cast<const int*>(hs + 1) != 0) {
        int zero = 0;
        cudaMemcpyToSymbol(g_wait_timeout, &zero, sizeof(int));
        return fail(c, NWB200_ERR_INVALID_RESULT, "a band waited too long for its header row (producer never published)");
    }
    if ((unsigned)(hs[0] >> 32) != c->epoch) return fail(c, NWB200_ERR_INVALID_RESULT, "the fill did not publish the score element");
    // un-shift: H[n][m] = P[n][m] + (n+m)*gap
    *align_cost = (int32_t)((long long)(int)(unsigned)hs[0] + ((long long)g.n + g.m) * c->gap);
    c->timing.align_cpy_dev = ev_ms(c->ev[0], c->ev[1]);
    c->timing.align_calc = ev_ms(c->ev[2], c->ev[3]);
    c->timing.align_cpy_host = with_moves ? ev_ms(c->ev[8], c->ev[9]) : ev_ms(c->ev[4], c->ev[5]);
    return NWB200_SUCCESS;
}

static void fill_hdr_info(const nwb200_ctx* c, nwb200_hdr_info* h)
{
    if (!h) return;
    const Geometry& g = c->g;
    h->tile_rows = g.By; h->tile_cols = g.Bx; h->trows = g.nb; h->tcols = g.tcols;
    h->hrow_elems = (int64_t)g.nb * g.tcols * (1 + g.Bx);
    h->hcol_elems = (int64_t)g.nb * g.tcols * (1 + g.By);
}

int nwb200_get_hdr_info(const nwb200_ctx* c, nwb200_hdr_info* hdr)
{
    if (!c || !hdr || !c->pair_resident) return NWB200_ERR_INVALID_VALUE;
    fill_hdr_info(c, hdr);
    return NWB200_SUCCESS;
}

static int enqueue_moves_copy(nwb200_ctx* c);

int nwb200_align_pair_u8(nwb200_ctx* c, const uint8_t* y, int64_t n, const uint8_t* x, int64_t m,
                         const nwb200_params* p, int flags, int32_t* align_cost, nwb200_hdr_info* hdr)
{
    int rc = nwb200_upload_pair_u8(c, y, n, x, m, p);
    if (rc) return rc;
    rc = nwb200_fill_resident(c, flags);
    if (rc) return rc;
    if (flags & NWB200_WITH_TRACE) {
        rc = nwb200_trace_resident(c);
        if (rc) return rc;
        rc = enqueue_moves_copy(c);
        if (rc) return rc;
    }
    rc = nwb200_fetch_score(c, align_cost);
    if (rc) return rc;
    fill_hdr_info(c, hdr);
    return NWB200_SUCCESS;
}

int nwb200_align_pair_i32(nwb200_ctx* c, const int32_t* seqY, int64_t adjrows, const int32_t* seqX, int64_t adjcols,
                          const nwb200_params* p, int flags, int32_t* align_cost, nwb200_hdr_info* hdr)
{
    int rc = upload_pair_i32(c, seqY, adjrows, seqX, adjcols, p);
    if (rc) return rc;
    rc = nwb200_fill_resident(c, flags);
    if (rc) return rc;
    if (flags & NWB200_WITH_TRACE) {
        rc = nwb200_trace_resident(c);
        if (rc) return rc;
        rc = enqueue_moves_copy(c);
        if (rc) return rc;
    }
    rc = nwb200_fetch_score(c, align_cost);
    if (rc) return rc;
    fill_hdr_info(c, hdr);
    return NWB200_SUCCESS;
}

#include "nwb200_capi_trace.inc"
#include "nwb200_capi_batch.inc"
#include "nwb200_capi_wave.inc"
#include "nwb200_capi_scan.inc"

// developer aid (not part of the public header): per-band globaltimer stamps of the fills that follow
NWB200_API int nwb200_debug_band_stamps(nwb200_ctx* c, int enable, int mode, unsigned long long* out, int max_bands)
{
    if (!c) return NWB200_ERR_INVALID_VALUE;
    c->dbg_stamps = (enable & 1) != 0; c->dbg_mode = mode;
    c->fuse_map = (enable & 4) == 0; c->inline_map = (enable & 8) != 0; c->half_map = (enable & 16) == 0; c->grouped = (enable & 32) == 0;
    c->cluster_max = (enable & 64) ? 1 : ((enable & 128) ? 2 : ((enable & 256) ? 4 : ((enable & 512) ? 8 : 16)));
    if (out && c->fill_done && c->d_dbg.p) {
        int nb = 3 * c->g.nb < max_bands ? 3 * c->g.nb : max_bands;
        cudaStreamSynchronize(c->stream);
        cudaMemcpy(out, c->d_dbg.p, sizeof(unsigned long long) * 4 * nb, cudaMemcpyDeviceToHost);
        return nb;
    }
    return 0;
}

int nwb200_last_cuda_error(const nwb200_ctx* c) { return c ? (int)c->last_cuda : 0; }
const char* nwb200_last_error(const nwb200_ctx* c) { return c ? c->last_error.c_str() : "null context"; }
int nwb200_get_timing(const nwb200_ctx* c, nwb200_timing* out)
{
    if (!c || !out) return NWB200_ERR_INVALID_VALUE;
    *out = c->timing;
    return NWB200_SUCCESS;
}
void* nwb200_stream(const nwb200_ctx* c) { return c ? (void*)c->stream : nullptr; }
int nwb200_sync(nwb200_ctx* c)
{
    if (!c) return NWB200_ERR_INVALID_VALUE;
    cudaSetDevice(c->device);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return fail(c, NWB200_ERR_KERNEL_FAILURE, "stream synchronize", e);
    return NWB200_SUCCESS;
}
int nwb200_kernel_launches(const nwb200_ctx* c) { return c ? c->launches : 0; }
int nwb200_get_memory_usage(const nwb200_ctx* c, nwb200_mem_usage* out)
{
    if (!c || !out) return NWB200_ERR_INVALID_VALUE;
    memset(out, 0, sizeof(*out));
    for (const DevBuf* b : {&c->d_sprime, &c->d_subst, &c->d_y, &c->d_HR, &c->d_snap, &c->d_lastcol, &c->d_sync, &c->d_order,
                            &c->d_map, &c->d_MID, &c->d_tmeta, &c->d_ops, &c->d_dense, &c->d_export, &c->d_HR2, &c->d_cut,
                            &c->d_bletters, &c->d_bmeta, &c->d_bscores, &c->d_bticket, &c->d_bmoves, &c->d_bmoff, &c->d_bcnt, &c->d_dbg, &c->d_wave, &c->d_corr})
        out->device_bytes += b->cap;
    for (const PinBuf* b : {&c->h_stage, &c->h_small, &c->h_trace, &c->h_batch, &c->h_export, &c->h_bscores, &c->h_bmoves}) out->pinned_host_bytes += b->cap;
    out->regs_per_thread = c->fill_regs; out->threads_per_block = c->fill_threads; out->blocks = c->fill_blocks;
    out->shared_bytes = (uint64_t)(c->fill_smem_static + c->fill_smem_dynamic) * (uint64_t)c->fill_blocks;
    out->local_bytes = (uint64_t)c->fill_local * (uint64_t)c->fill_threads * (uint64_t)c->fill_blocks;
    out->register_bytes = (uint64_t)c->fill_regs * 4u * (uint64_t)c->fill_threads * (uint64_t)c->fill_blocks;
    return NWB200_SUCCESS;
}
const char* nwb200_batch_kernel_name(const nwb200_ctx* c) { return c ? c->batch_kernel : ""; }

}  // extern "C"
