// nw_batch_launch.cu -- the batch kernels (nw_batch.cuh, nw_batch2.cuh) and their launchers: a translation unit of its own so that
// libnwb200.so builds in parallel (the other one is nwb200_capi.cu: the C ABI + the single-pair, traceback and scan kernels).
#include "nw_engine.cuh"
#include "nw_batch.cuh"
#include "nw_batch2.cuh"
#include "nw_batch3.cuh"
#include "nw_gotoh.cuh"

namespace nwb {
namespace {

template <int R, bool TMA = false>
int launch_batch_t(nwb200_ctx* c, const BatchArgs& a)
{
    constexpr int W = 4;
    size_t smem = Sched<R, 1>::warp_smem_bytes(c->S) * W;
    if (TMA) smem = ((smem + 15) & ~(size_t)15) + W * batch_tma_warp_bytes();
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(nw_batch_kernel<R, W, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(c, NWB200_ERR_KERNEL_FAILURE, "cudaFuncSetAttribute(batch)", e);
    }
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nw_batch_kernel<R, W, TMA>, W * 32, smem);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)c->sm_count * per_sm;
    const long long need = ((long long)(a.npairs - a.first) + W - 1) / W;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    nw_batch_kernel<R, W, TMA><<<(int)grid, W * 32, smem, c->stream>>>(a);
    c->launches++;
    c->batch_kernel = TMA ? "nw_batch_kernel (letters by TMA bulk copies)" : "nw_batch_kernel";
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, NWB200_ERR_KERNEL_FAILURE, "batch kernel launch", e);
    return NWB200_SUCCESS;
}

// Two pairs per warp in packed 16-bit halves (nw_batch2.cuh).  The CTA size follows from the shared memory a warp needs
// (two profiles): the CTA shape that puts most warps on an SM.
template <int R, int W, int NP>
int launch_batch2_w(nwb200_ctx* c, const BatchArgs& a, size_t smem, int per_sm)
{
    long long grid = (long long)c->sm_count * per_sm;
    const long long need = ((long long)(a.npairs - a.first) + 2 * W - 1) / (2 * W);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    nw_batch2_kernel<R, W, NP><<<(int)grid, W * 32, smem, c->stream>>>(a);
    c->launches++;
    c->batch_kernel = "nw_batch2_kernel";
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, NWB200_ERR_KERNEL_FAILURE, "packed batch kernel launch", e);
    return NWB200_SUCCESS;
}

template <int R, int W, int NP>
int batch2_occupancy(nwb200_ctx* c, size_t* smem_out)
{
    const size_t smem = Sched2<R>::warp_smem_bytes(c->S) * W;
    *smem_out = smem;
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(nw_batch2_kernel<R, W, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nw_batch2_kernel<R, W, NP>, W * 32, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    return per_sm;
}

template <int R, int NP>
int launch_batch2_t(nwb200_ctx* c, const BatchArgs& a)
{
    // CTA shapes tried: 4, 7 and 16 warps (16 warps of 14.3 KB are one CTA per SM at S = 25); the one that puts most warps on an SM
    // wins.  The choice depends on the alphabet size only (all devices of a process are B200s): kept per (R, NP, S).
    // (per device: the opt-in to more than 48 KB of dynamic shared memory is a per-device function attribute)
    struct Choice { int S = -1, w = 0, per_sm = 0; size_t smem = 0; };
    static Choice choices[64];
    Choice& ch = choices[c->device & 63];
    int& cached_S = ch.S; int& cached_w = ch.w; int& cached_per_sm = ch.per_sm; size_t& cached_smem = ch.smem;
    if (cached_S != c->S) {
        size_t s4 = 0, s7 = 0, s16 = 0;
        const int o4 = batch2_occupancy<R, 4, NP>(c, &s4), o7 = batch2_occupancy<R, 7, NP>(c, &s7), o16 = batch2_occupancy<R, 16, NP>(c, &s16);
        const char* force = getenv("NWB200_BATCH_WARPS");      // developer switch
        const int fw = force ? atoi(force) : 0;
        int best = 0;
        if (o4 > 0 && (fw == 0 || fw == 4)) { best = 4 * o4; cached_w = 4; cached_per_sm = o4; cached_smem = s4; }
        if (o7 > 0 && (fw == 0 || fw == 7) && 7 * o7 > best) { best = 7 * o7; cached_w = 7; cached_per_sm = o7; cached_smem = s7; }
        if (o16 > 0 && (fw == 0 || fw == 16) && 16 * o16 > best) { best = 16 * o16; cached_w = 16; cached_per_sm = o16; cached_smem = s16; }
        if (best == 0) return fail(c, NWB200_ERR_KERNEL_FAILURE, "packed batch kernel does not fit an SM");
        cached_S = c->S;
    }
    if (cached_w == 4) return launch_batch2_w<R, 4, NP>(c, a, cached_smem, cached_per_sm);
    if (cached_w == 7) return launch_batch2_w<R, 7, NP>(c, a, cached_smem, cached_per_sm);
    return launch_batch2_w<R, 16, NP>(c, a, cached_smem, cached_per_sm);
}

// Round-2 packed kernel (nw_batch3.cuh): same CTA-shape search.  G = lanes per packed pair of pairs (32: two pairs per warp, 16: four).
template <int R, int W, int MQ, int K, int G, bool SPLIT, bool EXP, bool PK>
int launch_batch3_w(nwb200_ctx* c, const BatchArgs& a, size_t smem, int per_sm)
{
    constexpr int PPW = 2 * (32 / G);                             // pairs per warp and ticket
    long long grid = (long long)c->sm_count * per_sm;
    const long long need = ((long long)(a.npairs - a.first) + PPW * W - 1) / (PPW * W);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    nw_batch3_kernel<R, W, MQ, K, G, SPLIT, EXP, PK><<<(int)grid, W * 32, smem, c->stream>>>(a);
    c->launches++;
    c->batch_kernel = PK ? "nw_batch3_kernel (5-bit packed letters)" : EXP ? "nw_batch3_kernel (expanded profiles)" : SPLIT ? "nw_batch3_kernel (four pairs per warp, split lanes)" : G == 16 ? "nw_batch3_kernel (four pairs per warp)" : K == 2 ? "nw_batch3_kernel (K = 2)" : (MQ == R / 4 ? "nw_batch3_kernel" : "nw_batch3_kernel (mixed IDP routes)");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, NWB200_ERR_KERNEL_FAILURE, "packed batch kernel launch", e);
    return NWB200_SUCCESS;
}

template <int R, int W, int MQ, int K, int G, bool SPLIT, bool EXP, bool PK>
int batch3_occupancy(nwb200_ctx* c, size_t* smem_out)
{
    const size_t smem = Sched3<R, K, G, SPLIT, EXP>::warp_smem_bytes(c->S) * W;
    *smem_out = smem;
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(nw_batch3_kernel<R, W, MQ, K, G, SPLIT, EXP, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nw_batch3_kernel<R, W, MQ, K, G, SPLIT, EXP, PK>, W * 32, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    return per_sm;
}

template <int R, int MQ, int K, int G = 32, bool SPLIT = false, bool EXP = false, bool PK = false>
int launch_batch3_t(nwb200_ctx* c, const BatchArgs& a)
{
    struct Choice { int S = -1, w = 0, per_sm = 0; size_t smem = 0; };
    static Choice choices[64];                                    // per (R, MQ, K, G) instance and device
    Choice& ch = choices[c->device & 63];
    if (ch.S != c->S) {
        size_t s4 = 0, s8 = 0, s16 = 0;
        const int o4 = batch3_occupancy<R, 4, MQ, K, G, SPLIT, EXP, PK>(c, &s4), o8 = batch3_occupancy<R, 8, MQ, K, G, SPLIT, EXP, PK>(c, &s8), o16 = batch3_occupancy<R, 16, MQ, K, G, SPLIT, EXP, PK>(c, &s16);
        const char* force = getenv("NWB200_BATCH_WARPS");      // developer switch
        const int fw = force ? atoi(force) : 0;
        int best = 0;
        if (o4 > 0 && (fw == 0 || fw == 4)) { best = 4 * o4; ch.w = 4; ch.per_sm = o4; ch.smem = s4; }
        if (o8 > 0 && (fw == 0 || fw == 8) && 8 * o8 > best) { best = 8 * o8; ch.w = 8; ch.per_sm = o8; ch.smem = s8; }
        if (o16 > 0 && (fw == 0 || fw == 16) && 16 * o16 > best) { best = 16 * o16; ch.w = 16; ch.per_sm = o16; ch.smem = s16; }
        if (best == 0) return fail(c, NWB200_ERR_KERNEL_FAILURE, "packed batch kernel does not fit an SM");
        ch.S = c->S;
    }
    if (ch.w == 4) return launch_batch3_w<R, 4, MQ, K, G, SPLIT, EXP, PK>(c, a, ch.smem, ch.per_sm);
    if (ch.w == 8) return launch_batch3_w<R, 8, MQ, K, G, SPLIT, EXP, PK>(c, a, ch.smem, ch.per_sm);
    return launch_batch3_w<R, 16, MQ, K, G, SPLIT, EXP, PK>(c, a, ch.smem, ch.per_sm);
}

bool batch_packed_enabled()
{
    const char* e = getenv("NWB200_BATCH_PACKED");      // developer switch: 0 = always the 32-bit kernel (A/B runs, tests)
    return !(e && e[0] == '0');
}

}  // namespace

int launch_batch(nwb200_ctx* c, const BatchArgs& a)
{
    if (a.npairs <= a.first) return NWB200_SUCCESS;
    if (a.packed5) {
        // 5-bit packed letters are read by the packed-halves kernel only
        if (!(c->max_sprime <= kBatch3MaxSprime && c->S <= kBatch3MaxLetters && c->batch_maxy <= 256))
            return fail(c, NWB200_ERR_INVALID_VALUE, "5-bit packed letters need at most 31 letters, subst - 2*gap <= 127 and pairs of at most 256 rows");
        return launch_batch3_t<8, 2, 1, 32, false, false, true>(c, a);
    }
    // packed halves: P <= min(n, m) * max s' <= 256 * 127 fits 15 bits and 512 * s' fits 16 (nw_batch2.cuh)
    if (c->max_sprime <= kBatch3MaxSprime && c->S <= kBatch3MaxLetters && c->batch_maxy <= 256 && batch_packed_enabled())
    {
        // developer switch for A/B runs: NWB200_BATCH_VARIANT=0 selects the round-1 kernel (nw_batch2.cuh), 2 the mixed-route instance, 3 the K = 2 instance (shuffle off the chain, twice the fill / drain)
        const char* e = getenv("NWB200_BATCH_VARIANT");
        const int v = e ? atoi(e) : 1;
        if (v == 0) return c->batch_maxy <= 128 ? launch_batch2_t<4, 0>(c, a) : launch_batch2_t<8, 4>(c, a);
        if (c->batch_maxy <= 128) return launch_batch3_t<4, 1, 1>(c, a);
        if (v == 2) return launch_batch3_t<8, 1, 1>(c, a);
        if (v == 3) return launch_batch3_t<8, 2, 2>(c, a);
        if (v == 4) return launch_batch3_t<16, 4, 1, 16>(c, a);      // four pairs per warp: 16 lanes x 16 rows per packed pair of pairs
        if (v == 5) return launch_batch3_t<16, 4, 2, 16>(c, a);      // the same with the shuffle off the dependent chain
        if (v == 7) return launch_batch3_t<8, 2, 1, 32, false, true>(c, a);      // profiles two rows per word: the merge is an integer add
        if (v == 6) return launch_batch3_t<16, 4, 1, 16, true>(c, a);      // four pairs per warp, every lane cut into two half-lanes a column apart
        return launch_batch3_t<8, 2, 1>(c, a);
    }
    {
        const char* e = getenv("NWB200_BATCH_TMA");        // developer switch (A/B): column letters staged by TMA bulk copies
        if (e && e[0] == '1') {
            if (c->batch_maxy <= 128) return launch_batch_t<4, true>(c, a);
            if (c->batch_maxy <= 256) return launch_batch_t<8, true>(c, a);
            return launch_batch_t<16, true>(c, a);
        }
    }
    if (c->batch_maxy <= 128) return launch_batch_t<4>(c, a);
    if (c->batch_maxy <= 256) return launch_batch_t<8>(c, a);
    return launch_batch_t<16>(c, a);
}

namespace {

template <int R>
int launch_batch_trace_t(nwb200_ctx* c, const BatchTraceArgs& a0, int need_chunks)
{
    BatchTraceArgs a = a0;
    const size_t base = Sched<R, 1>::warp_smem_bytes(c->S);
    const size_t per_chunk = (size_t)32 * 32 * (R / 4);
    size_t cap = (160u * 1024 - base) / per_chunk;                 // wider pairs take the single-pair path
    if ((size_t)need_chunks < cap) cap = (size_t)need_chunks;
    if (cap < 1) cap = 1;
    a.chunks_cap = (int)cap;
    const size_t smem = base + cap * per_chunk;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(nw_batch_trace_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(c, NWB200_ERR_KERNEL_FAILURE, "cudaFuncSetAttribute(batch trace)", e);
    }
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nw_batch_trace_kernel<R>, 32, smem);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)c->sm_count * per_sm;
    if (grid > (long long)(a.npairs - a.first)) grid = (long long)(a.npairs - a.first);
    if (grid < 1) grid = 1;
    nw_batch_trace_kernel<R><<<(int)grid, 32, smem, c->stream>>>(a);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, NWB200_ERR_KERNEL_FAILURE, "batch trace kernel launch", e);
    return NWB200_SUCCESS;
}


}  // namespace

namespace {
template <int R, bool LOCAL>
int launch_gotoh_t(nwb200_ctx* c, const GotohArgs& g)
{
    constexpr int W = 4;
    const size_t smem = (size_t)W * (size_t)c->S * 32 * R;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(nw_gotoh_batch_kernel<R, LOCAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(c, NWB200_ERR_KERNEL_FAILURE, "cudaFuncSetAttribute(gotoh)", e);
    }
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nw_gotoh_batch_kernel<R, LOCAL>, W * 32, smem);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)c->sm_count * per_sm;
    const long long need = ((long long)(g.b.npairs - g.b.first) + W - 1) / W;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    nw_gotoh_batch_kernel<R, LOCAL><<<(int)grid, W * 32, smem, c->stream>>>(g);
    c->launches++;
    c->batch_kernel = LOCAL ? "nw_gotoh_batch_kernel (local)" : "nw_gotoh_batch_kernel (global)";
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, NWB200_ERR_KERNEL_FAILURE, "gotoh batch kernel launch", e);
    return NWB200_SUCCESS;
}
}  // namespace

// affine-gap / local variants (nw_gotoh.cuh): bands of 256 or 512 rows
int launch_batch_gotoh(nwb200_ctx* c, const GotohArgs& g, bool local)
{
    if (g.b.npairs <= g.b.first) return NWB200_SUCCESS;
    if (c->batch_maxy <= 256) return local ? launch_gotoh_t<8, true>(c, g) : launch_gotoh_t<8, false>(c, g);
    return local ? launch_gotoh_t<16, true>(c, g) : launch_gotoh_t<16, false>(c, g);
}

int launch_batch_trace(nwb200_ctx* c, const BatchTraceArgs& a, int need_chunks)
{
    return c->batch_maxy <= 128 ? launch_batch_trace_t<4>(c, a, need_chunks) : c->batch_maxy <= 256 ? launch_batch_trace_t<8>(c, a, need_chunks)
                                                                                                    : launch_batch_trace_t<16>(c, a, need_chunks);
}

}  // namespace nwb
