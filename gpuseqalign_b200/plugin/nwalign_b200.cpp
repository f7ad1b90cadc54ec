// nwalign_b200.cpp -- the B200 engine as ONE MORE ENTRY of the reference's algorithm registry.
//
// This file is compiled against the reference's own headers (nw_algorithm.hpp, run_types.hpp, nw_fns.hpp,
// stopwatch.hpp -- found under $(REF)/src at build time, never copied) and adapts the C ABI of include/nwb200.h
// to the reference's plug-in signatures (nw_algorithm.hpp:11-15):
//     NwStat align(const NwAlgParams&, NwAlgInput&, NwAlgResult&)       <- NwAlign_B200
//     NwStat trace(NwAlgInput&, NwAlgResult&, bool calcDebugTrace)        <- NwTrace_B200
//     NwStat hash (NwAlgInput&, NwAlgResult&)                             <- NwHash_B200
// so that benchmark.cpp:473,484,488 call it exactly like NwAlign_Gpu9_Mlsp_DiagDiagDiag / NwTrace2_Sparse /
// NwHash2_Sparse.  Contract kept: inputs are nw.seqY / nw.seqX (int vectors, dummy element 0),
// nw.subst / nw.substsz / nw.gapoCost; outputs are res.align_cost, res.edit_trace, res.trace_hash, res.score_hash,
// the Stopwatch laps with the TSV column names (file_formats.cpp:505-518), res.cudaStat on CUDA failure; every
// failure is an NwStat, nothing throws.  Parameters (all optional in the param JSON, 0 = engine default):
//     "rowsPerLane" (4, 8, 16), "warpsPerBlock" (1, 4), "tileCols" (snapshot spacing, multiple of 32), "skew" (1, 2).
#include "nw_algorithm.hpp"
#include "nw_fns.hpp"
#include "nwalign_shared.hpp"
#include "nwb200.h"

#include "fmt_guard.hpp"

#include <algorithm>
#include <iomanip>
#include <string>
#include <vector>

namespace {

struct Engine {
    nwb200_ctx* ctx = nullptr;
    std::vector<int> subst;
    int substsz = 0;
    int gap = 0;
    bool scoring_set = false;
    ~Engine() { if (ctx) nwb200_destroy(ctx); }
};

Engine& engine()
{
    static Engine e;      // one context for the process: the reference is single-threaded, device 0 (benchmark.cpp:179)
    return e;
}

NwStat to_stat(int rc) { return static_cast<NwStat>(rc); }

int param_or_zero(const NwAlgParams& pr, const char* name)
{
    try { return pr.at(name).curr(); }
    catch (const std::exception&) { return 0; }
}

}  // namespace

NwStat NwAlign_B200(const NwAlgParams& pr, NwAlgInput& nw, NwAlgResult& res)
{
    Engine& e = engine();
    res.sw_align.start();
    if (!e.ctx) {
        int rc = nwb200_create(&e.ctx, 0);
        if (rc != NWB200_SUCCESS) { e.ctx = nullptr; return to_stat(rc); }
    }
    if (!e.scoring_set || e.substsz != nw.substsz || e.gap != nw.gapoCost || e.subst != nw.subst) {
        int rc = nwb200_set_scoring(e.ctx, nw.subst.data(), nw.substsz, nw.gapoCost);
        if (rc != NWB200_SUCCESS) { res.cudaStat = static_cast<cudaError_t>(nwb200_last_cuda_error(e.ctx)); return to_stat(rc); }
        e.subst = nw.subst; e.substsz = nw.substsz; e.gap = nw.gapoCost; e.scoring_set = true;
    }
    res.sw_align.lap("align.alloc");

    nwb200_params p;
    p.rows_per_lane = param_or_zero(pr, "rowsPerLane");
    p.warps_per_block = param_or_zero(pr, "warpsPerBlock");
    p.tile_cols = param_or_zero(pr, "tileCols");
    p.reserved = param_or_zero(pr, "skew");

    // The phases are separate C-ABI calls, each ending in a synchronisation, so that the Stopwatch laps are the reference's own
    // (nwalign_gpu9_mlsp_diagdiagdiag.cu:481 "align.cpy_dev", :687 "align.calc", :711 "align.cpy_host"; file_formats.cpp:505-510).
    auto failed = [&](int rc) {
        res.cudaStat = static_cast<cudaError_t>(nwb200_last_cuda_error(e.ctx));
        return to_stat(rc);
    };
    int rc = nwb200_upload_pair_i32(e.ctx, nw.seqY.data(), nw.adjrows, nw.seqX.data(), nw.adjcols, &p);
    if (rc == NWB200_SUCCESS) rc = nwb200_sync(e.ctx);
    if (rc != NWB200_SUCCESS) return failed(rc);
    res.sw_align.lap("align.cpy_dev");

    rc = nwb200_fill_resident(e.ctx, NWB200_KEEP_HEADERS);
    if (rc == NWB200_SUCCESS) rc = nwb200_sync(e.ctx);
    if (rc != NWB200_SUCCESS) return failed(rc);
    res.sw_align.lap("align.calc");

    int cost = 0;
    rc = nwb200_fetch_score(e.ctx, &cost);
    if (rc != NWB200_SUCCESS) return failed(rc);
    res.align_cost = cost;
    res.sw_align.lap("align.cpy_host");

    // publish the sparse geometry like gpu9 does (nwalign_gpu9_mlsp_diagdiagdiag.cu:696-699); lengths >= 2 keep the
    // reference's NwTrace2_GetTileAndElemIJ well defined should someone pair this align with NwHash2_Sparse
    nwb200_hdr_info info;
    if (nwb200_get_hdr_info(e.ctx, &info) == NWB200_SUCCESS) {
        nw.tileHdrMatRows = info.trows;
        nw.tileHdrMatCols = info.tcols;
        nw.tileHrowLen = 1 + info.tile_cols;
        nw.tileHcolLen = 1 + info.tile_rows;
    }
    // peak memory: host allocations of NwAlgInput as every algorithm reports them, and the engine's own device buffers and launch
    // resources in the places where the reference's kernels report theirs (nwalign_shared.cpp:16-24; the engine's buffers live
    // in its context, not in NwAlgInput, so measureDeviceAllocations() does not see them)
    updateNwAlgPeakMemUsage(nw, res);
    nwb200_mem_usage mu;
    if (nwb200_get_memory_usage(e.ctx, &mu) == NWB200_SUCCESS) {
        res.ramPeakAllocs = std::max<size_t>(res.ramPeakAllocs, nw.measureHostAllocations() + (size_t)mu.pinned_host_bytes);
        res.globalMemPeakAllocs = std::max<size_t>(res.globalMemPeakAllocs, nw.measureDeviceAllocations() + (size_t)mu.device_bytes);
        res.sharedMemPeakAllocs = std::max<size_t>(res.sharedMemPeakAllocs, (size_t)mu.shared_bytes);
        res.localMemPeakAllocs = std::max<size_t>(res.localMemPeakAllocs, (size_t)mu.local_bytes);
        res.regMemPeakAllocs = std::max<size_t>(res.regMemPeakAllocs, (size_t)mu.register_bytes);
    }
    return NwStat::success;
}

NwStat NwTrace_B200(NwAlgInput& nw, NwAlgResult& res, bool calcDebugTrace)
{
    Engine& e = engine();
    if (!e.ctx) return NwStat::errorInvalidValue;
    res.sw_trace.start();
    size_t len = 0;
    uint32_t hash = 0;
    std::string buf;
    try {
        buf.assign((size_t)nw.adjrows + (size_t)nw.adjcols + 64, '\0');
        if (calcDebugTrace) nw.trace.reserve((size_t)nw.adjrows - 1 + (size_t)nw.adjcols);      // nwtrace1_plain.cpp:15-18
    } catch (const std::exception&) {
        return NwStat::errorMemoryAllocation;
    }
    updateNwAlgPeakMemUsage(nw, res);
    res.sw_trace.lap("trace.alloc");
    int rc = nwb200_trace_pair(e.ctx, &buf[0], buf.size(), &len, &hash);
    if (rc == NWB200_ERR_INVALID_VALUE && len > buf.size()) {
        buf.resize(len);
        rc = nwb200_trace_pair(e.ctx, &buf[0], buf.size(), &len, &hash);
    }
    if (rc != NWB200_SUCCESS) {
        res.cudaStat = static_cast<cudaError_t>(nwb200_last_cuda_error(e.ctx));
        return to_stat(rc);
    }
    buf.resize(len);
    res.edit_trace = std::move(buf);
    res.sw_trace.lap("trace.calc");
    if (calcDebugTrace) {
        // the score-matrix values along the path, top-left -> bottom-right, folded into the hash behind the transcript
        // (nwtrace1_plain.cpp:34-38,107,120-126)
        size_t cnt = 0;
        nwb200_trace_values(e.ctx, nullptr, 0, &cnt);
        try { nw.trace.assign(cnt, 0); } catch (const std::exception&) { return NwStat::errorMemoryAllocation; }
        rc = nwb200_trace_values(e.ctx, nw.trace.data(), nw.trace.size(), &cnt);
        if (rc != NWB200_SUCCESS) {
            res.cudaStat = static_cast<cudaError_t>(nwb200_last_cuda_error(e.ctx));
            return to_stat(rc);
        }
        for (auto& curr : nw.trace) hash = ((hash << 5) + hash) ^ (unsigned)curr;
    }
    res.trace_hash = hash;
    return NwStat::success;
}

NwStat NwHash_B200(NwAlgInput& nw, NwAlgResult& res)
{
    (void)nw;
    Engine& e = engine();
    if (!e.ctx) return NwStat::errorInvalidValue;
    res.sw_hash.start();
    uint32_t h = 0;
    int rc = nwb200_score_hash(e.ctx, &h);
    if (rc != NWB200_SUCCESS) {
        res.cudaStat = static_cast<cudaError_t>(nwb200_last_cuda_error(e.ctx));
        return to_stat(rc);
    }
    res.score_hash = h;
    res.sw_hash.lap("hash.calc");
    return NwStat::success;
}

// The score matrix in the reference's text form (print_mat.hpp / NwPrintScore2_Sparse, nwtrace2_sparse.cpp:346-419: every element
// setw(4) and a comma, one matrix row per line); the rows are recomputed on the GPU a block at a time.
NwStat NwPrintScore_B200(std::ostream& os, const NwAlgInput& nw, NwAlgResult& res)
{
    Engine& e = engine();
    if (!e.ctx) return NwStat::errorInvalidValue;
    FormatFlagsGuard fg {os, 4};
    const long long rows = nw.adjrows, cols = nw.adjcols;
    const long long block = std::max<long long>(1, (long long)(16 << 20) / std::max<long long>(cols, 1));      // ~64 MB of ints at a time
    std::vector<int> buf;
    try { buf.resize((size_t)std::min(block, rows) * (size_t)cols); } catch (const std::exception&) { return NwStat::errorMemoryAllocation; }
    updateNwAlgPeakMemUsage(nw, res);
    for (long long r0 = 0; r0 < rows; r0 += block) {
        const long long k = std::min(block, rows - r0);
        int rc = nwb200_score_rows(e.ctx, r0, k, buf.data());
        if (rc != NWB200_SUCCESS) {
            res.cudaStat = static_cast<cudaError_t>(nwb200_last_cuda_error(e.ctx));
            return to_stat(rc);
        }
        for (long long i = 0; i < k; i++) {
            for (long long j = 0; j < cols; j++) os << std::setw(4) << buf[(size_t)(i * cols + j)] << ',';
            os << '\n';
        }
    }
    return NwStat::success;
}
