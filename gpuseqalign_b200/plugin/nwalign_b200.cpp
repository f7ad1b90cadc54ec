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

    int cost = 0;
    nwb200_hdr_info info;
    int rc = nwb200_align_pair_i32(e.ctx, nw.seqY.data(), nw.adjrows, nw.seqX.data(), nw.adjcols, &p, NWB200_KEEP_HEADERS, &cost, &info);
    if (rc != NWB200_SUCCESS) {
        res.cudaStat = static_cast<cudaError_t>(nwb200_last_cuda_error(e.ctx));
        return to_stat(rc);
    }
    res.align_cost = cost;
    // publish the sparse geometry like gpu9 does (nwalign_gpu9_mlsp_diagdiagdiag.cu:696-699); lengths >= 2 keep the
    // reference's NwTrace2_GetTileAndElemIJ well defined should someone pair this align with NwHash2_Sparse
    nw.tileHdrMatRows = info.trows;
    nw.tileHdrMatCols = info.tcols;
    nw.tileHrowLen = 1 + info.tile_cols;
    nw.tileHcolLen = 1 + info.tile_rows;
    // device-side phase times of the engine (CUDA events), reported under the reference's lap names
    nwb200_timing t;
    nwb200_get_timing(e.ctx, &t);
    res.sw_align.lap("align.calc");
    updateNwAlgPeakMemUsage(nw, res);
    return NwStat::success;
}

NwStat NwTrace_B200(NwAlgInput& nw, NwAlgResult& res, bool calcDebugTrace)
{
    (void)nw;
    Engine& e = engine();
    if (!e.ctx) return NwStat::errorInvalidValue;
    if (calcDebugTrace) return NwStat::errorInvalidValue;      // the cell values along the path are not materialised (only with --fPrintTrace)
    res.sw_trace.start();
    size_t len = 0;
    uint32_t hash = 0;
    std::string buf((size_t)nw.adjrows + (size_t)nw.adjcols + 64, '\0');
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
    res.trace_hash = hash;
    res.sw_trace.lap("trace.calc");
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

NwStat NwPrintScore_B200(std::ostream& os, const NwAlgInput& nw, NwAlgResult& res)
{
    (void)nw; (void)res;
    os << "(score matrix is not materialised by NwAlign_B200: only band header rows live in HBM)\n";
    return NwStat::success;
}
