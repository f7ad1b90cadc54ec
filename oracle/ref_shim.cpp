// ref_shim.cpp -- C-ABI harness around the UNMODIFIED reference implementation.
//
// TEST INFRASTRUCTURE ONLY (same rule as nw_oracle.c).  This file is our own code; it is
// compiled together with the reference's sources *where they lie* under /root/reference/src
// (see oracle/Makefile) into oracle/_ref/libnwref.so.  No reference source is copied into
// the repository.  It fills NwAlgInput exactly the way benchmark.cpp:175-223,411-426 does
// and then calls the reference's own entry points:
//   cpu1  NwAlign_Cpu1_St_Row     + NwHash1_Plain  + NwTrace1_Plain   (nw_algorithm.cpp:52)
//   cpu4  NwAlign_Cpu4_Mt_DiagRow + NwHash1_Plain  + NwTrace1_Plain   (nw_algorithm.cpp:55)
//   gpu9  NwAlign_Gpu9_Mlsp_DiagDiagDiag + NwHash2_Sparse + NwTrace2_Sparse (nw_algorithm.cpp:64; needs a GPU)
//   sparse-from-headers: NwTrace2_Sparse on caller-provided tile headers (validates the
//         restated header producer and the engine's header layout against the reference consumer)
#include "nw_fns.hpp"
#include "run_types.hpp"
#include <cstring>
#include <cuda_runtime.h>
#include <string>

extern "C" {

struct nwref_result
{
    int stat;            // NwStat of the failing step, 0 on success
    int step;            // 1 align, 2 hash, 3 trace (0 = none failed)
    int cuda_stat;
    int align_cost;
    unsigned score_hash;
    unsigned trace_hash;
    unsigned long long edit_len;
    float ms_align_alloc, ms_align_cpy_dev, ms_align_init_hdr, ms_align_calc, ms_align_cpy_host;
    float ms_hash_calc, ms_trace_alloc, ms_trace_calc;
    int tile_hdr_rows, tile_hdr_cols, tile_hrow_len, tile_hcol_len;
};

static void nwref_fill_input(NwAlgInput& nw, const int* seqY, int adjrows, const int* seqX, int adjcols,
                             const int* subst, int substsz, int gap)
{
    nw.subst.assign(subst, subst + (size_t)substsz * substsz);
    nw.substsz = substsz;
    nw.gapoCost = gap;
    nw.seqY.assign(seqY, seqY + adjrows);
    nw.seqX.assign(seqX, seqX + adjcols);
    nw.adjrows = adjrows;
    nw.adjcols = adjcols;
    nw.tileHdrMatRows = nw.tileHdrMatCols = nw.tileHrowLen = nw.tileHcolLen = 0;
    nw.sm_count = 1;
    nw.warpsz = 32;
    nw.maxThreadsPerBlock = 1024;
}

static void nwref_collect(const NwAlgResult& res, nwref_result* out, char* edit_buf, size_t cap)
{
    out->align_cost = res.align_cost;
    out->score_hash = res.score_hash;
    out->trace_hash = res.trace_hash;
    out->cuda_stat = (int)res.cudaStat;
    out->edit_len = res.edit_trace.size();
    if (edit_buf && cap > 0) {
        size_t n = res.edit_trace.size() < cap ? res.edit_trace.size() : cap;
        memcpy(edit_buf, res.edit_trace.data(), n);
    }
    out->ms_align_alloc = res.sw_align.get_or_default("align.alloc");
    out->ms_align_cpy_dev = res.sw_align.get_or_default("align.cpy_dev");
    out->ms_align_init_hdr = res.sw_align.get_or_default("align.init_hdr");
    out->ms_align_calc = res.sw_align.get_or_default("align.calc");
    out->ms_align_cpy_host = res.sw_align.get_or_default("align.cpy_host");
    out->ms_hash_calc = res.sw_hash.get_or_default("hash.calc");
    out->ms_trace_alloc = res.sw_trace.get_or_default("trace.alloc");
    out->ms_trace_calc = res.sw_trace.get_or_default("trace.calc");
}

// alg: 1 = cpu1, 4 = cpu4, 9 = gpu9.  params: cpu4 {blocksz}; gpu9 {threadsPerBlockA, subtileRows, subtileCols, subtileBx}.
int nwref_run(int alg, const int* seqY, int adjrows, const int* seqX, int adjcols,
              const int* subst, int substsz, int gap, const int* params, int n_params,
              int do_hash, int do_trace, char* edit_buf, unsigned long long edit_cap, nwref_result* out)
{
    memset(out, 0, sizeof(*out));
    try {
        NwAlgInput nw {};
        NwAlgResult res {};
        res.cudaStat = cudaSuccess;
        nwref_fill_input(nw, seqY, adjrows, seqX, adjcols, subst, substsz, gap);

        Dict<std::string, NwAlgParam> pd;
        NwStat (*alignFn)(const NwAlgParams&, NwAlgInput&, NwAlgResult&) = nullptr;
        NwStat (*hashFn)(NwAlgInput&, NwAlgResult&) = NwHash1_Plain;
        NwStat (*traceFn)(NwAlgInput&, NwAlgResult&, bool) = NwTrace1_Plain;
        if (alg == 1) {
            alignFn = NwAlign_Cpu1_St_Row;
        } else if (alg == 4) {
            alignFn = NwAlign_Cpu4_Mt_DiagRow;
            pd.insert("blocksz", NwAlgParam({n_params > 0 ? params[0] : 256}));
        } else if (alg == 9) {
#ifdef NWREF_WITH_GPU
            cudaDeviceProp prop {};
            if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { out->stat = (int)NwStat::errorCudaGeneral; out->step = 1; return 1; }
            nw.sm_count = prop.multiProcessorCount;
            nw.warpsz = prop.warpSize;
            nw.maxThreadsPerBlock = prop.maxThreadsPerBlock;
            nw.subst_gpu.init((size_t)substsz * substsz);
            if (cudaMemcpy(nw.subst_gpu.data(), nw.subst.data(), sizeof(int) * (size_t)substsz * substsz, cudaMemcpyHostToDevice) != cudaSuccess) {
                out->stat = (int)NwStat::errorMemoryTransfer; out->step = 1; return 1;
            }
            alignFn = NwAlign_Gpu9_Mlsp_DiagDiagDiag;
            hashFn = NwHash2_Sparse;
            traceFn = NwTrace2_Sparse;
            pd.insert("threadsPerBlockA", NwAlgParam({n_params > 0 ? params[0] : 128}));
            pd.insert("subtileRows", NwAlgParam({n_params > 1 ? params[1] : 4}));
            pd.insert("subtileCols", NwAlgParam({n_params > 2 ? params[2] : 4}));
            pd.insert("subtileBx", NwAlgParam({n_params > 3 ? params[3] : 48}));
#else
            out->stat = (int)NwStat::errorInvalidValue; out->step = 1; return 1;
#endif
        } else {
            out->stat = (int)NwStat::errorInvalidValue; out->step = 1; return 1;
        }
        NwAlgParams pr(pd);

        NwStat st = alignFn(pr, nw, res);
        if (st != NwStat::success) { out->stat = (int)st; out->step = 1; out->cuda_stat = (int)res.cudaStat; return 1; }
        if (do_hash) {
            st = hashFn(nw, res);
            if (st != NwStat::success) { out->stat = (int)st; out->step = 2; return 1; }
        }
        if (do_trace) {
            st = traceFn(nw, res, false);
            if (st != NwStat::success) { out->stat = (int)st; out->step = 3; return 1; }
        }
        nwref_collect(res, out, edit_buf, (size_t)edit_cap);
        out->tile_hdr_rows = nw.tileHdrMatRows; out->tile_hdr_cols = nw.tileHdrMatCols;
        out->tile_hrow_len = nw.tileHrowLen; out->tile_hcol_len = nw.tileHcolLen;
        return 0;
    } catch (...) {
        out->stat = (int)NwStat::errorMemoryAllocation; out->step = 1;
        return 1;
    }
}

// Run the reference's NwTrace2_Sparse (+ optionally NwHash2_Sparse) on caller-provided headers
// in the App. A-4 layout with tile sizes By x Bx; align_cost is recomputed from the last tile
// exactly as nwalign_gpu9_mlsp_diagdiagdiag.cu:713-716 does.
int nwref_trace_from_headers(const int* seqY, int adjrows, const int* seqX, int adjcols,
                             const int* subst, int substsz, int gap,
                             const int* hrow, const int* hcol, int By, int Bx, int do_hash,
                             char* edit_buf, unsigned long long edit_cap, nwref_result* out)
{
    memset(out, 0, sizeof(*out));
    try {
        NwAlgInput nw {};
        NwAlgResult res {};
        res.cudaStat = cudaSuccess;
        nwref_fill_input(nw, seqY, adjrows, seqX, adjcols, subst, substsz, gap);
        int trows = (adjrows - 1 + By - 1) / By; if (trows < 1) trows = 1;
        int tcols = (adjcols - 1 + Bx - 1) / Bx; if (tcols < 1) tcols = 1;
        nw.tileHdrMatRows = trows; nw.tileHdrMatCols = tcols;
        nw.tileHrowLen = 1 + Bx; nw.tileHcolLen = 1 + By;
        size_t nrow = (size_t)trows * tcols * (1 + Bx), ncol = (size_t)trows * tcols * (1 + By);
        nw.tileHrowMat.init(nrow); nw.tileHcolMat.init(ncol);
        memcpy(nw.tileHrowMat.data(), hrow, nrow * sizeof(int));
        memcpy(nw.tileHcolMat.data(), hcol, ncol * sizeof(int));
        std::vector<int> tile((size_t)(1 + By) * (1 + Bx), 0);
        std::swap(nw.tile, tile);
        TileAndElemIJ co;
        NwTrace2_GetTileAndElemIJ(nw, nw.adjrows - 1, nw.adjcols - 1, co);
        NwTrace2_AlignTile(nw.tile, nw, co);
        res.align_cost = el(nw.tile, 1 + Bx, co.iTileElem, co.jTileElem);
        NwStat st;
        if (do_hash) {
            st = NwHash2_Sparse(nw, res);
            if (st != NwStat::success) { out->stat = (int)st; out->step = 2; return 1; }
        }
        st = NwTrace2_Sparse(nw, res, false);
        if (st != NwStat::success) { out->stat = (int)st; out->step = 3; return 1; }
        nwref_collect(res, out, edit_buf, (size_t)edit_cap);
        return 0;
    } catch (...) {
        out->stat = (int)NwStat::errorMemoryAllocation; out->step = 1;
        return 1;
    }
}

// cfg3 on the host, the way benchmark.cpp:406-517 would run it: the reference's own cpu4 (or cpu1) align called ONCE PER
// PAIR on a freshly filled NwAlgInput (benchmark.cpp:411-426, reset after every pair :436-439).  A 256 x 256 pair is a single
// cpu4 tile (blocksz 256), so the reference has no parallelism inside a pair; to give it every host core
// the PAIRS are dealt to `threads` OpenMP threads here (our loop, the reference's functions; its inner `omp parallel` then runs
// with one thread because nested parallelism is off by default).  threads <= 1: the strictly serial loop of benchmark.cpp.
// Returns 0 and the scores; *ms_total = wall time of the loop.
int nwref_batch_cpu(int alg, const unsigned char* letters, const unsigned long long* offY, const unsigned* lenY,
                    const unsigned long long* offX, const unsigned* lenX, unsigned long long n_pairs,
                    const int* subst, int substsz, int gap, int blocksz, int threads, int* scores, double* ms_total)
{
    if (alg != 1 && alg != 4) return 1;
    int bad = 0;
    Stopwatch sw {};
    sw.start();
    #pragma omp parallel for schedule(dynamic, 8) num_threads(threads > 1 ? threads : 1) reduction(| : bad)
    for (long long p = 0; p < (long long)n_pairs; p++) {
        try {
            NwAlgInput nw {};
            NwAlgResult res {};
            res.cudaStat = cudaSuccess;
            std::vector<int> sy((size_t)lenY[p] + 1, 0), sx((size_t)lenX[p] + 1, 0);
            for (unsigned i = 0; i < lenY[p]; i++) sy[i + 1] = letters[offY[p] + i];
            for (unsigned j = 0; j < lenX[p]; j++) sx[j + 1] = letters[offX[p] + j];
            nwref_fill_input(nw, sy.data(), (int)sy.size(), sx.data(), (int)sx.size(), subst, substsz, gap);
            Dict<std::string, NwAlgParam> pd;
            if (alg == 4) pd.insert("blocksz", NwAlgParam({blocksz > 0 ? blocksz : 256}));
            NwAlgParams pr(pd);
            NwStat st = (alg == 4) ? NwAlign_Cpu4_Mt_DiagRow(pr, nw, res) : NwAlign_Cpu1_St_Row(pr, nw, res);
            if (st != NwStat::success) bad |= 1;
            scores[p] = res.align_cost;
        } catch (...) {
            bad |= 1;
        }
    }
    sw.lap("batch");
    if (ms_total) *ms_total = sw.get_or_default("batch");
    return bad;
}

int nwref_has_gpu9(void)
{
#ifdef NWREF_WITH_GPU
    return 1;
#else
    return 0;
#endif
}

} // extern "C"
