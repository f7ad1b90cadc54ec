// nw_engine.cuh -- host-side engine state behind the C ABI (include/nwb200.h).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <utility>
#include <cuda_runtime.h>
#include "../../include/nwb200.h"

namespace nwb {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Geometry {
    int R = 4, W = 4, K = 2;
    int By = 128, Bx = 512;          // band height (32*R), snapshot spacing in columns
    int n = 0, m = 0;
    int nb = 0, pad = 0;             // bands, padding rows above row 1 (rows are aligned to the bottom)
    int tcols = 0;
    int nlc = 0;                     // 32-step chunks per band
    int snap_chunks = 16, nsnap = 0;
    long long ldr = 0;               // header-row pitch in elements
    long long npad = 0;              // nb * By
};

}  // namespace nwb

struct nwb200_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[10] = {};
    // scoring
    std::vector<int32_t> subst;
    int S = 0;
    int gap = 0;
    bool scoring_set = false;
    int max_sprime = 0;
    nwb::DevBuf d_sprime, d_subst;
    // resident pair
    nwb::Geometry g;
    bool pair_resident = false;
    bool headers_valid = false;
    bool fill_done = false;
    nwb::DevBuf d_y, d_HR, d_snap, d_lastcol, d_sync;      // d_y: the row letters, then (at x_off) the column letters
    size_t x_off = 0;
    nwb::PinBuf h_stage, h_small, h_trace;
    // traceback
    nwb::DevBuf d_map, d_MID, d_tmeta, d_ops, d_dense, d_export, d_HR2, d_cut;
    nwb::PinBuf h_export;
    bool trace_done = false;
    bool map_valid = false;          // the last fill launch also produced the origin maps
    bool fuse_map = true;
    bool inline_map = false;
    bool half_map = true;
    bool grouped = true;
    int cluster_max = 16;            // largest thread-block cluster the grouped fill may use (1 = no cluster launch)
    bool map_is_half = false;
    bool edit_cached = false;
    bool moves_on_host = false;      // the move list of the last traceback has been copied into h_trace already
    int* d_timeout_flag = nullptr;   // device address of g_wait_timeout
    std::string last_edit;
    unsigned last_hash = 0;
    // batch
    nwb::DevBuf d_bletters, d_bmeta, d_bscores, d_bticket;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t slice_ev[48] = {};
    cudaStream_t d2h_stream = nullptr;          // scores of a finished slice travel back while later slices are still arriving
    cudaEvent_t slice_done_ev[48] = {};
    nwb::PinBuf h_bscores;
    nwb::DevBuf d_bmoves, d_bmoff, d_bcnt;      // transcripts of a batch: move lists, their offsets, their lengths
    nwb::PinBuf h_bmoves;
    size_t batch_pairs = 0, batch_letters = 0;
    int batch_maxy = 0;
    bool batch_resident = false;
    nwb::PinBuf h_batch;
    // cross-GPU column-block wavefront
    nwb::DevBuf d_order;             // ticket -> (block round, band) in wavefront order
    nwb::DevBuf d_corr;              // corridor passes of the traceback: (first, last) segment per band and level
    nwb::DevBuf d_wave;              // [flags (nq+1)*nb | err, pad | recv (nq+1)*recv_stride] -- ONE allocation, exported over CUDA IPC
    void* wave_peer_base = nullptr;  // the right neighbour's d_wave mapped into this process (nullptr: loopback)
    int wave_rank = 0, wave_world = 1, wave_wc = 0, wave_nq = 0, wave_nblocks = 0;
    long long wave_ldr = 0, wave_hr_stride = 0, wave_recv_stride = 0;
    bool wave_ready = false, wave_connected = false, wave_filled = false;
    // traceback after a cross-GPU fill: header rows and snapshots in the layout of the whole matrix on every rank (its own column blocks
    // filled in), the other ranks' buffers mapped on the rank that walks the path
    bool wave_keep = false, wave_global = false;
    int plan_world = 1;                 // ranks the pair being uploaded is spread over (a hint for plan_geometry)
    int scan_warps = 16;                // strips per group of the prefix-max scorer (nw_scan.cuh)
    void* wave_peer_hr[16] = {};
    void* wave_peer_snap[16] = {};
    // row-parallel prefix-max scorer
    int scan_total = 0, scan_chunk0 = 0, scan_nchunks = 0, scan_per = 0;
    unsigned scan_epoch = 0;
    bool scan_ready = false, scan_filled = false;
    // developer aids
    nwb::DevBuf d_dbg;
    bool dbg_stamps = false;
    int dbg_mode = 0;
    // bookkeeping
    nwb200_timing timing = {};
    unsigned epoch = 0;
    int launches = 0;
    const char* batch_kernel = "";
    bool batch_packed5 = false;          // the resident batch holds 5-bit packed letters
    int corr_w = 0, corr_nseg = 0;      // the last traceback: segments per band of the corridor pass (0: none) out of corr_nseg
    int* d_miss = nullptr;              // its miss flags (device)
    int corr_levels = 0, corr_w_lv[2] = {0, 0}, corr_d_lv[2] = {0, 0};      // the corridor passes of the last traceback (narrow, wide)
    // resources of the last fill launch (nwb200_get_memory_usage; the reference's updateNwAlgPeakMemUsage, nwalign_shared.cpp:5-25)
    int fill_regs = 0, fill_threads = 0, fill_blocks = 0;
    size_t fill_smem_static = 0, fill_smem_dynamic = 0, fill_local = 0;
    cudaError_t last_cuda = cudaSuccess;
    std::string last_error;
};

// ---- host helpers shared by the translation units of libnwb200.so
namespace nwb {

inline int fail(nwb200_ctx* c, int stat, const char* msg, cudaError_t e = cudaSuccess)
{
    if (c) {
        c->last_error = msg;
        if (e != cudaSuccess) { c->last_cuda = e; c->last_error += std::string(": ") + cudaGetErrorString(e); }
    }
    return stat;
}

#define CU(c, call, stat, msg) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return nwb::fail((c), (stat), (msg), e__); } while (0)

inline float ev_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0.f; cudaEventElapsedTime(&ms, a, b); return ms; }

// What one launch of a batch kernel works on (nw_batch.cuh / nw_batch2.cuh; launched from nw_batch_launch.cu).
struct BatchArgs {
    const uint8_t* letters;              // byte letters of all sequences
    const unsigned long long* offY;      // per pair: offset of the row sequence
    const unsigned* lenY;
    const unsigned long long* offX;
    const unsigned* lenX;
    unsigned long long first;            // this launch aligns pairs [first, npairs)
    unsigned long long npairs;
    const uint8_t* sprime;
    int S;
    int gap;
    int* scores;                         // H[lenY][lenX] per pair; kBatchTooTall if lenY > 32*R (the host re-runs those as single pairs)
    unsigned long long* ticket;          // zero at launch
    int packed5;                         // 1: the letters are 5-bit packed (8 letters in 5 bytes), offsets are BYTE offsets of the sequences' bit streams
    int* err;                            // set to 1 when a letter outside the alphabet is met (the pair's score is then meaningless), 2 when the
                                         // s' table did not arrive
};
constexpr int kBatchTooTall = (int)0x80000000;

struct BatchTraceArgs {
    const uint8_t* letters;
    const unsigned long long* offY;
    const unsigned* lenY;
    const unsigned long long* offX;
    const unsigned* lenX;
    unsigned long long first, npairs;    // this launch walks pairs [first, npairs)
    const uint8_t* sprime;
    int S;
    int negg;                            // -gap
    const unsigned long long* moff;      // [npairs - first]: start of pair (first + q)'s move list in `moves` (lenY + lenX bytes each)
    unsigned char* moves;
    int* cnt;                            // [npairs - first]: moves emitted; -1 = not handled here (empty sequence, taller than one band, or
                                         // more columns than the CTA's shared memory holds codes for): the host takes the single-pair path
    int chunks_cap;                      // 32-column chunks of move codes that fit behind the warp's sweep buffers
};

// nw_batch_launch.cu: picks the kernel instance for the resident batch (c->batch_maxy, c->S, c->max_sprime) and launches it on c->stream
int launch_batch(nwb200_ctx* c, const BatchArgs& a);
int launch_batch_trace(nwb200_ctx* c, const BatchTraceArgs& a, int need_chunks);
struct GotohArgs;
int launch_batch_gotoh(nwb200_ctx* c, const GotohArgs& g, bool local);

}  // namespace nwb

