/*
 * nwb200.h -- C ABI of the B200-native Needleman-Wunsch (linear gap) engine.
 *
 * This is the drop-in boundary for the hot path of markods/GpuSeqAlign: everything the
 * reference's algorithm plug-ins do between benchmark.cpp:473 (alg.align), :484 (alg.hash)
 * and :488 (alg.trace).  The C++ adaptor in gpuseqalign_b200/plugin/nwalign_b200.cpp wraps
 * these calls in the reference's NwAlignFn / NwTraceFn / NwHashFn signatures
 * (nw_algorithm.hpp:11-13) so the engine registers as one more entry of
 * getNwAlgorithmMap (nw_algorithm.cpp:48-68); INTEGRATION.md shows the binding.
 *
 * Conventions (identical to the reference's data contract, run_types.hpp:70-110):
 *   - sequences are arrays of letter indices; the *_i32 entry points take the reference's
 *     `int` vectors with a dummy element 0 in front (adjrows = lenY + 1, adjcols = lenX + 1,
 *     file_formats.cpp:43-47, benchmark.cpp:411-426); the *_u8 / batch entry points take
 *     plain byte letters without the dummy element;
 *   - seqY indexes matrix rows, seqX matrix columns; subst[y * substsz + x] (benchmark.cpp:192);
 *   - gap is the (normally negative) linear gap score nw.gapoCost (cmd_parser.cpp:302);
 *   - every function returns an NwStat value (run_types.hpp:12-24): 0 = success,
 *     2 errorCudaGeneral, 3 errorMemoryAllocation, 4 errorMemoryTransfer, 5 errorKernelFailure,
 *     8 errorInvalidValue, 9 errorInvalidResult.  Nothing throws across this boundary.
 *   - a context is not thread-safe; it owns one CUDA device, one stream and all device buffers
 *     (the reference instead re-allocates inside every timed align, nwalign_gpu9...cu:438-449).
 *
 * Results are bit-exact with the reference: integer score, score hash, trace hash and edit
 * transcript (nwtrace1_plain.cpp / nwtrace2_sparse.cpp semantics).  There is no CPU fallback:
 * without a CUDA device nwb200_create fails with errorCudaGeneral.
 */
#ifndef NWB200_H
#define NWB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NWB200_API __attribute__((visibility("default")))
#else
#define NWB200_API
#endif

/* NwStat values (run_types.hpp:12-24). */
enum {
    NWB200_SUCCESS = 0,
    NWB200_ERR_CUDA_GENERAL = 2,
    NWB200_ERR_MEMORY_ALLOCATION = 3,
    NWB200_ERR_MEMORY_TRANSFER = 4,
    NWB200_ERR_KERNEL_FAILURE = 5,
    NWB200_ERR_INVALID_VALUE = 8,
    NWB200_ERR_INVALID_RESULT = 9
};

typedef struct nwb200_ctx nwb200_ctx;

/* Tile geometry: the analogue of the reference's per-algorithm parameters
 * (param_best.json: subtileRows/subtileCols/subtileBx for gpu9; nwalign_gpu9...cu:384-398).
 * tile_rows is fixed by the kernel shape (rows per lane x 32 lanes); tile_cols is the spacing of the
 * register snapshots that play the role of the reference's header columns.  0 = engine default. */
typedef struct nwb200_params {
    int32_t rows_per_lane;   /* 4, 8 or 16: rows per lane, band height = 32 x this (0: 4 up to 75 776 rows, else 8) */
    int32_t warps_per_block; /* 1 or 4 warps per CTA of the fill launch (0: 4)                                       */
    int32_t tile_cols;       /* snapshot spacing Bx in columns, multiple of 32, <= 1024 (0: 512)                      */
    int32_t reserved;        /* lane skew K: 1 or 2 (0: 2)                                                            */
} nwb200_params;

/* Flags for nwb200_align_pair_*. */
#define NWB200_SCORE_ONLY   0x0   /* keep only what the score needs                        */
#define NWB200_KEEP_HEADERS 0x1   /* keep tile header rows/columns for trace / copy_headers */
#define NWB200_WITH_TRACE   0x2   /* (implies KEEP_HEADERS) the align call also enqueues the traceback kernels and the copy of the move
                                     list behind the fill, before its single synchronisation: the trace call that follows only
                                     formats the transcript.  Same results; saves a host round trip per pair. */

/* Geometry of the sparse score-matrix representation kept on the device after an align
 * (what the reference publishes in nw.tileHdrMatRows/Cols, tileHrowLen/tileHcolLen,
 * nwalign_gpu9...cu:696-699). */
typedef struct nwb200_hdr_info {
    int32_t tile_rows;     /* By */
    int32_t tile_cols;     /* Bx */
    int32_t trows;         /* tileHdrMatRows */
    int32_t tcols;         /* tileHdrMatCols */
    int64_t hrow_elems;    /* trows*tcols*(1+Bx) ints in the reference layout */
    int64_t hcol_elems;    /* trows*tcols*(1+By) */
} nwb200_hdr_info;

/* Per-phase device/host times of the last call, in ms (the reference's Stopwatch lap names,
 * file_formats.cpp:505-518). */
typedef struct nwb200_timing {
    float align_cpy_dev;   /* H2D of the sequences                     */
    float align_calc;      /* fill kernel(s), CUDA events              */
    float align_cpy_host;  /* D2H of the score (and headers if copied) */
    float trace_calc;      /* traceback kernels, CUDA events           */
    float trace_cpy_host;  /* D2H of the transcript                    */
} nwb200_timing;

/* Resources of the context and of its last fill launch: what updateNwAlgPeakMemUsage collects from cudaFuncAttributes x
 * active blocks for the reference's kernels (nwalign_shared.cpp:5-25). */
typedef struct nwb200_mem_usage {
    uint64_t device_bytes;        /* device buffers owned by the context (letters, header rows, snapshots, maps, batch buffers ...) */
    uint64_t pinned_host_bytes;   /* pinned staging buffers                                                                       */
    uint64_t shared_bytes;        /* (static + dynamic shared memory per CTA) x CTAs of the last fill launch                      */
    uint64_t local_bytes;         /* local memory per thread x threads x CTAs                                                      */
    uint64_t register_bytes;      /* registers per thread x 4 x threads x CTAs                                                     */
    int32_t  regs_per_thread, threads_per_block, blocks, reserved;
} nwb200_mem_usage;

NWB200_API int  nwb200_create(nwb200_ctx** out, int device);
NWB200_API void nwb200_destroy(nwb200_ctx* ctx);

/* Replaces initNwInput's subst upload + gapoCost (benchmark.cpp:185-220).  Restrictions the reference does not
 * have (it takes any int): substsz <= 63 and subst[i] - 2*gap <= 255 for every entry (the kernels keep
 * s' = max(subst - 2*gap, 0) as a byte profile); violations return NWB200_ERR_INVALID_VALUE. */
NWB200_API int  nwb200_set_scoring(nwb200_ctx* ctx, const int32_t* subst, int substsz, int gap);

/* Replaces NwAlign_Gpu9_Mlsp_DiagDiagDiag (nwalign_gpu9_mlsp_diagdiagdiag.cu:368-722):
 * H2D of the two sequences, score-matrix fill on the GPU, align_cost back on the host. */
NWB200_API int  nwb200_align_pair_i32(nwb200_ctx* ctx, const int32_t* seqY, int64_t adjrows,
                                      const int32_t* seqX, int64_t adjcols,
                                      const nwb200_params* params /* nullable */, int flags,
                                      int32_t* align_cost, nwb200_hdr_info* hdr /* nullable */);
/* Same with byte letters (no dummy element). */
NWB200_API int  nwb200_align_pair_u8(nwb200_ctx* ctx, const uint8_t* y, int64_t len_y,
                                     const uint8_t* x, int64_t len_x,
                                     const nwb200_params* params, int flags,
                                     int32_t* align_cost, nwb200_hdr_info* hdr);

/* Split form used by the benchmark to time the device part with inputs resident in HBM:
 * upload once, then (re)run fill / trace any number of times. */
NWB200_API int  nwb200_upload_pair_u8(nwb200_ctx* ctx, const uint8_t* y, int64_t len_y,
                                      const uint8_t* x, int64_t len_x, const nwb200_params* params);
NWB200_API int  nwb200_upload_pair_i32(nwb200_ctx* ctx, const int32_t* seqY, int64_t adjrows,
                                       const int32_t* seqX, int64_t adjcols, const nwb200_params* params);
NWB200_API int  nwb200_get_hdr_info(const nwb200_ctx* ctx, nwb200_hdr_info* hdr); /* geometry of the resident pair */
NWB200_API int  nwb200_fill_resident(nwb200_ctx* ctx, int flags);            /* async on the ctx stream */
NWB200_API int  nwb200_trace_resident(nwb200_ctx* ctx);                      /* async on the ctx stream */
NWB200_API int  nwb200_fetch_score(nwb200_ctx* ctx, int32_t* align_cost);    /* syncs */
NWB200_API int  nwb200_fetch_trace(nwb200_ctx* ctx, char* edit_buf, size_t cap, size_t* len,
                                   uint32_t* trace_hash);                    /* syncs */

/* Replaces NwTrace2_Sparse (nwtrace2_sparse.cpp:102-257) for the pair of the last
 * align call made with NWB200_KEEP_HEADERS: tile recompute + walk on the GPU.
 * edit_buf receives the run-length transcript ("<count><op>..." with ops = X I D,
 * nwtrace1_plain.cpp:81-103), not NUL-terminated; *len its length.  If cap is too small
 * the call returns NWB200_ERR_INVALID_VALUE and *len holds the required size. */
NWB200_API int  nwb200_trace_pair(nwb200_ctx* ctx, char* edit_buf, size_t cap, size_t* len,
                                  uint32_t* trace_hash);

/* Replaces the D2H of both header matrices (nwalign_gpu9...cu:701-708): converts the
 * device-resident headers to the reference's tile-major layout so that
 * the reference's own NwTrace2_Sparse / NwHash2_Sparse / NwPrintScore2_Sparse accept them. */
NWB200_API int  nwb200_copy_headers(nwb200_ctx* ctx, int32_t* hrow_host, int32_t* hcol_host);

/* Replaces NwHash2_Sparse / NwHash1_Plain (nwtrace2_sparse.cpp:263-340, nwtrace1_plain.cpp:133-154):
 * the djb2-xor fold over every cell of the score matrix.  The fold is inherently sequential;
 * rows are recomputed on the GPU in slabs from the device-resident sequences and folded on
 * the host as they stream back. */
NWB200_API int  nwb200_score_hash(nwb200_ctx* ctx, uint32_t* score_hash);

/* Diagnostics of the last traceback.  For long pairs the origin maps (pass A of the traceback) are first computed only in corridors
 * around the straight line from (lenY, lenX) to the origin -- 1024 columns either side, then max(4096, lenX / 32), then everything:
 * *corridor_segments of the *segments map segments per band in the first corridor (0: no corridor pass); *corridor_missed = 1 when the
 * path left every corridor and the full pass ran (the result is the same either way).  NWB200_CORRIDOR=<half width in columns> makes
 * it one corridor of that width; 0 switches the corridors off. */
NWB200_API int  nwb200_trace_info(nwb200_ctx* ctx, int* corridor_segments, int* segments, int* corridor_missed);

/* Replaces NwPrintScore2_Sparse's row recomputation (nwtrace2_sparse.cpp:346-419): rows [row0, row0 + nrows) of the
 * full score matrix, adjcols = lenX + 1 ints per row (row 0 / column 0 included), recomputed on the GPU. */
NWB200_API int  nwb200_score_rows(nwb200_ctx* ctx, int64_t row0, int64_t nrows, int32_t* out);

/* calcDebugTrace (nwtrace1_plain.cpp:34-38,107,120-126): the score-matrix values along the traceback path, top-left ->
 * bottom-right, both corners included (*count = moves + 1; NWB200_ERR_INVALID_VALUE with *count set when cap is too
 * small).  Call after nwb200_trace_pair / nwb200_fetch_trace.  O(matrix) work: a debugging aid for small pairs. */
NWB200_API int  nwb200_trace_values(nwb200_ctx* ctx, int32_t* values, size_t cap, size_t* count);

/* Batch of independent pairs (BASELINE config 3): byte letters in one pool, per-pair
 * offsets/lengths.  scores[n_pairs] always; if edits != NULL each pair's transcript is written
 * at edits + edit_off[p] (edit_off has n_pairs + 1 entries: pair p owns the bytes [edit_off[p], edit_off[p+1]);
 * lengths in edit_len[p] out) and its trace hash into trace_hashes[p].  Letters must be < substsz
 * (NWB200_ERR_INVALID_VALUE otherwise).  Pairs of any height are accepted (those taller than 512 rows are
 * re-run through the single-pair kernels). */
NWB200_API int  nwb200_align_batch(nwb200_ctx* ctx, const uint8_t* letters, size_t n_letters,
                                   const uint64_t* offY, const uint32_t* lenY,
                                   const uint64_t* offX, const uint32_t* lenX, size_t n_pairs,
                                   int32_t* scores, char* edits /* nullable */, const uint64_t* edit_off,
                                   uint32_t* edit_len, uint32_t* trace_hashes);
/* Split form: upload once, run on the stream, fetch.  Pairs with more than 512 rows are rejected here
 * (NWB200_ERR_INVALID_VALUE): the split form has no single-pair fix-up pass. */
NWB200_API int  nwb200_upload_batch(nwb200_ctx* ctx, const uint8_t* letters, size_t n_letters,
                                    const uint64_t* offY, const uint32_t* lenY,
                                    const uint64_t* offX, const uint32_t* lenX, size_t n_pairs);
NWB200_API int  nwb200_batch_resident(nwb200_ctx* ctx);                      /* async on the ctx stream */
NWB200_API int  nwb200_fetch_batch_scores(nwb200_ctx* ctx, int32_t* scores); /* syncs */

/* The same batch calls with 5-bit PACKED letters: `packed` holds, from byte offY[p] / offX[p] on, the little-endian bit stream of a
 * sequence's letters, 5 bits each (8 letters in 5 bytes; a sequence of L letters occupies ceil(5 L / 8) bytes); n_bytes = size of the
 * pool.  Scores only.  The one-shot call is bound by the host-to-device copy of the letters: packing takes 3/8 of it away.  Needs an
 * alphabet of at most 31 letters, subst - 2*gap <= 127 and pairs of at most 256 rows (the packed-halves kernel reads the stream). */
NWB200_API int  nwb200_align_batch_packed5(nwb200_ctx* ctx, const uint8_t* packed, size_t n_bytes, const uint64_t* offY, const uint32_t* lenY,
                                           const uint64_t* offX, const uint32_t* lenX, size_t n_pairs, int32_t* scores);
NWB200_API int  nwb200_upload_batch_packed5(nwb200_ctx* ctx, const uint8_t* packed, size_t n_bytes, const uint64_t* offY, const uint32_t* lenY,
                                            const uint64_t* offX, const uint32_t* lenX, size_t n_pairs);

/* Variants the reference lists as future work (README.md:6-29: NW_AG, SW_LG, SW_AG; --gapeCost, cmd_parser.cpp:143,213 is parsed
 * and unused there): affine gaps (Gotoh) and local alignment (Smith-Waterman) for the RESIDENT batch (nwb200_upload_batch), scores
 * only, fetched with nwb200_fetch_batch_scores.  A gap of L residues costs gap_open + (L - 1) * gap_extend -- with
 * gap_extend == gap_open NWB200_VARIANT_NW_AFFINE is the reference's linear-gap recurrence.  No reference implementation exists to
 * compare with ("parity unpinned" except for that case: DESIGN.md).  Pairs of up to 512 rows, substitution
 * scores in [-128, 127]. */
#define NWB200_VARIANT_NW_AFFINE 1
#define NWB200_VARIANT_SW_LINEAR 2
#define NWB200_VARIANT_SW_AFFINE 3
NWB200_API int  nwb200_batch_resident_variant(nwb200_ctx* ctx, int variant, int gap_open, int gap_extend);   /* async on the ctx stream */

/* One very long pair as a cross-GPU wavefront (BASELINE configs 4 and 5; no reference counterpart -- the reference
 * is single-GPU, benchmark.cpp:179).  One context per GPU / process; the columns are dealt to the ranks in blocks of
 * block_cols; border columns travel GPU-to-GPU as peer stores over NVLink into the neighbour's receive buffer, which
 * is shared through a 64-byte CUDA IPC handle:
 *   every rank: nwb200_wave_upload -> nwb200_wave_export(handle) -> [exchange handles] ->
 *               nwb200_wave_connect(handle of rank (r+1) % world; NULL when world == 1) -> [barrier] ->
 *               nwb200_wave_fill(epoch, same non-zero value on every rank) -> nwb200_wave_fetch.
 * The score is returned by the rank that owns the last column block (*has_score = 1). */
NWB200_API int  nwb200_wave_upload(nwb200_ctx* ctx, const uint8_t* y, int64_t len_y, const uint8_t* x, int64_t len_x,
                                   const nwb200_params* params, int rank, int world, int block_cols);
NWB200_API int  nwb200_wave_export(nwb200_ctx* ctx, void* handle64);
NWB200_API int  nwb200_wave_connect(nwb200_ctx* ctx, const void* right_peer_handle64);
NWB200_API int  nwb200_wave_fill(nwb200_ctx* ctx, unsigned epoch);           /* async on the ctx stream */
NWB200_API int  nwb200_wave_fetch(nwb200_ctx* ctx, int* has_score, int32_t* align_cost);   /* syncs */

/* Traceback of a cross-GPU fill (BASELINE config 5 at N > 1; the reference's semantics: nwtrace2_sparse.cpp:102-257 on the header
 * matrices of the whole pair).  Every rank keeps the header rows and snapshots of its own column blocks in the layout of the whole
 * matrix; the rank that walks the path pulls the other ranks' parts over NVLink (peer copies from buffers mapped with CUDA IPC; only
 * the bands of a block that the traceback's corridor can touch unless `full`) and then runs the ordinary traceback:
 *   every rank:  nwb200_wave_keep_headers(1) -> nwb200_wave_upload (block_cols a multiple of tile_cols) -> nwb200_wave_export +
 *                nwb200_wave_export_headers -> [exchange] -> nwb200_wave_connect -> (tracing rank: nwb200_wave_connect_headers)
 *                -> [barrier] -> nwb200_wave_fill -> nwb200_wave_fetch -> [barrier: every fill is complete]
 *   tracing rank: nwb200_wave_gather_headers(0) -> nwb200_trace_resident -> nwb200_trace_info; if the path left the corridor:
 *                nwb200_wave_gather_headers(1) -> nwb200_trace_resident;  -> nwb200_fetch_trace. */
NWB200_API int  nwb200_wave_keep_headers(nwb200_ctx* ctx, int on);
NWB200_API int  nwb200_wave_export_headers(nwb200_ctx* ctx, void* hr_handle64, void* snap_handle64);
NWB200_API int  nwb200_wave_connect_headers(nwb200_ctx* ctx, const void* hr_handles /* world x 64 B */, const void* snap_handles /* world x 64 B */);
NWB200_API int  nwb200_wave_gather_headers(nwb200_ctx* ctx, int full);          /* async on the ctx stream */

/* Few rows x very many columns (BASELINE config 4): the row-parallel prefix-max scorer, one warp per strip of 512
 * columns (groups of 16 strips dealt to the ranks), one int per row crossing each strip (and GPU) boundary.  Score only.  Multi-GPU use follows the wavefront
 * protocol: nwb200_scan_upload -> nwb200_wave_export -> nwb200_wave_connect -> [barrier] -> nwb200_scan_fill -> nwb200_scan_fetch. */
NWB200_API int  nwb200_scan_upload(nwb200_ctx* ctx, const uint8_t* y, int64_t len_y, const uint8_t* x, int64_t len_x, int rank, int world);
NWB200_API int  nwb200_scan_fill(nwb200_ctx* ctx, unsigned epoch);           /* async on the ctx stream */
NWB200_API int  nwb200_scan_fetch(nwb200_ctx* ctx, int* has_score, int32_t* align_cost);   /* syncs */

/* Introspection. */
NWB200_API int         nwb200_last_cuda_error(const nwb200_ctx* ctx);
NWB200_API const char* nwb200_last_error(const nwb200_ctx* ctx);
NWB200_API int         nwb200_get_timing(const nwb200_ctx* ctx, nwb200_timing* out);
NWB200_API void*       nwb200_stream(const nwb200_ctx* ctx);                 /* cudaStream_t */
NWB200_API int         nwb200_sync(nwb200_ctx* ctx);
NWB200_API int         nwb200_kernel_launches(const nwb200_ctx* ctx);        /* kernels launched so far */
NWB200_API int         nwb200_get_memory_usage(const nwb200_ctx* ctx, nwb200_mem_usage* out);
NWB200_API const char* nwb200_batch_kernel_name(const nwb200_ctx* ctx);      /* kernel the last batch launch used ("" before one) */
NWB200_API const char* nwb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NWB200_H */
