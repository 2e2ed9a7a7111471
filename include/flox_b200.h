/* flox_b200 — C ABI of the B200 grouped-reduction engine.
 *
 * This is the drop-in boundary for flox's per-chunk hot path.  The reference has no FFI (it is
 * pure Python); its plug-in seam is the *engine function contract*
 *     f(group_idx, array, *, axis=-1, size=None, fill_value=None, dtype=None, **kw) -> ndarray
 * resolved by name in flox/aggregations.py:85-90 (generic_aggregate) and called from
 * flox/core.py:368-370 (chunk_reduce).  Each entry point below cites the reference function(s)
 * it replaces.  A maintainer binds them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter is documented as host;
 *   - no entry point allocates or synchronises; work is enqueued on `stream` (a cudaStream_t
 *     passed as void*, NULL = the legacy default stream);
 *   - return value: 0 on success, negative FLOXB_E* otherwise (floxb_strerror gives text);
 *   - a chunk is values[A, n, B] (C-contiguous) + codes[n]; codes outside [0, n_groups) are
 *     skipped (flox/factorize.py:44-53,201-210 + flox/core.py:381-383 fused away);
 *   - accumulators are laid out [A][n_groups] when B == 1 and GROUP-MAJOR [n_groups][A*B] when B > 1 (kept
 *     trailing dims: a row's B values then update one contiguous run); floxb_finalize(group_major_ab = A*B)
 *     turns the latter into the reference's [A, B, n_groups] result order.  They are WIDE: f64 for float inputs, i64 for integer
 *     inputs, i64 counters.  Several calls may accumulate into the same buffers (init=0):
 *     that is flox's _simple_combine (flox/dask.py:90-144) fused into the chunk kernel.
 */
#ifndef FLOX_B200_H_
#define FLOX_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLOXB_OK 0
#define FLOXB_EINVAL (-1)  /* bad argument (null pointer, negative size, ...) */
#define FLOXB_EDTYPE (-2)  /* unsupported dtype combination */
#define FLOXB_ECUDA (-3)   /* a CUDA runtime call failed; see floxb_last_cuda_error() */
#define FLOXB_ERANGE (-4)  /* size beyond what the kernels index */

/* element types */
#define FLOXB_F32 0
#define FLOXB_F64 1
#define FLOXB_I32 2
#define FLOXB_I64 3

/* floxb_moments op */
#define FLOXB_OP_SUM 0   /* aggregate_flox.sum / nansum            (aggregate_flox.py:210-211) */
#define FLOXB_OP_PROD 1  /* aggregate_flox.prod / nanprod          (aggregate_flox.py:212-213) */
#define FLOXB_OP_SUMSQ 2 /* aggregate_flox.(nan)sum_of_squares     (aggregate_flox.py:225-244) */

/* floxb_finalize kind */
#define FLOXB_FIN_SUM 0     /* acc -> out; NaN-propagating unless skipna          */
#define FLOXB_FIN_NANLEN 1  /* cnt -> out                      (aggregate_flox.py:247-248) */
#define FLOXB_FIN_MEAN 2    /* acc/cnt                         (aggregate_flox.py:251-266) */
#define FLOXB_FIN_MAX 3     /* key acc -> out                  (aggregate_flox.py:214-215) */
#define FLOXB_FIN_MIN 4     /*                                 (aggregate_flox.py:216-217) */
#define FLOXB_FIN_VAR 5     /* M2/(n-ddof)                     (aggregations.py:460-468)   */
#define FLOXB_FIN_STD 6     /* sqrt of the above               (aggregations.py:471-472)   */
#define FLOXB_FIN_ARG 7     /* index acc -> out, fill where no winner                       */
#define FLOXB_FIN_PICK 8    /* picked value acc -> out (first / last); fill where the group has no row, or (skipna)
                               no non-NaN row                                    (aggregations.py:634-651) */

/* floxb_combine kind (flox/dask.py:131-143 with agg.simple_combine = np.sum/np.max/np.min) */
#define FLOXB_COMB_SUM 0
#define FLOXB_COMB_MAX 1
#define FLOXB_COMB_MIN 2
#define FLOXB_COMB_PROD 3
#define FLOXB_COMB_NANMAX 4 /* np.nanmax: NaN blocks are skipped */
#define FLOXB_COMB_NANMIN 5

typedef struct floxb_chunk {
  const void* values; /* device, [A, n, B] */
  const void* codes;  /* device, [n] */
  int64_t n;          /* length of the reduced axis in this chunk */
} floxb_chunk_t;

/* Library / build info. */
int floxb_version(void);
const char* floxb_strerror(int code);
const char* floxb_last_cuda_error(void);
/* Largest n_groups the shared-memory (privatised-bin) kernel path takes for this op family;
 * above it the L2-atomic path is used.  Informational. */
int64_t floxb_smem_group_limit(int value_dtype, int family /*0 moments, 1 minmax*/);
/* How many chunks (dask blocks, flox/core.py:1614-1626 `map_blocks(chunk_reduce)`) one persistent launch of the
 * shared-memory path walks; longer chunk lists are split into several launches.  Informational. */
int floxb_max_chunks_per_launch(void);

/* Grouped sum / product / sum-of-squares plus non-NaN and NaN counts, one pass.
 * Replaces aggregate_flox._np_grouped_op(op=np.add|multiply.reduceat) and nanlen
 * (flox/aggregate_flox.py:133-192,195-217,225-248) and the sort in _prepare_for_flox
 * (flox/aggregate_flox.py:9-23), which a scatter reduction does not need.
 *   chunks     HOST array of n_chunks descriptors
 *   acc        f64[A*B*G] for float values, i64[A*B*G] for integer values (non-NaN values only)
 *   cnt, nnan  i64[A*B*G]: number of non-NaN values per group; nnan > 0 iff the group has NaN values
 *              (an exact NaN count on the L2-atomic and column paths, the number of CTAs that saw a
 *              NaN on the shared-memory f32 path)
 *   init       1 = reset the accumulators first (0 / 1 for PROD), 0 = accumulate on top
 *   force_path 0 = automatic, 1 = shared-memory bins, 2 = L2 atomics, 3 = shared-memory bins with the f32 CAS-loop
 *              level 1 instead of the fixed-point one, 4 = the per-element L2-atomic kernel for small kept trailing
 *              dims instead of the run-merging one, 5 = run-merging with per-element counts (testing / profiling)
 *   workspace  optional scratch, see floxb_moments_workspace_bytes */
int floxb_moments(const floxb_chunk_t* chunks, int64_t n_chunks, int value_dtype, int code_dtype,
                  int64_t A, int64_t B, int64_t n_groups, int op, void* acc, int64_t* cnt,
                  int64_t* nnan, int init, int force_path, void* workspace, int64_t workspace_bytes,
                  void* stream);
/* Optional scratch for floxb_moments (device memory, need not be initialised, may be NULL): with at least
 * this many bytes the f32 shared-memory kernel keeps its f64 level in L2-resident global memory instead of
 * shared memory (8 instead of 16 bytes of shared memory per group -> more L1 for the streaming loads, and
 * twice the groups fit).  0 = the call has no use for a workspace.  Must not be shared by concurrent calls. */
int64_t floxb_moments_workspace_bytes(int value_dtype, int64_t A, int64_t B, int64_t n_groups);

/* Grouped max / min of the non-NaN values as order-preserving i64 keys, plus the counts.
 * Replaces aggregate_flox.max/min/nanmax/nanmin (flox/aggregate_flox.py:214-217,195-207).
 *   acc_key    i64[A*B*G]; init sets it to INT64_MIN (max) / INT64_MAX (min) */
int floxb_minmax(const floxb_chunk_t* chunks, int64_t n_chunks, int value_dtype, int code_dtype,
                 int64_t A, int64_t B, int64_t n_groups, int is_max, int64_t* acc_key,
                 int64_t* cnt, int64_t* nnan, int init, int force_path, void* stream);

/* Row permutation grouped by code, for the column-owner kernels: perm[group_offsets[g] .. group_offsets[g+1])
 * lists the rows whose code is g (order within a group unspecified); rows with codes outside [0, n_groups) are
 * left out.  cursor: i64[n_groups] scratch.  Replaces the stable argsort of _prepare_for_flox
 * (flox/aggregate_flox.py:9-23) for this layout — an O(n) counting sort over the LABELS only. */
int floxb_group_index(const void* codes, int code_dtype, int64_t n, int64_t n_groups, int64_t* perm,
                      int64_t* group_offsets /* [n_groups+1] */, int64_t* cursor, void* stream);

/* Column-owner reduction of ONE chunk values[A, n, B] (B > 1: kept trailing dims, e.g. (time, lat, lon) grouped
 * along time — the layout flox reaches via _move_reduce_dims_to_end + reshape copy, flox/core.py:1014-1018).
 * mode: 0 sum, 1 prod, 2 sum of squares (acc = f64/i64), 3 max, 4 min (acc = i64 keys); cnt/nnan as in
 * floxb_moments.  No atomics: each (a, b, g) bin is produced in registers by one thread.
 * init=1 overwrites the accumulators, init=0 merges into them.
 * group_major=1 (B > 1 only): accumulators are laid out [n_groups][A*B] instead of [A*B][n_groups] so that
 * the owner threads write full sectors; pass the same A*B to floxb_finalize(group_major_ab). */
int floxb_reduce_cols(const void* values, int value_dtype, const int64_t* perm, const int64_t* group_offsets,
                      int64_t A, int64_t n, int64_t B, int64_t n_groups, int mode, void* acc, int64_t* cnt,
                      int64_t* nnan, int init, int group_major, void* stream);

/* Second pass of var_chunk: sum over rows of (x - mean[group])^2
 * (flox/aggregations.py:375-387).  `means` is f64[A*B*G] (see floxb_group_means). */
int floxb_sqdev(const floxb_chunk_t* chunks, int64_t n_chunks, int value_dtype, int code_dtype,
                int64_t A, int64_t B, int64_t n_groups, const double* means, int skipna,
                double* acc_m2, int init, void* stream);

/* means[g] = acc[g]/cnt[g], rounded through `value_dtype` like `array_sums / array_lens`
 * (flox/aggregations.py:375-376); NaN when skipna==0 and the group saw a NaN. */
int floxb_group_means(int value_dtype, const void* acc, const int64_t* cnt, const int64_t* nnan,
                      int64_t n_bins, int skipna, double* means, void* stream);

/* Position of the first extremum per group (numpy first-occurrence semantics; the reference
 * pins them at tests/test_core.py:295,388-408; chunk_argreduce flox/core.py:157-211).
 * Run floxb_minmax first; this pass finds, per group, the smallest `index_base + i` whose
 * value equals the extremum (or is the first NaN when skipna==0 and the group has a NaN).
 *   acc_idx    i64[A*B*G]; init sets INT64_MAX (= "no winner") */
int floxb_argext(const floxb_chunk_t* chunks, int64_t n_chunks, int value_dtype, int code_dtype,
                 int64_t A, int64_t B, int64_t n_groups, const int64_t* acc_key,
                 const int64_t* nnan, int skipna, const int64_t* index_base /*host, per chunk*/,
                 int64_t* acc_idx, int init, void* stream);

/* first / last / nanfirst / nanlast (flox/aggregations.py:634-651; combine: xrutils.nanfirst / nanlast,
 * flox/xrutils.py:386-412).  floxb_firstlast: acc_pos[bin] = smallest (is_last: largest) `index_base + i` over the rows of
 * the group — over its non-NaN rows when skipna; init sets INT64_MAX (INT64_MIN) = "no row".  floxb_take: the row whose
 * position equals pos[bin] writes its value into acc[bin] (f64 for float values, i64 for integers; init zeroes acc, so that
 * ranks merge by SUM after merging positions by MIN / MAX). */
int floxb_firstlast(const floxb_chunk_t* chunks, int64_t n_chunks, int value_dtype, int code_dtype, int64_t A, int64_t B,
                    int64_t n_groups, int is_last, int skipna, const int64_t* index_base /*host, per chunk*/,
                    int64_t* acc_pos, int init, void* stream);
int floxb_take(const floxb_chunk_t* chunks, int64_t n_chunks, int value_dtype, int code_dtype, int64_t A, int64_t B,
               int64_t n_groups, const int64_t* pos, const int64_t* index_base /*host, per chunk*/, void* acc, int init,
               void* stream);

/* Typed result from wide accumulators: flox/core.py:410-475 (_finalize_results) pieces that
 * are arithmetic — fill of absent groups (aggregate_flox.py:174-180), NaN propagation,
 * mean = sum/count (aggregations.py:323-325), var = M2/(n-ddof) (aggregations.py:460-468),
 * min_count masking (core.py:438-459).
 *   acc        f64/i64 sums, i64 keys (MAX/MIN) or i64 indices (ARG); for VAR/STD the M2 (f64)
 *   acc2       VAR/STD only: the f64 sum (unused otherwise, may be NULL)
 *   out_dtype  FLOXB_* of `out`
 *   fill_f/fill_i  value for groups with no rows (float / integer `out`)
 *   min_count  >0: groups with cnt < min_count receive mc_fill_* instead (core.py:438-459)
 *   ddof       VAR/STD only
 *   group_major_ab  0: accumulators are [A*B][G] like `out`; A*B: accumulators are [G][A*B]
 *                   (written by floxb_reduce_cols(group_major=1)), `out` is still [A*B][G] */
int floxb_finalize(int kind, int value_dtype, int out_dtype, const void* acc, const void* acc2,
                   const int64_t* cnt, const int64_t* nnan, int64_t n_bins, int skipna,
                   double fill_f, int64_t fill_i, int64_t min_count, double mc_fill_f,
                   int64_t mc_fill_i, int64_t ddof, void* out, int64_t group_major_ab, void* stream);

/* Reduce `nblk` stacked accumulator blocks [nblk, n_bins] into out[n_bins]
 * (flox/dask.py:131-143).  dtype: FLOXB_F64 or FLOXB_I64. */
int floxb_combine(int kind, int dtype, const void* stacked, int64_t nblk, int64_t n_bins,
                  void* out, void* stream);

/* Chan merge of stacked (M2, S, n) triples, each f64[nblk, n_bins] -> f64[n_bins]
 * (flox/aggregations.py:392-451, _var_combine). */
int floxb_var_combine(const double* m2, const double* s, const double* n, int64_t nblk,
                      int64_t n_bins, double* out_m2, double* out_s, double* out_n, void* stream);

/* Label encoding on device (flox/factorize.py:44-82).
 *   floxb_digitize: codes[i] = bin index of labels[i] in `edges` (closed right or left), -1 when
 *   outside or NaN — np.digitize + the pd.cut fix-up (factorize.py:55-82).  label_dtype F32/F64/I64.
 *   floxb_ravel_codes: combined = c0*size1 + c1, -1 if either is -1 (factorize.py:102-108). */
int floxb_digitize(const void* labels, int label_dtype, int64_t n, const double* edges,
                   int64_t n_edges, int closed_right, int64_t* codes, void* stream);
int floxb_ravel_codes(const int64_t* c0, const int64_t* c1, int64_t n, int64_t size0, int64_t size1,
                      int64_t* out, void* stream);

/* Discovery of UNKNOWN integer labels on the device (expected_groups=None: the reference runs pd.factorize on
 * the host, flox/factorize.py:95-99, a hash pass of ~5 ns per label):
 *   floxb_label_minmax   : minmax[0] = min(labels), minmax[1] = max(labels)   (device int64[2]; n > 0)
 *   floxb_label_presence : table[labels[i] - lo] = 1 for lo <= labels[i] < lo + range   (table: device u8[range],
 *                          zeroed by the call).  The sorted unique labels are lo + nonzero(table).
 *   floxb_lut_codes      : codes[i] = lut[labels[i] - lo] if lo <= labels[i] < lo + range else -1 — also the
 *                          searchsorted + isin of known integer groups (factorize.py:84-94).
 * label_dtype FLOXB_I32 / FLOXB_I64. */
int floxb_label_minmax(const void* labels, int label_dtype, int64_t n, int64_t* minmax, void* stream);
int floxb_label_presence(const void* labels, int label_dtype, int64_t n, int64_t lo, int64_t range,
                         uint8_t* table, void* stream);
int floxb_lut_codes(const void* labels, int label_dtype, int64_t n, int64_t lo, int64_t range,
                    const int64_t* lut, int64_t* codes, void* stream);

/* Label staging for host-resident chunks.  flox hands every block int64 codes (flox/factorize.py:201-210,
 * flox/core.py:298-316); over PCIe they are 2/3 of the bytes of an f32 reduction.  When n_groups < 2^15 (2^31) the
 * codes fit 16 (32) bits:
 *   floxb_host_pack_codes  (HOST function, no CUDA call; thread-safe, callers parallelise over chunks):
 *       out[i] = codes[i] if 0 <= codes[i] < n_groups else -1, narrowed to int16 / int32.  `out` is normally a
 *       pinned staging buffer.  code_dtype FLOXB_I32 / FLOXB_I64, out_bits 16 or 32.
 *   floxb_widen_codes      (device): int16 codes -> int32 codes, the narrowest type the reduction kernels read. */
int floxb_host_pack_codes(const void* codes, int code_dtype, int64_t n, int64_t n_groups, void* out, int out_bits);
int floxb_widen_codes(const int16_t* codes16, int64_t n, int32_t* codes32, void* stream);

/* ---- sorted-segment family: grouped quantile / median ----------------------------------------
 * Replaces aggregate_flox.quantile_ (flox/aggregate_flox.py:50-130) behind quantile / nanquantile / median / nanmedian
 * (:218-221).  The caller sorts every row by (group, value), NaNs last within their group (the reference partitions a
 * complex array group + 1j * value for the same order, :100-113), and passes
 *   sorted   [n_rows, n] values (F32 / F64)                  starts   i64[n_groups + 1]: segment of group g in every row
 *   floxb_segment_valid     nvalid[n_rows, n_groups] = non-NaN values per (row, group)
 *   floxb_segment_quantile  out[n_q, n_rows, n_groups]: numpy's method="linear" quantile of the valid values, NaN for a
 *                           group without valid values (or, skipna = 0, holding a NaN), `fill` for a group that does not
 *                           occur.  before[n_rows, n_groups] = exclusive running sum of nvalid along the groups of a row:
 *                           the reference's virtual index is formed on the row's compacted valid values and this offset
 *                           decides the last bits of the interpolation weight (:85-90); results are bit-identical. */
int floxb_segment_valid(const void* sorted, int value_dtype, int64_t n_rows, int64_t n, const int64_t* starts,
                        int64_t n_groups, int64_t* nvalid, void* stream);
int floxb_segment_quantile(const void* sorted, int value_dtype, int64_t n_rows, int64_t n, const int64_t* starts,
                           int64_t n_groups, const int64_t* nvalid, const int64_t* before, const double* q, int n_q,
                           int skipna, double fill, int out_dtype, void* out, void* stream);

/* Grouped scans within groups, in order of appearance (flox/aggregate_flox.py:296-329 cumsum / nancumsum, :269-293 ffill).
 *   values        [n_rows, n] in the ORIGINAL order (F32 / F64 / I32 / I64); out has the same shape, order and dtype
 *   perm          i64[n]: stable argsort of the group codes (the reference's _prepare_for_flox permutation, labels only)
 *   sorted_codes  i64[n]: codes[perm]
 *   kind          FLOXB_SCAN_CUMSUM (skipna = 1: NaN counts as 0 = nancumsum; 0: a NaN poisons the rest of ITS group)
 *                 or FLOXB_SCAN_FFILL (a NaN takes the last valid value before it within its group)
 *   scratch       device memory of at least floxb_segment_scan_scratch_bytes(n_rows, n)
 * A true segmented scan: sums accumulate in f64 / i64 per group and are rounded once (the reference subtracts the total that
 * precedes the group from one global running sum, so its results carry that total's rounding). */
#define FLOXB_SCAN_CUMSUM 0
#define FLOXB_SCAN_FFILL 1
int64_t floxb_segment_scan_scratch_bytes(int64_t n_rows, int64_t n);
int floxb_segment_scan(int kind, const void* values, int value_dtype, const int64_t* perm, const int64_t* sorted_codes,
                       int64_t n_rows, int64_t n, int skipna, void* scratch, int64_t scratch_bytes, void* out, void* stream);

/* ---- cross-GPU combine over NVLink peer memory (one box, one process per GPU) ---------------------------------------
 * Replaces the exchange step of flox's map-reduce — `_tree_reduce(combine=_simple_combine)`, flox/dask.py:440-460 with
 * flox/dask_array_ops.py:29-121 — for accumulator sets small enough to be latency-bound (NCCL remains the path for
 * large ones, flox_b200/parallel.py).  Every rank owns a SLAB:
 *     [header: FLOXB_XCHG_HEADER_BYTES][buffer 0][buffer 1]        (the caller double-buffers: collective s uses s % 2)
 * allocated with floxb_ipc_alloc and mapped by every peer with floxb_ipc_open (CUDA IPC; handles travel through the
 * host-side process group).  The reduction kernels accumulate straight into the slab; floxb_xchg_allreduce then runs a
 * flag barrier over peer memory and merges word j of all `world` buffers in rank order (bit-identical results on every
 * rank) into local `outs`.  Collectives must be issued in the same order with the same `epoch` (1, 2, 3, ...) on all ranks.
 *   floxb_ipc_alloc   cudaMalloc + zero + export: *dptr = slab, handle64 = 64-byte cudaIpcMemHandle_t (HOST memory)
 *   floxb_ipc_open    map a peer's handle (peer access is enabled lazily); floxb_ipc_close / floxb_ipc_free undo both
 *   floxb_xchg_allreduce
 *       slab_bases   HOST array of `world` device pointers: this process's mapping of every rank's slab (own at [rank])
 *       data_offset  byte offset of the buffer in use inside each slab (>= FLOXB_XCHG_HEADER_BYTES, multiple of 8)
 *       seg_words / seg_ops / outs   HOST arrays, n_segments <= FLOXB_XCHG_MAX_SEGMENTS: consecutive runs of 8-byte
 *                    words in the buffer, how to combine each (FLOXB_XCHG_*), and the local device destination of each
 *   floxb_xchg_status *failed_epoch != 0: a barrier gave up waiting (about a minute) for a peer in that collective (the results
 *                    of it and everything after are undefined).  Synchronises the device. */
#define FLOXB_XCHG_MAX_RANKS 8
#define FLOXB_XCHG_MAX_SEGMENTS 4
#define FLOXB_XCHG_HEADER_BYTES 128
#define FLOXB_XCHG_SUM_F64 0
#define FLOXB_XCHG_SUM_I64 1
#define FLOXB_XCHG_MAX_I64 2
#define FLOXB_XCHG_MIN_I64 3
#define FLOXB_XCHG_PROD_F64 4
#define FLOXB_XCHG_PROD_I64 5
int floxb_ipc_alloc(int64_t bytes, void** dptr, void* handle64);
int floxb_ipc_open(const void* handle64, void** dptr);
int floxb_ipc_close(void* dptr);
int floxb_ipc_free(void* dptr);
int floxb_xchg_allreduce(void* const* slab_bases, int world, int rank, uint64_t epoch, int64_t data_offset,
                         int n_segments, const int64_t* seg_words, const int* seg_ops, void* const* outs, void* stream);
int floxb_xchg_status(const void* slab_base, uint64_t* failed_epoch);

#ifdef __cplusplus
}
#endif
#endif /* FLOX_B200_H_ */
