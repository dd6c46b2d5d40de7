/*
 * warpcore.h -- C ABI of the B200-native execution core that replaces WarpDB's
 * jit.cpp / operator half of warpdb.cpp / optimizer.cpp / multi_gpu_utils.cpp.
 *
 * The reference has no FFI layer; its seam is the five free functions of
 * include/jit.hpp:7-27 and include/multi_gpu_utils.hpp:10-12 (callers:
 * src/warpdb.cpp:247,365,371,450,454,541,585; src/optimizer.cpp:51; src/main.cu:18,44,331).
 * Each entry point below names the reference interface it replaces.  The C++ shims with the
 * reference's exact signatures live in warpdb_b200/csrc/host/ (jit.hpp, multi_gpu_utils.hpp,
 * optimizer.hpp, warpdb.hpp) and INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; wdb_last_error() returns the
 *    message (thread local).  NVRTC failures use the reference's text "Kernel compilation
 *    failed." (src/jit.cpp:128) with the compile log on stderr (src/jit.cpp:123-125).
 *  - expressions are the CUDA-C strings produced by ASTNode::to_cuda_expr()
 *    (include/expression.hpp:29-79): `col[idx]` names a column, literals carry an `f` suffix.
 *    They are compiled with NVRTC for sm_100a with the reference's flags (arch only, so
 *    --fmad=true, IEEE div/sqrt: src/jit.cpp:114-117), after the UDF source (custom.cu).
 *  - device pointers are plain CUDA pointers of the primary context of `device`; `stream` is a
 *    cudaStream_t (NULL = legacy default stream).  Calls are asynchronous on `stream` unless a
 *    host output pointer (h_*) is passed, in which case they synchronise that stream.
 *  - row counts are 64-bit (the reference's `int N` overflows at 4e9/8e9 rows: SURVEY F8).
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef WARPCORE_H
#define WARPCORE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WDB_ABI_VERSION 2

/* DataType, in the enum order of include/csv_loader.hpp:13 */
enum { WDB_INT32 = 0, WDB_INT64 = 1, WDB_FLOAT32 = 2, WDB_FLOAT64 = 3, WDB_STRING = 4 };
/* AggregationType, in the enum order of include/expression.hpp:86 */
enum { WDB_SUM = 0, WDB_AVG = 1, WDB_COUNT = 2, WDB_MIN = 3, WDB_MAX = 4 };
/* output modes of wdb_project_filter */
enum {
  WDB_DENSE = 0,      /* reference semantics (src/jit.cpp:55-61): rows failing cond leave out[i] untouched */
  WDB_COMPACT = 1,    /* stable stream compaction: surviving values packed in row order */
  WDB_DENSE_ZERO = 2  /* dense, rows failing cond write 0.0f (what WarpDB::query returns: SURVEY App. D) */
};
/* group output order */
enum { WDB_ORDER_FIRST = 0 /* src/jit.cpp:196-213 */, WDB_ORDER_KEY_ASC = 1 /* std::map, src/warpdb.cpp:425 */, WDB_ORDER_KEY_DESC = 2 };

/* ColumnDesc of include/csv_loader.hpp:15-20 (length widened to 64 bit) */
typedef struct wdb_col {
  const char *name;
  int dtype;
  const void *dptr;
  int64_t len;
} wdb_col_t;

/* ---- library / device state ------------------------------------------------------------ */
int wdb_abi_version(void);
const char *wdb_last_error(void);
/* Retain the primary context of `device`, check it is an sm_100-family part and create the
 * per-device kernel cache.  Idempotent.  Replaces cuInit/cuCtxCreate of src/jit.cpp:104,153-155. */
int wdb_init(int device);
int wdb_shutdown(void);
int wdb_device_count(int *out);
/* Source prepended to every generated kernel: the contents of ./custom.cu (src/jit.cpp:65-73).
 * NULL or "" clears it.  The C++ shim re-reads the file per call like the reference does. */
int wdb_set_udf_source(const char *cuda_src);
/* Tuning knobs ("project.variant", "project.block", "project.unroll", ...); see DESIGN.md.
 * value == INT64_MIN removes the override (built-in default). */
int wdb_set_option(const char *key, int64_t value);
int wdb_get_option(const char *key, int64_t *value);
/* Statistics of the kernel cache and of the last call on this thread. */
typedef struct wdb_stats {
  int64_t kernels_compiled;   /* NVRTC compilations so far */
  int64_t cache_hits;
  int64_t launches;           /* kernel launches issued by this library so far */
  double last_compile_ms;     /* NVRTC + module load time of the last miss */
} wdb_stats_t;
int wdb_get_stats(wdb_stats_t *out);

/* ---- fused filter + project:  jit_compile_and_launch (include/jit.hpp:7-10, src/jit.cpp:48-174)
 * cols: all columns of the table (only the ones the strings mention are read).
 * expr: value expression; cond: "" or NULL for none.
 * d_out: float[n] (DENSE*) or float[>= survivors] (COMPACT; n is always enough).
 * d_count: optional device int64 receiving the survivor count (COMPACT) or n (dense modes).
 * h_count: optional host int64; if non-NULL the call synchronises `stream`. */
int wdb_project_filter(int device, void *stream, const wdb_col_t *cols, int ncols, const char *expr,
                       const char *cond, float *d_out, int64_t n, int mode, int64_t *d_count,
                       int64_t *h_count);

/* ---- hash GROUP BY:  jit_group_sum (include/jit.hpp:15-18, src/jit.cpp:179-246) and the
 * std::map aggregation of src/warpdb.cpp:373-437.  key = (int)key_expr, fp64 accumulators.
 * An aggregation table is an opaque per-device object so that chunks (query_multi_gpu_csv) and
 * partial aggregates of other GPUs can be folded into it. */
typedef struct wdb_agg wdb_agg_t;
/* needs: which accumulators the table maintains */
enum { WDB_NEED_SUM = 1, WDB_NEED_COUNT = 2, WDB_NEED_MINMAX = 4, WDB_NEED_FIRST_ROW = 8 };
int wdb_agg_create(int device, int64_t expected_groups, int needs, wdb_agg_t **out);
int wdb_agg_destroy(wdb_agg_t *t);
int wdb_agg_reset(wdb_agg_t *t, void *stream);
/* Optimizer statistics (TableStats, include/csv_loader.hpp:22-37; the reference never fills them,
 * src/optimizer.cpp:13-17): every key the next consume calls will produce lies in [lo, hi] (e.g.
 * wdb_column_minmax of the GROUP BY column).  A small span selects direct-indexed accumulators;
 * keys outside the range stay correct (they take the general path).  known = 0 forgets the range. */
int wdb_agg_set_key_range(wdb_agg_t *t, int known, int64_t lo, int64_t hi);
/* fold n rows into the table (cond "" = none).  row_base is the global index of row 0 (first-appearance order). */
int wdb_agg_consume(wdb_agg_t *t, void *stream, const wdb_col_t *cols, int ncols, const char *val_expr,
                    const char *key_expr, const char *cond, int64_t n, int64_t row_base);
/* fold m partial groups (as exported with the raw arrays below) into the table */
int wdb_agg_merge(wdb_agg_t *t, void *stream, const int32_t *d_keys, const double *d_sums,
                  const int64_t *d_counts, const double *d_mins, const double *d_maxs,
                  const int64_t *d_first, int64_t m);
/* number of groups (synchronises) */
int wdb_agg_size(wdb_agg_t *t, void *stream, int64_t *h_groups);
/* tuning statistic: rows (mod 2^32) since the last reset that the small-cardinality kernel could not
 * keep in its shared-memory accumulators and folded straight into the global table */
int wdb_agg_spilled(wdb_agg_t *t, void *stream, int64_t *h_rows);
/* export groups in `order`; any output pointer may be NULL.  d_vals holds float(agg result)
 * per src/warpdb.cpp:429-435.  cap = capacity of the output arrays.  Synchronises. */
int wdb_agg_export(wdb_agg_t *t, void *stream, int agg, int order, int32_t *d_keys, float *d_vals,
                   double *d_sums, int64_t *d_counts, double *d_mins, double *d_maxs, int64_t *d_first,
                   int64_t cap, int64_t *h_groups);
/* one-shot convenience: create + consume + export + destroy */
int wdb_group_agg(int device, void *stream, const wdb_col_t *cols, int ncols, const char *val_expr,
                  const char *key_expr, const char *cond, int agg, int order, int64_t n,
                  int64_t expected_groups, int32_t *d_keys, float *d_vals, int64_t cap, int64_t *h_groups);

/* ---- ORDER BY ... [LIMIT k [OFFSET o]]:  jit_sort_float + truncate
 * (include/jit.hpp:26-27, src/jit.cpp:283-307, src/warpdb.cpp:453-455,483-495).
 * key_expr orders, val_expr (NULL = key_expr) is returned; ties keep row order (stable).
 * Writes min(k, survivors - offset) values to d_out_vals / d_out_keys (either may be NULL);
 * k < 0 means no LIMIT (full sort; d_out must hold n).  Synchronises when h_n != NULL. */
int wdb_topk(int device, void *stream, const wdb_col_t *cols, int ncols, const char *key_expr,
             const char *val_expr, const char *cond, int descending, int64_t k, int64_t offset,
             int64_t n, float *d_out_vals, float *d_out_keys, int64_t *h_n);

/* ---- in-place stable device sorts: jit_sort_float / jit_sort_pairs
 * (include/jit.hpp:22-27, src/jit.cpp:248-307) */
int wdb_sort_float(int device, void *stream, float *d_vals, int64_t count, int ascending);
int wdb_sort_pairs(int device, void *stream, int32_t *d_keys, float *d_vals, int64_t count, int ascending);

/* ---- optimizer support: column statistics (TableStats, include/csv_loader.hpp:22-37) and
 * zone maps (per-tile min/max) used for pruning; new work, the reference's analyze_condition is
 * a stub (src/optimizer.cpp:13-17). */
int wdb_column_minmax(int device, void *stream, const wdb_col_t *col, double *h_min, double *h_max);
/* Zone map: min/max of every zone of zone_rows rows (power of two >= 2048; 0 = 4096), built in one
 * streaming pass (typically at load time).  Values are recorded as the filter kernel sees them when
 * the column is compared with a float literal (integers converted to float). */
typedef struct wdb_zonemap wdb_zonemap_t;
int wdb_zonemap_build(int device, void *stream, const wdb_col_t *col, int64_t zone_rows, wdb_zonemap_t **out);
/* Resident ingest (upload_to_gpu, src/csv_loader.cpp:126-161: one synchronous pageable cudaMemcpy per
 * column, no statistics): host column (h_col->dptr is a HOST pointer) -> d_dst in 64 MB chunks, pinned
 * sources copied asynchronously, pageable ones through a pinned two-slot ring; the zone map
 * (zone_rows as above; < 0 or out_zm == NULL: none) is filled on the device chunk by chunk while the
 * next chunk is on the wire, and the exact column min/max (TableStats) is returned when asked for. */
int wdb_upload_column(int device, void *stream, const wdb_col_t *h_col, void *d_dst, int64_t zone_rows, wdb_zonemap_t **out_zm,
                      double *h_min, double *h_max);
int wdb_zonemap_destroy(wdb_zonemap_t *z);
int wdb_zonemap_info(const wdb_zonemap_t *z, int64_t *zone_rows, int64_t *nzones);
/* one `column <op> constant` term of a conjunction; op: 0 '>' 1 '>=' 2 '<' 3 '<=' 4 '==' 5 '!=' */
typedef struct wdb_prune {
  const wdb_zonemap_t *zonemap;
  int op;
  double value;
} wdb_prune_t;
/* wdb_project_filter with zone-map pruning: `cond` is still evaluated on every row that is read,
 * preds (implied by cond: the AND-ed col-vs-constant terms of it) only decide which zones are read
 * at all.  Results are identical to wdb_project_filter.  h_zones_live (optional) = zones not pruned. */
int wdb_project_filter_pruned(int device, void *stream, const wdb_col_t *cols, int ncols, const char *expr,
                              const char *cond, float *d_out, int64_t n, int mode, int64_t *d_count,
                              int64_t *h_count, const wdb_prune_t *preds, int npreds, int64_t *h_zones_live);

/* The same pruning for GROUP BY and ORDER BY ... LIMIT with a WHERE clause: the kernels run over the
 * maximal runs of live zones only (few, long runs on sorted / clustered columns; on a random layout the
 * plain call is taken).  Results are identical to wdb_agg_consume / wdb_topk. */
int wdb_agg_consume_pruned(wdb_agg_t *t, void *stream, const wdb_col_t *cols, int ncols, const char *val_expr, const char *key_expr,
                           const char *cond, int64_t n, int64_t row_base, const wdb_prune_t *preds, int npreds, int64_t *h_zones_live);
int wdb_topk_pruned(int device, void *stream, const wdb_col_t *cols, int ncols, const char *key_expr, const char *val_expr, const char *cond,
                    int descending, int64_t k, int64_t offset, int64_t n, float *d_out_vals, float *d_out_keys, int64_t *h_n,
                    const wdb_prune_t *preds, int npreds, int64_t *h_zones_live);

/* ---- inner equi-join on integer keys.  `JOIN <table> ON <expr>` is parsed by the reference
 * (JoinClause, include/expression.hpp:123-135; src/expression.cpp:375-401) and validated
 * (src/warpdb.cpp:321-323) but never executed ("Currently JOIN loads the same table for
 * demonstration purposes", include/warpdb.hpp:22); these entry points are what a query_sql that
 * does execute it binds.  Result = every (probe row i, build row j) with probe_key[i] ==
 * build_key[j], ordered by i, then j (nested-loop order).  Keys are WDB_INT32 or WDB_INT64 columns
 * (the two sides may differ); the build side holds at most 2^32 - 1 rows.
 *   wdb_join_build   sorts (key, row) of the build column once; the index can be probed many times
 *                    (from `stream`, or after it has drained).  A dense key range -- at most 8 key
 *                    values per build row -- is also tabulated for direct addressing.
 *   wdb_join_probe   d_probe_rows == d_build_rows == NULL: count only (*h_pairs); otherwise the pairs
 *                    are written (cap = capacity of each array in pairs; more pairs than cap is an
 *                    error and nothing is written).  Synchronous: returns after the stream drained.
 *                    A count-only call leaves its per-tile counts in the index, and an emitting call
 *                    on the same probe column (same pointer, length and type) that follows it reuses
 *                    them instead of counting again -- do not modify the column between the two.
 *   wdb_gather       d_dst[i] = src[d_rows[i]] for i < count (4- and 8-byte column types; d_rows ==
 *                    NULL copies the first count rows): materialises the columns a joined query reads,
 *                    which the other operators then consume unchanged.  Rows are not range-checked. */
typedef struct wdb_join wdb_join_t;
int wdb_join_build(int device, void *stream, const wdb_col_t *build_key, wdb_join_t **out);
int wdb_join_probe(wdb_join_t *j, void *stream, const wdb_col_t *probe_key, int64_t *d_probe_rows, int64_t *d_build_rows, int64_t cap,
                   int64_t *h_pairs);
/* direct_span: number of key values tabulated for direct addressing (0: probes binary-search) */
int wdb_join_info(const wdb_join_t *j, int64_t *build_rows, int *key_dtype, int64_t *direct_span);
int wdb_join_destroy(wdb_join_t *j);
int wdb_gather(int device, void *stream, const wdb_col_t *src, const int64_t *d_rows, int64_t count, void *d_dst);

/* ---- multi-GPU: run_multi_gpu_jit_host (include/multi_gpu_utils.hpp:10-12,
 * src/multi_gpu_utils.cpp:5-63).  Host columns in, host floats out; rows are split into
 * contiguous shards chunk = ceil(n/ndev) and all devices run concurrently on their own
 * streams (the reference loops over them sequentially).  h_cols[i].dptr are HOST pointers (pinned
 * memory makes the copies asynchronous).  ndev <= 0: all devices; devices: NULL = 0..ndev-1, else
 * the device ids to use (one process per GPU passes {its device}). */
int wdb_multi_project_filter_host(int ndev, const int *devices, const wdb_col_t *h_cols, int ncols,
                                  const char *expr, const char *cond, float *h_out, int64_t n, int mode,
                                  int64_t *h_count);
/* shard [start,end) of device dev: src/multi_gpu_utils.cpp:24-31 */
int wdb_shard_range(int64_t n, int ndev, int dev, int64_t *start, int64_t *end);

/* ---- sharded operators with the cross-GPU merge inside the core.
 * The reference's multi-GPU driver (src/multi_gpu_utils.cpp:5-63, src/warpdb.cpp:508-542) can only
 * project / filter and moves every shard's output through the host; these entry points keep the
 * row-range shards (chunk = ceil(n/ndev), :24-31) resident on their GPUs and merge partial
 * aggregates / top-k candidates GPU to GPU with NCCL over NVLink.
 *
 * A communicator binds one GPU to a group of `nranks` GPUs.  One process per GPU: rank 0 calls
 * wdb_comm_unique_id, the 128 bytes travel by any means (torch.distributed, MPI, a file) and every
 * rank calls wdb_comm_init_rank.  One process driving all GPUs (the reference's shape):
 * wdb_comm_init_all fills out[0..ndev) and each communicator is then used from its own host thread.
 * nranks == 1 needs neither an id nor NCCL. */
typedef struct wdb_comm wdb_comm_t;
int wdb_comm_unique_id(void *id128);
int wdb_comm_init_rank(int device, int nranks, int rank, const void *id128, wdb_comm_t **out);
int wdb_comm_init_all(int ndev, const int *devices, wdb_comm_t **out);
int wdb_comm_destroy(wdb_comm_t *c);
int wdb_comm_info(const wdb_comm_t *c, int *rank, int *nranks, int *device);
/* wdb_project_filter on this rank's shard (no data-path collective: results stay sharded, rank order is
 * row order) plus an all-gather of one count per rank: count3 = {rows written by this rank, global
 * offset of its first row, global total}.  d_count3 (device) / h_count3 (host, synchronises) optional. */
int wdb_multi_project_filter(wdb_comm_t *c, void *stream, const wdb_col_t *cols, int ncols, const char *expr, const char *cond,
                             float *d_out, int64_t n_local, int mode, int64_t *d_count3, int64_t *h_count3);
/* GROUP BY over all shards (jit_group_sum + the std::map merge, src/jit.cpp:179-246, src/warpdb.cpp:373-437):
 * cols hold this rank's shard (n_local rows, the first one being global row row_base); every rank ends
 * with the same final groups in d_keys / d_vals (capacity cap) and their number in *d_groups (device)
 * and *h_groups (host; synchronises; NULL = fully asynchronous).  range_known != 0: every key of every
 * shard lies in [key_lo, key_hi] (TableStats of the key column); otherwise the core measures it. */
int wdb_multi_group_agg(wdb_comm_t *c, void *stream, const wdb_col_t *cols, int ncols, const char *val_expr, const char *key_expr,
                        const char *cond, int agg, int order, int64_t n_local, int64_t row_base, int64_t expected_groups,
                        int range_known, int64_t key_lo, int64_t key_hi, int32_t *d_keys, float *d_vals, int64_t cap,
                        int64_t *d_groups, int64_t *h_groups);
/* ORDER BY key [DESC] LIMIT k OFFSET offset over all shards (jit_sort_float + truncate, src/jit.cpp:283-307,
 * src/warpdb.cpp:483-495): every rank ends with the same min(k, survivors - offset) values; ties keep
 * global row order.  *d_n (device) / *h_n (host; synchronises) receive the number of values. */
int wdb_multi_topk(wdb_comm_t *c, void *stream, const wdb_col_t *cols, int ncols, const char *key_expr, const char *val_expr,
                   const char *cond, int descending, int64_t k, int64_t offset, int64_t n_local, int64_t row_base,
                   float *d_out_vals, float *d_out_keys, int64_t *d_n, int64_t *h_n);
/* Host columns in, host results out, one process driving ndev GPUs (<= 0: all) from one thread per
 * device: the aggregate / ORDER BY counterparts of wdb_multi_project_filter_host. */
int wdb_multi_group_agg_host(int ndev, const int *devices, const wdb_col_t *h_cols, int ncols, const char *val_expr, const char *key_expr,
                             const char *cond, int agg, int order, int64_t n, int64_t expected_groups, int32_t *h_keys, float *h_vals,
                             int64_t cap, int64_t *h_groups);
int wdb_multi_topk_host(int ndev, const int *devices, const wdb_col_t *h_cols, int ncols, const char *key_expr, const char *val_expr,
                        const char *cond, int descending, int64_t k, int64_t offset, int64_t n, float *h_out_vals, int64_t *h_n);

/* ---- synthetic columns for benchmarks/tests (counter based, bit-identical to the oracle's
 * orc_synth_*; not part of the reference) */
int wdb_synth_f32(int device, void *stream, float *d_out, int64_t n, uint64_t seed, float lo, float hi, int64_t row0);
int wdb_synth_i32(int device, void *stream, int32_t *d_out, int64_t n, uint64_t seed, int32_t lo, int32_t hi_excl, int64_t row0);

/* ---- introspection for tests: generate (and compile for `arch`, e.g. "sm_100a") the kernel a
 * call would use, without a device.  kind: "project", "filter", "compact", "group", "topk".
 * Returns the CUDA source and/or the CUBIN through malloc'd buffers the caller frees with wdb_free. */
int wdb_debug_compile(const char *kind, const wdb_col_t *cols, int ncols, const char *expr_a,
                      const char *expr_b, const char *cond, int mode, const char *arch, char **out_source,
                      void **out_cubin, size_t *out_cubin_size);
void wdb_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
