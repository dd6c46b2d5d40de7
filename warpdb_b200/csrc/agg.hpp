// agg.hpp -- the aggregation-table object behind wdb_agg_t, shared by ops_group.cu (consume / merge /
// export) and ops_comm.cu (cross-GPU merge of partial aggregates).  Internal, not part of the C ABI.
#pragma once
#include <functional>

#include "core.hpp"
#include "kernels/group_table.cuh"

struct wdb_agg {
  wdb::Device *dev = nullptr;
  int needs = 0;
  int64_t cap = 0;          // power of two
  char *mem = nullptr;
  wdb_table T{};
  // direct-addressed side table (T.dsums / T.dcnts), allocated on first use
  char *dense_mem = nullptr;
  int64_t dense_cap = 0;    // allocated entries
  bool dense_live = false;  // holds aggregates (T.dspan > 0)
  bool have_range = false;  // optimizer statistics: every key of the next consume calls lies in [key_lo, key_hi]
  int64_t key_lo = 0, key_hi = -1;
  // cross-GPU merge overlapped with the aggregation: called by wdb_agg_consume after it has launched the kernel of
  // index slice [lo, hi) of the direct-addressed table (ops_comm.cu all-reduces that slice on a side stream while the
  // next slice's kernel runs)
  std::function<int(unsigned lo, unsigned hi)> after_slice;
};

namespace wdb {
struct KeyRange { bool known; int64_t lo, hi; };
int needs_for_agg(int agg);
// table whose initialisation is ordered on `stream` (wdb_agg_create initialises on the NULL stream and waits)
int agg_create_on(int device, int64_t expected_groups, int needs, cudaStream_t stream, wdb_agg **out);
// min/max of the key EXPRESSION over the rows (one streaming pass + a host read-back); known = false when there is nothing to scan
int auto_key_range(Device *d, cudaStream_t s, const wdb_col_t *cols, int ncols, const char *key_expr, int64_t n, KeyRange *out);
// make [lo, lo + span) the live direct-addressed side table of t (initialised on `s`)
int dense_prepare(wdb_agg *t, cudaStream_t s, int64_t lo, int64_t span);
// fold every hash-table entry whose key lies in the side table's range into the side table; entries
// outside it are counted in meta[3]'s neighbour (T.meta[4]) so that the caller can detect stale statistics
int agg_hash_to_dense(wdb_agg *t, cudaStream_t s);
// ordered export of a table whose whole result sits in the side table, without any host
// synchronisation: groups go to d_keys / d_vals (at most cap), their number to *d_groups
int agg_export_dense_async(wdb_agg *t, cudaStream_t s, int agg, int order, int32_t *d_keys, float *d_vals, double *d_sums,
                           int64_t *d_counts, double *d_mins, double *d_maxs, int64_t cap, long long *d_groups);
}  // namespace wdb
