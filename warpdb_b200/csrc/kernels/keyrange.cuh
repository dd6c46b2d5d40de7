// keyrange.cuh -- optimizer statistics on demand: min and max of the GROUP BY key EXPRESSION,
// evaluated exactly as the aggregation kernels evaluate it (`int key = <expr>`, src/jit.cpp:200), in
// one streaming pass over the columns the expression reads.  The reference declares TableStats
// (include/csv_loader.hpp:22-37) but never fills them (src/optimizer.cpp:13-17).
// out[0] = min (starts at INT_MAX), out[1] = max (starts at INT_MIN).
// Host-supplied macros: WDB_BLOCK, WDB_UNROLL, WDB_VEC; generated WDB_KEY.
extern "C" __global__ void __launch_bounds__(WDB_BLOCK)
wdb_keyrange(const wdb_cols C, const i64 n, int *__restrict__ out) {
  int lo = 0x7fffffff, hi = (int)0x80000000;
  const i64 nvec = n / WDB_VEC;
  const i64 tile_vecs = (i64)WDB_BLOCK * WDB_UNROLL;
  const i64 ntiles = (nvec + tile_vecs - 1) / tile_vecs;
  for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const i64 v0 = tile * tile_vecs + threadIdx.x;
    wdb_rows R[WDB_UNROLL];
    const bool full = (tile + 1) * tile_vecs <= nvec;
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u)
      if (full || v0 + (i64)u * WDB_BLOCK < nvec) wdb_load_rows(C, (v0 + (i64)u * WDB_BLOCK) * WDB_VEC, R[u]);
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u) {
      if (!(full || v0 + (i64)u * WDB_BLOCK < nvec)) continue;
#pragma unroll
      for (int j = 0; j < WDB_VEC; ++j) {
        const int k = WDB_KEY(R[u], j);
        lo = min(lo, k);
        hi = max(hi, k);
      }
    }
  }
  if (blockIdx.x == 0) {  // ragged tail
    const i64 row = nvec * WDB_VEC + threadIdx.x;
    if (row < n) {
      wdb_rows R;
      wdb_load_row1(C, row, R, 0);
      const int k = WDB_KEY(R, 0);
      lo = min(lo, k);
      hi = max(hi, k);
    }
  }
  lo = __reduce_min_sync(WDB_FULL_MASK, lo);
  hi = __reduce_max_sync(WDB_FULL_MASK, hi);
  __shared__ int s_lo[WDB_BLOCK / 32], s_hi[WDB_BLOCK / 32];
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < WDB_BLOCK / 32; ++w) { lo = min(lo, s_lo[w]); hi = max(hi, s_hi[w]); }
    atomicMin(&out[0], lo);
    atomicMax(&out[1], hi);
  }
}
