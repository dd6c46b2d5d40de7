// group.cuh -- fused filter + hash GROUP BY consume kernel (replaces the <<<1,1>>> linear-probe
// `group_kernel` of src/jit.cpp:192-215 and the std::map loop of src/warpdb.cpp:373-385).
//
// key = (int)KEY(row) exactly as `int key = <expr>` (src/jit.cpp:200); accumulators are fp64
// (src/warpdb.cpp:379-384).  Every CTA pre-aggregates into a private shared-memory hash table
// (WDB_SMEM_SLOTS slots, native 32-bit CAS on the key, CAS-loop fp64 add -- sm_100 has no native
// shared fp64 atomic) and folds it into the global open-addressing table once, at the end; rows
// that do not find a shared slot within WDB_SMEM_PROBES probes go to the global table directly
// (native RED.ADD.F64 in L2).  With WDB_SMEM_SLOTS == 0 every row goes straight to global memory
// (large group counts).
//
// Roofline: nominally HBM (8 B/row for price,quantity) but in practice bound by the shared-memory
// atomic rate (small G) or by L2/DRAM random read-modify-write (G >> L2); see DESIGN.md.
//
// Host-supplied macros: WDB_BLOCK, WDB_UNROLL, WDB_VEC, WDB_NEEDS, WDB_SMEM_SLOTS (0 or power of
// two), WDB_SMEM_LOG2, WDB_SMEM_PROBES, WDB_HAS_COND; generated WDB_VAL, WDB_KEY, WDB_COND.

#if WDB_SMEM_SLOTS > 0
struct wdb_smem_table {
  int *keys;
  double *sums;
  u32 *cnts;
  i64 *mins;
  i64 *maxs;
  i64 *first;
};
__device__ __forceinline__ wdb_smem_table wdb_smem_carve(unsigned char *base) {
  wdb_smem_table S;
  size_t off = 0;
  S.sums = reinterpret_cast<double *>(base + off); off += sizeof(double) * WDB_SMEM_SLOTS;
  S.mins = reinterpret_cast<i64 *>(base + off); off += (WDB_NEEDS & WDB_NEED_MINMAX_BIT) ? sizeof(i64) * WDB_SMEM_SLOTS : 0;
  S.maxs = reinterpret_cast<i64 *>(base + off); off += (WDB_NEEDS & WDB_NEED_MINMAX_BIT) ? sizeof(i64) * WDB_SMEM_SLOTS : 0;
  S.first = reinterpret_cast<i64 *>(base + off); off += (WDB_NEEDS & WDB_NEED_FIRST_BIT) ? sizeof(i64) * WDB_SMEM_SLOTS : 0;
  S.keys = reinterpret_cast<int *>(base + off); off += sizeof(int) * WDB_SMEM_SLOTS;
  S.cnts = reinterpret_cast<u32 *>(base + off);   // present only when counts are needed
  return S;
}
#endif

__device__ __forceinline__ void wdb_group_row(const wdb_table &T,
#if WDB_SMEM_SLOTS > 0
                                              const wdb_smem_table &S,
#endif
                                              int key, float val, i64 row, const u32 pass_bits, const u32 pass) {
  const double dv = (double)val;
  const u32 hsh = wdb_hash32(key);
  // multi-pass mode (large tables): this launch only folds the keys whose hash prefix is `pass`,
  // i.e. one contiguous, L2-sized region of the global table
  if (pass_bits && (hsh >> (32u - pass_bits)) != pass) return;
#if WDB_SMEM_SLOTS > 0
  if (key != WDB_KEY_EMPTY) {
    u32 h = hsh >> (32 - WDB_SMEM_LOG2);
#pragma unroll 1
    for (int p = 0; p < WDB_SMEM_PROBES; ++p) {
      const int k = S.keys[h];
      bool hit = (k == key);
      if (!hit && k == WDB_KEY_EMPTY) {
        const int prev = atomicCAS(&S.keys[h], WDB_KEY_EMPTY, key);
        hit = (prev == WDB_KEY_EMPTY) || (prev == key);
      }
      if (hit) {
        if (WDB_NEEDS & WDB_NEED_SUM_BIT) atomicAdd(&S.sums[h], dv);
        if (WDB_NEEDS & WDB_NEED_CNT_BIT) atomicAdd(&S.cnts[h], 1u);
        if (WDB_NEEDS & WDB_NEED_MINMAX_BIT) { const i64 e = wdb_f64_enc(dv); atomicMin(&S.mins[h], e); atomicMax(&S.maxs[h], e); }
        if (WDB_NEEDS & WDB_NEED_FIRST_BIT) atomicMin(&S.first[h], row);
        return;
      }
      h = (h + 1u) & (WDB_SMEM_SLOTS - 1u);
    }
  }
#endif
  const i64 s = wdb_table_slot_h(T, key, hsh);
  if (s >= 0) { const i64 e = wdb_f64_enc(dv); wdb_table_add<WDB_NEEDS>(T, s, dv, 1ull, e, e, row); }
}

extern "C" __global__ void __launch_bounds__(WDB_BLOCK)
wdb_group(const wdb_cols C, const i64 n, const i64 row_base, const wdb_table T, const u32 pass_bits, const u32 pass) {
#if WDB_SMEM_SLOTS > 0
  extern __shared__ __align__(16) unsigned char wdb_smem[];
  const wdb_smem_table S = wdb_smem_carve(wdb_smem);
  for (int s = threadIdx.x; s < WDB_SMEM_SLOTS; s += WDB_BLOCK) {
    S.keys[s] = WDB_KEY_EMPTY;
    S.sums[s] = 0.0;
    if (WDB_NEEDS & WDB_NEED_CNT_BIT) S.cnts[s] = 0u;
    if (WDB_NEEDS & WDB_NEED_MINMAX_BIT) { S.mins[s] = WDB_ENC_PLUS_INF; S.maxs[s] = WDB_ENC_MINUS_INF; }
    if (WDB_NEEDS & WDB_NEED_FIRST_BIT) S.first[s] = 0x7fffffffffffffffll;
  }
  __syncthreads();
#define WDB_GROUP_ROW(key, val, row) wdb_group_row(T, S, key, val, row, pass_bits, pass)
#else
#define WDB_GROUP_ROW(key, val, row) wdb_group_row(T, key, val, row, pass_bits, pass)
#endif
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
      const i64 row = (v0 + (i64)u * WDB_BLOCK) * WDB_VEC;
#pragma unroll
      for (int j = 0; j < WDB_VEC; ++j) {
#if WDB_HAS_COND
        if (!WDB_COND(R[u], j)) continue;
#endif
        WDB_GROUP_ROW(WDB_KEY(R[u], j), WDB_VAL(R[u], j), row_base + row + j);
      }
    }
  }
  if (blockIdx.x == 0) {  // ragged tail
    const i64 row = nvec * WDB_VEC + threadIdx.x;
    if (row < n) {
      wdb_rows R;
      wdb_load_row1(C, row, R, 0);
#if WDB_HAS_COND
      if (WDB_COND(R, 0))
#endif
        WDB_GROUP_ROW(WDB_KEY(R, 0), WDB_VAL(R, 0), row_base + row);
    }
  }
#if WDB_SMEM_SLOTS > 0
  __syncthreads();
  for (int s = threadIdx.x; s < WDB_SMEM_SLOTS; s += WDB_BLOCK) {
    const int key = S.keys[s];
    if (key == WDB_KEY_EMPTY) continue;
    const i64 g = wdb_table_slot(T, key);
    if (g >= 0)
      wdb_table_add<WDB_NEEDS>(T, g, S.sums[s], (WDB_NEEDS & WDB_NEED_CNT_BIT) ? (u64)S.cnts[s] : 0ull, (WDB_NEEDS & WDB_NEED_MINMAX_BIT) ? S.mins[s] : 0,
                               (WDB_NEEDS & WDB_NEED_MINMAX_BIT) ? S.maxs[s] : 0, (WDB_NEEDS & WDB_NEED_FIRST_BIT) ? S.first[s] : 0);
  }
#endif
}

#if WDB_WP_SLOTS > 0
// ---- small-cardinality kernel: warp-private shared-memory tables, no shared-memory atomics on
// the accumulation path.  sm_100 has no native 64-bit (or floating-point) shared-memory atomic:
// atomicAdd(double*) on shared memory is a CAS loop on the single ATOMS pipe, which caps the
// kernel above at ~0.3 rows/clk/SM.  Here every warp owns a private open-addressing table
// (WDB_WP_SLOTS slots of key + fp64 sum [+ u32 count]); the lanes of a warp that hit the same slot
// in the same step arbitrate through a one-byte tag per slot, and the winner does a plain
// LDS/DADD/STS read-modify-write.  A 32-bit CAS is used only to claim an
// empty slot (once per distinct key per warp).  Tables are folded into the global table at the end.
#define WDB_WP_WARPS (WDB_BLOCK / 32)
struct wdb_wp_table {
  int *keys;
  double *sums;
  u32 *cnts;
  unsigned char *tags;
};
__device__ __noinline__ void wdb_wp_row(const wdb_table &T, const wdb_wp_table &W, const bool valid, const int key,
                                           const float val, const i64 row, const u32 lane) {
  // 1. slot of this lane's key in the warp's table
  u32 h = 0;
  bool in_table = false;
  if (valid && key != WDB_KEY_EMPTY) {
    u32 s = wdb_hash32(key) >> (32 - WDB_WP_LOG2);
#pragma unroll 1
    for (int p = 0; p < WDB_WP_PROBES; ++p) {
      const int k = W.keys[s];
      if (k == key) { in_table = true; break; }
      if (k == WDB_KEY_EMPTY) {
        const int prev = atomicCAS(&W.keys[s], WDB_KEY_EMPTY, key);
        if (prev == WDB_KEY_EMPTY || prev == key) { in_table = true; break; }
      }
      s = (s + 1u) & (WDB_WP_SLOTS - 1u);
    }
    if (in_table) h = s;
  }
  // 2. conflict-free accumulate: lanes that hit the same slot in this step arbitrate through a
  // one-byte tag per slot (write lane id, re-read, the lane that reads back its own id owns the slot
  // for this round); losers go round again.  (MATCH.ANY would find the peers in one instruction
  // but measured ~400 cycles per warp instruction on B200.)
  bool pending = in_table;
  while (__any_sync(WDB_FULL_MASK, pending)) {
    if (pending) W.tags[h] = (unsigned char)lane;
    __syncwarp();
    if (pending && W.tags[h] == (unsigned char)lane) {
      if (WDB_NEEDS & WDB_NEED_SUM_BIT) W.sums[h] += (double)val;
      if (WDB_NEEDS & WDB_NEED_CNT_BIT) W.cnts[h] += 1u;
      pending = false;
    }
    __syncwarp();
  }
  if (!in_table && valid) {  // table full or sentinel key: straight to the global table
    const i64 g = wdb_table_slot(T, key);
    if (g >= 0) wdb_table_add<WDB_NEEDS>(T, g, (double)val, 1ull, 0, 0, row);
  }
  __syncwarp();
}

extern "C" __global__ void __launch_bounds__(WDB_BLOCK, 1)
wdb_group_wp(const wdb_cols C, const i64 n, const i64 row_base, const wdb_table T) {
  extern __shared__ __align__(16) unsigned char wdb_smem[];
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  double *all_sums = reinterpret_cast<double *>(wdb_smem);
  int *all_keys = reinterpret_cast<int *>(wdb_smem + sizeof(double) * WDB_WP_SLOTS * WDB_WP_WARPS);
  unsigned char *all_tags = wdb_smem + (sizeof(double) + sizeof(int)) * WDB_WP_SLOTS * WDB_WP_WARPS;
  u32 *all_cnts = reinterpret_cast<u32 *>(wdb_smem + (sizeof(double) + sizeof(int) + 1) * WDB_WP_SLOTS * WDB_WP_WARPS);
  for (int s = threadIdx.x; s < WDB_WP_SLOTS * WDB_WP_WARPS; s += WDB_BLOCK) {
    all_keys[s] = WDB_KEY_EMPTY;
    all_sums[s] = 0.0;
    if (WDB_NEEDS & WDB_NEED_CNT_BIT) all_cnts[s] = 0u;
  }
  __syncthreads();
  wdb_wp_table W;
  W.keys = all_keys + warp * WDB_WP_SLOTS;
  W.sums = all_sums + warp * WDB_WP_SLOTS;
  W.cnts = all_cnts + warp * WDB_WP_SLOTS;
  W.tags = all_tags + warp * WDB_WP_SLOTS;

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
      const bool have = full || v0 + (i64)u * WDB_BLOCK < nvec;     // warp-uniform except in the last tile
      const i64 row = (v0 + (i64)u * WDB_BLOCK) * WDB_VEC;
#pragma unroll
      for (int j = 0; j < WDB_VEC; ++j) {
        bool valid = have;
        int key = 0;
        float val = 0.0f;
        if (have) {
#if WDB_HAS_COND
          valid = WDB_COND(R[u], j);
#endif
          key = WDB_KEY(R[u], j);
          val = WDB_VAL(R[u], j);
        }
        wdb_wp_row(T, W, valid, key, val, row_base + row + j, lane);
      }
    }
  }
  if (blockIdx.x == 0) {  // ragged tail: one row per thread of CTA 0
    const i64 row = nvec * WDB_VEC + threadIdx.x;
    bool valid = row < n;
    int key = 0;
    float val = 0.0f;
    if (valid) {
      wdb_rows R;
      wdb_load_row1(C, row, R, 0);
#if WDB_HAS_COND
      valid = WDB_COND(R, 0);
#endif
      key = WDB_KEY(R, 0);
      val = WDB_VAL(R, 0);
    }
    wdb_wp_row(T, W, valid, key, val, row_base + row, lane);
  }
  __syncthreads();
  for (int s = threadIdx.x; s < WDB_WP_SLOTS * WDB_WP_WARPS; s += WDB_BLOCK) {
    const int key = all_keys[s];
    if (key == WDB_KEY_EMPTY) continue;
    const i64 g = wdb_table_slot(T, key);
    if (g >= 0) wdb_table_add<WDB_NEEDS>(T, g, all_sums[s], (WDB_NEEDS & WDB_NEED_CNT_BIT) ? (u64)all_cnts[s] : 0ull, 0, 0, 0);
  }
}
#endif
