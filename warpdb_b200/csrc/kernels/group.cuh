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
// atomicAdd(double*) on shared memory is a CAS loop on the ATOMS path (~2 clk per lane), which caps
// `wdb_group` above at ~0.7 rows/clk/SM however the table is shaped.  Plain LDS/STS cost one clock
// per conflict-free warp access, so here every warp owns a private table and updates it with
// ordinary read-modify-writes:
//   slot (16 B) = { f64 sum ; i32 key ; u32 tag },  slot of a key = xor-folded key (dense integer
//   ranges map without collisions; anything else still works, just with more slow-path rows)
//   step (WDB_WP_ILP rows per lane):
//     1. STS.32  every lane writes a unique tag (lane + 32*i) into the home slot of its key
//     2. LDS.128 reads the slot back: {sum, key, tag}
//     3. the lane that finds its own key AND its own tag is the only writer of that slot in this
//        step: STS.64 sum + value (plus a u32 count in a side array when COUNT/AVG need it)
//     4. everybody else (same key twice in one step, key displaced from its home slot, empty
//        slot, table full) takes the slow path AFTER a warp barrier: CAS claim + CAS-loop add in the
//        warp's table, or the global table when WDB_WP_PROBES slots were all taken.
// Column vectors are double-buffered in registers (the next tile's loads are in flight while the
// current one is folded) because a CTA has only a handful of warps.  Tables are folded into the
// global table once, at the end.
#define WDB_WP_WARPS (WDB_BLOCK / 32)
#define WDB_WP_HAS_CNT ((WDB_NEEDS & WDB_NEED_CNT_BIT) != 0)

__device__ __forceinline__ u32 wdb_wp_home(int key) {
  u32 x = (u32)key;
  x ^= x >> WDB_WP_LOG2;
  x ^= x >> ((2 * WDB_WP_LOG2) < 32 ? (2 * WDB_WP_LOG2) : 31);
  return x & (WDB_WP_SLOTS - 1u);
}

// rare rows: atomic path (all plain writers of the step have finished: caller put a warp barrier in between)
__device__ __noinline__ void wdb_wp_slow(const wdb_table &T, uint4 *slots, u32 *cnts, const int key, const float val, const i64 row) {
  const double dv = (double)val;
  if (key != WDB_KEY_EMPTY) {
    u32 s = wdb_wp_home(key);
#pragma unroll 1
    for (int p = 0; p < WDB_WP_PROBES; ++p) {
      int *kp = reinterpret_cast<int *>(&slots[s]) + 2;
      int k = *reinterpret_cast<volatile int *>(kp);
      if (k == WDB_KEY_EMPTY) {
        const int prev = atomicCAS(kp, WDB_KEY_EMPTY, key);
        k = (prev == WDB_KEY_EMPTY) ? key : prev;
      }
      if (k == key) {
        if (WDB_NEEDS & WDB_NEED_SUM_BIT) atomicAdd(reinterpret_cast<double *>(&slots[s]), dv);
        if (WDB_WP_HAS_CNT) atomicAdd(&cnts[s], 1u);
        return;
      }
      s = (s + 1u) & (WDB_WP_SLOTS - 1u);
    }
  }
  const i64 g = wdb_table_slot(T, key);
  if (g >= 0) wdb_table_add<WDB_NEEDS>(T, g, dv, 1ull, 0, 0, row);
}

// one step: NI rows per lane
template <int NI>
__device__ __forceinline__ void wdb_wp_step(const wdb_table &T, uint4 *slots, u32 *cnts, const u32 lane, const int (&key)[NI],
                                            const float (&val)[NI], const bool (&valid)[NI], const i64 row0) {
  u32 h[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    h[i] = wdb_wp_home(key[i]);
    if (valid[i]) reinterpret_cast<u32 *>(&slots[h[i]])[3] = lane + 32u * i;
  }
  __syncwarp();
  uint4 s[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i)
    s[i] = slots[h[i]];
  bool slow[NI];
  bool any_slow = false;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const bool win = valid[i] && s[i].z == (u32)key[i] && s[i].w == lane + 32u * i;
    if (win) {
      if (WDB_NEEDS & WDB_NEED_SUM_BIT)
        *reinterpret_cast<double *>(&slots[h[i]]) = __hiloint2double((int)s[i].y, (int)s[i].x) + (double)val[i];
      if (WDB_WP_HAS_CNT) cnts[h[i]] += 1u;
    }
    slow[i] = valid[i] && !win;
    any_slow |= slow[i];
  }
  __syncwarp();
  if (__any_sync(WDB_FULL_MASK, any_slow)) {
#pragma unroll
    for (int i = 0; i < NI; ++i)
      if (slow[i]) wdb_wp_slow(T, slots, cnts, key[i], val[i], row0 + i);
    __syncwarp();
  }
}

__device__ __forceinline__ void wdb_wp_tile(const wdb_table &T, uint4 *slots, u32 *cnts, const u32 lane, const wdb_rows (&R)[WDB_UNROLL],
                                            const i64 v0, const i64 nvec, const bool full, const i64 row_base) {
#pragma unroll
  for (int u = 0; u < WDB_UNROLL; ++u) {
    const bool have = full || v0 + (i64)u * WDB_BLOCK < nvec;     // warp-uniform except in the last tile
    const i64 row = row_base + (v0 + (i64)u * WDB_BLOCK) * WDB_VEC;
#pragma unroll
    for (int j0 = 0; j0 < WDB_VEC; j0 += WDB_WP_ILP) {
      int key[WDB_WP_ILP];
      float val[WDB_WP_ILP];
      bool valid[WDB_WP_ILP];
#pragma unroll
      for (int i = 0; i < WDB_WP_ILP; ++i) {
        valid[i] = have;
        key[i] = 0;
        val[i] = 0.0f;
        if (have) {
#if WDB_HAS_COND
          valid[i] = WDB_COND(R[u], j0 + i);
#endif
          key[i] = WDB_KEY(R[u], j0 + i);
          val[i] = WDB_VAL(R[u], j0 + i);
        }
      }
      wdb_wp_step<WDB_WP_ILP>(T, slots, cnts, lane, key, val, valid, row + j0);
    }
  }
}

__device__ __forceinline__ void wdb_wp_load(const wdb_cols &C, wdb_rows (&R)[WDB_UNROLL], const i64 v0, const i64 nvec, const bool full) {
#pragma unroll
  for (int u = 0; u < WDB_UNROLL; ++u)
    if (full || v0 + (i64)u * WDB_BLOCK < nvec) wdb_load_rows(C, (v0 + (i64)u * WDB_BLOCK) * WDB_VEC, R[u]);
}

extern "C" __global__ void __launch_bounds__(WDB_BLOCK, 1)
wdb_group_wp(const wdb_cols C, const i64 n, const i64 row_base, const wdb_table T) {
  extern __shared__ __align__(16) unsigned char wdb_smem[];
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint4 *all_slots = reinterpret_cast<uint4 *>(wdb_smem);
  u32 *all_cnts = reinterpret_cast<u32 *>(wdb_smem + sizeof(uint4) * WDB_WP_SLOTS * WDB_WP_WARPS);
  for (int s = threadIdx.x; s < WDB_WP_SLOTS * WDB_WP_WARPS; s += WDB_BLOCK) {
    all_slots[s] = make_uint4(0u, 0u, (u32)WDB_KEY_EMPTY, 0xffffffffu);
    if (WDB_WP_HAS_CNT) all_cnts[s] = 0u;
  }
  __syncthreads();
  uint4 *slots = all_slots + warp * WDB_WP_SLOTS;
  u32 *cnts = all_cnts + warp * WDB_WP_SLOTS;

  const i64 nvec = n / WDB_VEC;
  const i64 tile_vecs = (i64)WDB_BLOCK * WDB_UNROLL;
  const i64 ntiles = (nvec + tile_vecs - 1) / tile_vecs;
  const i64 stride = gridDim.x;
  i64 tile = blockIdx.x;
  wdb_rows A[WDB_UNROLL], B[WDB_UNROLL];
  if (tile < ntiles) wdb_wp_load(C, A, tile * tile_vecs + threadIdx.x, nvec, (tile + 1) * tile_vecs <= nvec);
  while (tile < ntiles) {
    const i64 t1 = tile + stride;
    if (t1 < ntiles) wdb_wp_load(C, B, t1 * tile_vecs + threadIdx.x, nvec, (t1 + 1) * tile_vecs <= nvec);
    wdb_wp_tile(T, slots, cnts, lane, A, tile * tile_vecs + threadIdx.x, nvec, (tile + 1) * tile_vecs <= nvec, row_base);
    if (t1 >= ntiles) break;
    const i64 t2 = t1 + stride;
    if (t2 < ntiles) wdb_wp_load(C, A, t2 * tile_vecs + threadIdx.x, nvec, (t2 + 1) * tile_vecs <= nvec);
    wdb_wp_tile(T, slots, cnts, lane, B, t1 * tile_vecs + threadIdx.x, nvec, (t1 + 1) * tile_vecs <= nvec, row_base);
    tile = t2;
  }
  if (blockIdx.x == 0) {  // ragged tail: one row per thread of CTA 0
    const i64 row = nvec * WDB_VEC + threadIdx.x;
    bool valid[1] = {row < n};
    int key[1] = {0};
    float val[1] = {0.0f};
    if (valid[0]) {
      wdb_rows R;
      wdb_load_row1(C, row, R, 0);
#if WDB_HAS_COND
      valid[0] = WDB_COND(R, 0);
#endif
      key[0] = WDB_KEY(R, 0);
      val[0] = WDB_VAL(R, 0);
    }
    wdb_wp_step<1>(T, slots, cnts, lane, key, val, valid, row_base + row);
  }
  __syncthreads();
  for (int s = threadIdx.x; s < WDB_WP_SLOTS * WDB_WP_WARPS; s += WDB_BLOCK) {
    const uint4 v = all_slots[s];
    if ((int)v.z == WDB_KEY_EMPTY) continue;
    const i64 g = wdb_table_slot(T, (int)v.z);
    if (g >= 0) wdb_table_add<WDB_NEEDS>(T, g, __hiloint2double((int)v.y, (int)v.x), WDB_WP_HAS_CNT ? (u64)all_cnts[s] : 0ull, 0, 0, 0);
  }
}
#endif
