// group.cuh -- fused filter + GROUP BY consume kernels (replace the <<<1,1>>> linear-probe
// `group_kernel` of src/jit.cpp:192-215 and the std::map loop of src/warpdb.cpp:373-385).
//
// key = (int)KEY(row) exactly as `int key = <expr>` (src/jit.cpp:200); accumulators are fp64
// (src/warpdb.cpp:379-384).  Two kernels, four accumulator layouts (the host picks one from the
// key's min/max statistics, ops_group.cu):
//   wdb_group     every CTA pre-aggregates into a private shared-memory hash table
//                 (WDB_SMEM_SLOTS slots, native 32-bit CAS on the key, CAS-loop fp64 add -- sm_100 has
//                 no native shared fp64 atomic) and folds it into the global open-addressing table
//                 once, at the end; rows that do not find a shared slot within WDB_SMEM_PROBES probes
//                 go to the global table directly (native RED.ADD.F64 in L2).  WDB_SMEM_SLOTS == 0:
//                 every row goes straight to global memory (large group counts).  WDB_DENSE: keys
//                 inside the known range go to a direct-addressed table instead (one RED, no probe).
//   wdb_group_wp  narrow key ranges: warp-private, directly indexed accumulators in shared memory
//                 updated with plain read-modify-writes (no shared-memory atomics); see below.
//
// Roofline: nominally HBM (8 B/row for price,quantity) but in practice bound by shared-memory
// wavefronts (wdb_group_wp), the shared-memory atomic rate (wdb_group, small G) or L2 atomics /
// DRAM random read-modify-write (large G); see DESIGN.md section 3.4.
//
// Host-supplied macros: WDB_BLOCK, WDB_UNROLL, WDB_VEC, WDB_NEEDS, WDB_SMEM_SLOTS (0 or power of
// two), WDB_SMEM_LOG2, WDB_SMEM_PROBES, WDB_HAS_COND, WDB_DENSE, WDB_WP_IDS (0: no wdb_group_wp),
// WDB_WP_ILP; generated WDB_VAL, WDB_KEY, WDB_COND.

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
#if WDB_DENSE
  {  // direct-addressed table: one native RED per accumulator, no probe (keys outside the promised range fall through)
    // a table larger than the L2 is filled in several launches, launch i folding only the index
    // slice [pass_bits, pass) (the two parameters carry the slice bounds in this mode), so that the
    // random read-modify-writes of a launch stay L2-resident; the launch whose slice starts at 0
    // also handles the keys outside the range
    const u32 di = (u32)key - (u32)T.dlo;
    if (di < T.dspan) {
      if (di >= pass_bits && di < pass) {
        const i64 e = (WDB_NEEDS & WDB_NEED_MINMAX_BIT) ? wdb_f64_enc(dv) : 0;
        wdb_dense_add<WDB_NEEDS & 7>(T, di, dv, 1ull, e, e);
      }
      return;
    }
    if (pass_bits != 0u) return;
  }
  const u32 hsh = wdb_hash32(key);
#else
  const u32 hsh = wdb_hash32(key);
  // multi-pass mode (large tables): this launch only folds the keys whose hash prefix is `pass`,
  // i.e. one contiguous, L2-sized region of the global table
  if (pass_bits && (hsh >> (32u - pass_bits)) != pass) return;
#endif
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

#if WDB_WP_IDS > 0
// ---- small-cardinality kernel for keys with a known, narrow range: warp-private, directly indexed
// accumulators and NO shared-memory atomics.  sm_100 has no native 64-bit (or floating-point)
// shared-memory atomic: atomicAdd(double*) on shared memory is a CAS loop on the ATOMS path (~2 clk
// per lane), which caps `wdb_group` above at ~0.7 rows/clk/SM however the table is shaped.  Plain
// LDS/STS cost one clock per conflict-free wavefront, so here every warp owns private accumulator
// arrays and updates them with ordinary read-modify-writes.
//   group id of a key:   key - key_base; the optimizer knows the key column's min/max
//                        (wdb_agg_set_key_range / wdb_column_minmax).  Keys outside the promised
//                        range stay correct: they are folded into the global table directly.
//   accumulators:        struct of arrays per warp -- f64 sums[WDB_WP_IDS], i64 mins[] / maxs[], u32 tags[], u32 counts[]
//                        (an array of 16-byte structs would put every tag in one of 8 banks)
//   step (WDB_WP_ILP rows per lane), repeated while any lane is pending:
//     1. STS.32  every pending lane writes a unique tag (lane + 32*i) for its id
//     2. LDS.32 + LDS.64  read the tag back, and the sum with it
//     3. the lane that reads its own tag is the only writer of that id in this round:
//        DADD / STS.64 on the sum (and count); lanes that lost (same key twice in one round)
//        stay pending
//   Control flow is warp-uniform (votes decide every loop; calls into non-inlined functions are
//   followed by a warp barrier).
// Measured (ncu, 1 K keys): 21 shared-memory wavefronts per 32 rows (13 of them bank conflicts of
// the random accesses), the shared-memory pipe 81 % busy: 375 Grows/s against 202 Grows/s of the
// atomic kernel.  A variant that mapped arbitrary keys to ids through a CTA-shared hash index cost
// another ~10 wavefronts and a dependent LDS per row and lost to the atomic kernel (140-170 Grows/s).
// Column vectors are double-buffered in registers (the next tile's loads are in flight while the
// current one is folded) because a CTA has only a handful of warps.  The warps' accumulators are
// summed in shared memory and folded into the global table once per CTA, at the end.
#define WDB_WP_WARPS (WDB_BLOCK / 32)
#define WDB_WP_HAS_SUM ((WDB_NEEDS & WDB_NEED_SUM_BIT) != 0)
#define WDB_WP_HAS_CNT ((WDB_NEEDS & WDB_NEED_CNT_BIT) != 0)
#define WDB_WP_HAS_MM ((WDB_NEEDS & WDB_NEED_MINMAX_BIT) != 0)
#define WDB_WP_HAS_FIRST ((WDB_NEEDS & WDB_NEED_FIRST_BIT) != 0)
#define WDB_WP_NOID 0xffffffffu

// All of the kernel's shared memory is addressed as offsets from this one symbol so that every
// access stays an LDS/STS (pointers kept in a struct decay to generic LD/ST once the struct is
// passed to a non-inlined function).
// layout (each array [WARPS][IDS], present only when the aggregation needs it):
//   sums f64 | mins i64 | maxs i64 (order-preserving encodings of f64) | first i64 (smallest row id) | tags u32 | counts u32
extern __shared__ __align__(16) unsigned char wdb_wp_smem[];
struct wdb_wp_state {
  u32 sums, mins, maxs, first, tags, cnts;   // byte offsets of this warp's accumulator arrays
  int key_base;
};
// accumulator entries per warp: one per id, or -- WDB_WP_MODE == 2, tiny key ranges -- one per (id, lane)
#define WDB_WP_LANES (WDB_WP_MODE == 2 ? 32 : 1)
#define WDB_WP_SLOTS (WDB_WP_IDS * WDB_WP_LANES)
#define WDB_WP_OFF_SUMS 0u
#define WDB_WP_OFF_MINS (WDB_WP_OFF_SUMS + (WDB_WP_HAS_SUM ? 8u : 0u) * WDB_WP_SLOTS * WDB_WP_WARPS)
#define WDB_WP_OFF_MAXS (WDB_WP_OFF_MINS + (WDB_WP_HAS_MM ? 8u : 0u) * WDB_WP_SLOTS * WDB_WP_WARPS)
#define WDB_WP_OFF_FIRST (WDB_WP_OFF_MAXS + (WDB_WP_HAS_MM ? 8u : 0u) * WDB_WP_SLOTS * WDB_WP_WARPS)
#define WDB_WP_OFF_TAGS (WDB_WP_OFF_FIRST + (WDB_WP_HAS_FIRST ? 8u : 0u) * WDB_WP_SLOTS * WDB_WP_WARPS)
#define WDB_WP_OFF_CNTS (WDB_WP_OFF_TAGS + (WDB_WP_MODE >= 1 ? 0u : 4u) * WDB_WP_SLOTS * WDB_WP_WARPS)
#define WDB_WP_MIN(W, id) (*reinterpret_cast<i64 *>(wdb_wp_smem + (W).mins + 8u * (id)))
#define WDB_WP_MAX(W, id) (*reinterpret_cast<i64 *>(wdb_wp_smem + (W).maxs + 8u * (id)))
#define WDB_WP_FIRST(W, id) (*reinterpret_cast<i64 *>(wdb_wp_smem + (W).first + 8u * (id)))
#define WDB_WP_SUM(W, id) (*reinterpret_cast<double *>(wdb_wp_smem + (W).sums + 8u * (id)))
#define WDB_WP_TAG(W, id) (*reinterpret_cast<u32 *>(wdb_wp_smem + (W).tags + 4u * (id)))
#define WDB_WP_CNT(W, id) (*reinterpret_cast<u32 *>(wdb_wp_smem + (W).cnts + 4u * (id)))

__device__ __noinline__ void wdb_wp_global_row(const wdb_table &T, const int key, const float val, const i64 row) {
  atomicAdd(&T.meta[3], 1u);   // statistics: rows that bypassed the shared-memory accumulators (wdb_agg_spilled)
  const i64 g = wdb_table_slot(T, key);
  const i64 e = wdb_f64_enc((double)val);
  if (g >= 0) wdb_table_add<WDB_NEEDS>(T, g, (double)val, 1ull, e, e, row);
}

// one step: NI rows per lane
template <int NI>
__device__ __forceinline__ void wdb_wp_step(const wdb_table &T, const wdb_wp_state &W, const u32 lane, const int (&key)[NI],
                                            const float (&val)[NI], const bool (&valid)[NI], const i64 row0) {
  u32 id[NI];
  bool pending[NI];
  bool any_global = false, any_pending = false;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    id[i] = (u32)key[i] - (u32)W.key_base;
    pending[i] = valid[i] && id[i] < (u32)WDB_WP_IDS;
    any_pending |= pending[i];
    any_global |= valid[i] && !pending[i];
  }
  if (__any_sync(WDB_FULL_MASK, any_global)) {   // stale statistics: a key outside the promised range
#pragma unroll
    for (int i = 0; i < NI; ++i)
      if (valid[i] && !pending[i]) wdb_wp_global_row(T, key[i], val[i], row0 + i);
    __syncwarp();
  }
#pragma unroll 1
  while (__any_sync(WDB_FULL_MASK, any_pending)) {
#pragma unroll
    for (int i = 0; i < NI; ++i)
      if (pending[i]) WDB_WP_TAG(W, id[i]) = lane + 32u * i;
    __syncwarp();
    u32 t[NI], c[NI];
    double a[NI];
    i64 mn[NI], mx[NI], fr[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i)
      if (pending[i]) {       // tag and accumulators are read together: they do not wait for the tag compare
        t[i] = WDB_WP_TAG(W, id[i]);
        if (WDB_WP_HAS_SUM) a[i] = WDB_WP_SUM(W, id[i]);
        if (WDB_WP_HAS_CNT) c[i] = WDB_WP_CNT(W, id[i]);
        if (WDB_WP_HAS_MM) { mn[i] = WDB_WP_MIN(W, id[i]); mx[i] = WDB_WP_MAX(W, id[i]); }
        if (WDB_WP_HAS_FIRST) fr[i] = WDB_WP_FIRST(W, id[i]);
      }
    any_pending = false;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      if (pending[i] && t[i] == lane + 32u * i) {
        if (WDB_WP_HAS_SUM) WDB_WP_SUM(W, id[i]) = a[i] + (double)val[i];
        if (WDB_WP_HAS_CNT) WDB_WP_CNT(W, id[i]) = c[i] + 1u;
        if (WDB_WP_HAS_MM) {   // extrema settle quickly: the stores are rare
          const i64 e = wdb_f64_enc((double)val[i]);
          if (e < mn[i]) WDB_WP_MIN(W, id[i]) = e;
          if (e > mx[i]) WDB_WP_MAX(W, id[i]) = e;
        }
        if (WDB_WP_HAS_FIRST && row0 + i < fr[i]) WDB_WP_FIRST(W, id[i]) = row0 + i;   // first appearance = smallest row id
        pending[i] = false;
      }
      any_pending |= pending[i];
    }
    __syncwarp();
  }
}

#if WDB_WP_MODE == 1
// Leader aggregation: the lanes of a warp that hit the same id in this step find each other with
// MATCH.ANY; the lowest lane of every peer group (the leader) collects the group's values with
// shuffles and is the only one to touch shared memory -- one LDS.64 + STS.64 per distinct id and no
// tag round-trip (half the shared-memory wavefronts of the arbitration above, and no second round).
// Untouched accumulators are recognised by their initial values (-0.0 sums, zero counts, +inf / -inf
// extrema, MAX first row), so this mode keeps no tag array at all.
template <int NI>
__device__ __forceinline__ void wdb_wp_step_match(const wdb_table &T, const wdb_wp_state &W, const u32 lane, const int (&key)[NI],
                                                  const float (&val)[NI], const bool (&valid)[NI], const i64 row0) {
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const u32 id = (u32)key[i] - (u32)W.key_base;
    const bool mine = valid[i] && id < (u32)WDB_WP_IDS;
    if (__any_sync(WDB_FULL_MASK, valid[i] && !mine)) {   // stale statistics: a key outside the promised range
      if (valid[i] && !mine) wdb_wp_global_row(T, key[i], val[i], row0 + i);
      __syncwarp();
    }
    const u32 peers = __match_any_sync(WDB_FULL_MASK, mine ? id : (0x80000000u | lane));
    const bool leader = mine && (peers & wdb_lanemask_lt()) == 0u;
    u32 rest = leader ? (peers & (peers - 1u)) : 0u;        // the peers above the leader
    double acc = (double)val[i];
    u32 cnt = 1u;
    i64 emin = WDB_WP_HAS_MM ? wdb_f64_enc((double)val[i]) : 0, emax = emin;
#pragma unroll 1
    while (__any_sync(WDB_FULL_MASK, rest != 0u)) {         // rare: two rows of one warp step share a key
      const int src = rest ? (__ffs(rest) - 1) : (int)lane;
      const float v = __shfl_sync(WDB_FULL_MASK, val[i], src);
      if (rest) {
        acc += (double)v;
        ++cnt;
        if (WDB_WP_HAS_MM) { const i64 e = wdb_f64_enc((double)v); emin = min(emin, e); emax = max(emax, e); }
        rest &= rest - 1u;
      }
    }
    if (leader) {
      if (WDB_WP_HAS_SUM) WDB_WP_SUM(W, id) += acc + 0.0;
      if (WDB_WP_HAS_CNT) WDB_WP_CNT(W, id) += cnt;
      if (WDB_WP_HAS_MM) {
        if (emin < WDB_WP_MIN(W, id)) WDB_WP_MIN(W, id) = emin;
        if (emax > WDB_WP_MAX(W, id)) WDB_WP_MAX(W, id) = emax;
      }
      if (WDB_WP_HAS_FIRST && row0 + i < WDB_WP_FIRST(W, id)) WDB_WP_FIRST(W, id) = row0 + i;   // the leader is the lowest lane = the smallest row
    }
    __syncwarp();
  }
}
#define WDB_WP_STEP wdb_wp_step_match
#elif WDB_WP_MODE == 2
// Tiny key ranges (a few dozen ids: GROUP BY status / region / weekday): every LANE owns a private copy of every
// accumulator, entry id * 32 + lane, i.e. bank = lane -- no bank conflicts, no duplicates to arbitrate, no tags,
// no votes: a row is one LDS.64 + DADD + STS.64 at two wavefronts each (4 per 32 rows against the 21 of the
// tag arbitration, which at 10 keys also needs ~4 rounds per step because most lanes collide: 6.1 ms per 1e9 rows).
template <int NI>
__device__ __forceinline__ void wdb_wp_step_lane(const wdb_table &T, const wdb_wp_state &W, const u32 lane, const int (&key)[NI],
                                                 const float (&val)[NI], const bool (&valid)[NI], const i64 row0) {
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const u32 id = (u32)key[i] - (u32)W.key_base;
    const bool mine = valid[i] && id < (u32)WDB_WP_IDS;
    if (__any_sync(WDB_FULL_MASK, valid[i] && !mine)) {   // stale statistics: a key outside the promised range
      if (valid[i] && !mine) wdb_wp_global_row(T, key[i], val[i], row0 + i);
      __syncwarp();
    }
    if (mine) {
      const u32 s = id * 32u + lane;
      if (WDB_WP_HAS_SUM) WDB_WP_SUM(W, s) += (double)val[i] + 0.0;
      if (WDB_WP_HAS_CNT) WDB_WP_CNT(W, s) += 1u;
      if (WDB_WP_HAS_MM) {
        const i64 e = wdb_f64_enc((double)val[i]);
        if (e < WDB_WP_MIN(W, s)) WDB_WP_MIN(W, s) = e;
        if (e > WDB_WP_MAX(W, s)) WDB_WP_MAX(W, s) = e;
      }
      if (WDB_WP_HAS_FIRST && row0 + i < WDB_WP_FIRST(W, s)) WDB_WP_FIRST(W, s) = row0 + i;
    }
  }
}
#define WDB_WP_STEP wdb_wp_step_lane
#else
#define WDB_WP_STEP wdb_wp_step
#endif

__device__ __forceinline__ void wdb_wp_tile(const wdb_table &T, const wdb_wp_state &W, const u32 lane, const wdb_rows (&R)[WDB_UNROLL],
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
      WDB_WP_STEP<WDB_WP_ILP>(T, W, lane, key, val, valid, row + j0);
    }
  }
}

__device__ __forceinline__ void wdb_wp_load(const wdb_cols &C, wdb_rows (&R)[WDB_UNROLL], const i64 v0, const i64 nvec, const bool full) {
#pragma unroll
  for (int u = 0; u < WDB_UNROLL; ++u)
    if (full || v0 + (i64)u * WDB_BLOCK < nvec) wdb_load_rows(C, (v0 + (i64)u * WDB_BLOCK) * WDB_VEC, R[u]);
}

extern "C" __global__ void __launch_bounds__(WDB_BLOCK, 1)
wdb_group_wp(const wdb_cols C, const i64 n, const i64 row_base, const wdb_table T, const int key_base) {
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  double *all_sums = reinterpret_cast<double *>(wdb_wp_smem + WDB_WP_OFF_SUMS);
  i64 *all_mins = reinterpret_cast<i64 *>(wdb_wp_smem + WDB_WP_OFF_MINS);
  i64 *all_maxs = reinterpret_cast<i64 *>(wdb_wp_smem + WDB_WP_OFF_MAXS);
  i64 *all_first = reinterpret_cast<i64 *>(wdb_wp_smem + WDB_WP_OFF_FIRST);
  u32 *all_tags = reinterpret_cast<u32 *>(wdb_wp_smem + WDB_WP_OFF_TAGS);
  u32 *all_cnts = reinterpret_cast<u32 *>(wdb_wp_smem + WDB_WP_OFF_CNTS);
  wdb_wp_state W;
  W.sums = WDB_WP_OFF_SUMS + 8u * WDB_WP_SLOTS * warp;
  W.mins = WDB_WP_OFF_MINS + 8u * WDB_WP_SLOTS * warp;
  W.maxs = WDB_WP_OFF_MAXS + 8u * WDB_WP_SLOTS * warp;
  W.first = WDB_WP_OFF_FIRST + 8u * WDB_WP_SLOTS * warp;
  W.tags = WDB_WP_OFF_TAGS + 4u * WDB_WP_SLOTS * warp;
  W.cnts = WDB_WP_OFF_CNTS + 4u * WDB_WP_SLOTS * warp;
  W.key_base = key_base;
  for (int s = threadIdx.x; s < WDB_WP_SLOTS * WDB_WP_WARPS; s += WDB_BLOCK) {
    if (WDB_WP_HAS_SUM) all_sums[s] = WDB_WP_MODE >= 1 ? -0.0 : 0.0;
    if (WDB_WP_HAS_MM) { all_mins[s] = WDB_ENC_PLUS_INF; all_maxs[s] = WDB_ENC_MINUS_INF; }
    if (WDB_WP_HAS_FIRST) all_first[s] = 0x7fffffffffffffffll;
    if (WDB_WP_MODE == 0) all_tags[s] = WDB_WP_NOID;
    if (WDB_WP_HAS_CNT) all_cnts[s] = 0u;
  }
  __syncthreads();

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
    wdb_wp_tile(T, W, lane, A, tile * tile_vecs + threadIdx.x, nvec, (tile + 1) * tile_vecs <= nvec, row_base);
    if (t1 >= ntiles) break;
    const i64 t2 = t1 + stride;
    if (t2 < ntiles) wdb_wp_load(C, A, t2 * tile_vecs + threadIdx.x, nvec, (t2 + 1) * tile_vecs <= nvec);
    wdb_wp_tile(T, W, lane, B, t1 * tile_vecs + threadIdx.x, nvec, (t1 + 1) * tile_vecs <= nvec, row_base);
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
    WDB_WP_STEP<1>(T, W, lane, key, val, valid, row_base + row);
  }
  __syncthreads();
  // fold: one thread per id sums the warps' accumulators (fixed order) and adds the total to the global table
  for (u32 id = threadIdx.x; id < (u32)WDB_WP_IDS; id += WDB_BLOCK) {
    double sum = 0.0;
    u64 cnt = 0ull;
    i64 mn = WDB_ENC_PLUS_INF, mx = WDB_ENC_MINUS_INF, fr = 0x7fffffffffffffffll;
    bool touched = false;
#pragma unroll 1
    for (int w = 0; w < WDB_WP_WARPS; ++w)
#pragma unroll 1
      for (int l = 0; l < WDB_WP_LANES; ++l) {
        const int e = w * WDB_WP_SLOTS + (int)id * WDB_WP_LANES + l;
#if WDB_WP_MODE >= 1
        bool hit;   // every row updates every accumulator the table tracks: any one of them tells
        if (WDB_WP_HAS_CNT) hit = all_cnts[e] != 0u;
        else if (WDB_WP_HAS_SUM) hit = (u64)__double_as_longlong(all_sums[e]) != WDB_DENSE_EMPTY;
        else if (WDB_WP_HAS_MM) hit = all_mins[e] != WDB_ENC_PLUS_INF || all_maxs[e] != WDB_ENC_MINUS_INF;
        else hit = all_first[e] != 0x7fffffffffffffffll;
        if (!hit) continue;
#else
        if (all_tags[e] == WDB_WP_NOID) continue;
#endif
        touched = true;
        if (WDB_WP_HAS_SUM) sum += all_sums[e];
        if (WDB_WP_HAS_CNT) cnt += all_cnts[e];
        if (WDB_WP_HAS_MM) { mn = min(mn, all_mins[e]); mx = max(mx, all_maxs[e]); }
        if (WDB_WP_HAS_FIRST) fr = min(fr, all_first[e]);
      }
    if (!touched) continue;
    const int key = (int)((u32)key_base + id);
#if WDB_DENSE
    const u32 di = (u32)key - (u32)T.dlo;     // the host made the direct-addressed side table cover [key_base, key_base + WDB_WP_IDS)
    if (di < T.dspan && !WDB_WP_HAS_FIRST) {   // the side table holds sums, counts and extrema
      wdb_dense_add<WDB_NEEDS & 7>(T, di, sum, cnt, mn, mx);
      continue;
    }
#endif
    const i64 g = wdb_table_slot(T, key);
    if (g >= 0) wdb_table_add<WDB_NEEDS>(T, g, sum, cnt, mn, mx, fr);
  }
}
#endif
