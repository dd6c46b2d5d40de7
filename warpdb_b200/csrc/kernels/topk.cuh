// topk.cuh -- ORDER BY key [ASC|DESC] LIMIT k kernels (replace jit_sort_float's <<<1,1>>> O(N^2)
// bubble sort of the whole projected column followed by a host-side truncate:
// src/jit.cpp:283-307, src/warpdb.cpp:453-455,483-495).
//
//  wdb_topk_scan   one streaming pass; every thread keeps its WDB_K best (key,row) pairs in
//                  registers (the common case per row is a single compare against its current
//                  worst), lanes merge with warp shuffles, warps merge through shared memory and
//                  each CTA writes WDB_K candidates.
//  wdb_topk_final  the same merge over the CTA candidates (one CTA).
//  wdb_topk_emit   evaluates the SELECT expression at the winning rows.
//  wdb_tile_best   (large k) best key per tile, used to derive a threshold for the compaction pass.
//
// Order: DESC -> larger key first, ASC -> smaller key first; equal keys keep row order (the
// reference's bubble sort is stable).  Roofline: HBM, 4 B/row for a single float key column.
//
// Host-supplied macros: WDB_BLOCK, WDB_UNROLL, WDB_VEC, WDB_K, WDB_DESC, WDB_HAS_COND; generated
// WDB_KEY (float), WDB_VAL (float), WDB_COND.
#define WDB_ROW_NONE 0x7fffffffffffffffll
#if WDB_DESC
#define WDB_KEY_WORST (__int_as_float(0xff800000))   // -inf
#else
#define WDB_KEY_WORST (__int_as_float(0x7f800000))   // +inf
#endif

__device__ __forceinline__ bool wdb_better(float ak, i64 ar, float bk, i64 br) {
#if WDB_DESC
  return (ak > bk) || (ak == bk && ar < br);
#else
  return (ak < bk) || (ak == bk && ar < br);
#endif
}

struct wdb_list {
  float k[WDB_K];
  i64 r[WDB_K];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < WDB_K; ++i) { k[i] = WDB_KEY_WORST; r[i] = WDB_ROW_NONE; }
  }
  __device__ __forceinline__ void offer(float ck, i64 cr) {
    if (wdb_better(ck, cr, k[WDB_K - 1], r[WDB_K - 1])) {
      k[WDB_K - 1] = ck; r[WDB_K - 1] = cr;
#pragma unroll
      for (int i = WDB_K - 1; i > 0; --i)
        if (wdb_better(k[i], r[i], k[i - 1], r[i - 1])) {
          const float tk = k[i]; k[i] = k[i - 1]; k[i - 1] = tk;
          const i64 tr = r[i]; r[i] = r[i - 1]; r[i - 1] = tr;
        }
    }
  }
  __device__ __forceinline__ void pop() {
#pragma unroll
    for (int i = 0; i + 1 < WDB_K; ++i) { k[i] = k[i + 1]; r[i] = r[i + 1]; }
    k[WDB_K - 1] = WDB_KEY_WORST; r[WDB_K - 1] = WDB_ROW_NONE;
  }
};

// Merge the 32 sorted lists of a warp: WDB_K rounds of a shuffle arg-best over the list heads;
// lane i < WDB_K ends up holding the i-th best pair of the warp.
__device__ __forceinline__ void wdb_warp_merge(wdb_list &L, float &rk, i64 &rr) {
  const u32 lane = wdb_lane();
  rk = WDB_KEY_WORST; rr = WDB_ROW_NONE;
#pragma unroll 1
  for (int round = 0; round < WDB_K; ++round) {
    float bk = L.k[0];
    i64 br = L.r[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ok = __shfl_xor_sync(WDB_FULL_MASK, bk, o);
      const i64 orow = __shfl_xor_sync(WDB_FULL_MASK, br, o);
      if (wdb_better(ok, orow, bk, br)) { bk = ok; br = orow; }
    }
    if (br != WDB_ROW_NONE && L.r[0] == br) L.pop();
    if ((int)lane == round) { rk = bk; rr = br; }
  }
}

// Block-wide merge; the first WDB_K lanes of warp 0 write the CTA's best pairs to out_k/out_r.
__device__ __forceinline__ void wdb_block_topk(wdb_list &L, float *__restrict__ out_k, i64 *__restrict__ out_r) {
  __shared__ float s_k[(WDB_BLOCK / 32) * WDB_K];
  __shared__ i64 s_r[(WDB_BLOCK / 32) * WDB_K];
  const u32 lane = wdb_lane(), warp = threadIdx.x >> 5;
  float rk; i64 rr;
  wdb_warp_merge(L, rk, rr);
  if (lane < WDB_K) { s_k[warp * WDB_K + lane] = rk; s_r[warp * WDB_K + lane] = rr; }
  __syncthreads();
  if (warp == 0) {
    L.clear();
    for (int i = lane; i < (WDB_BLOCK / 32) * WDB_K; i += 32) L.offer(s_k[i], s_r[i]);
    wdb_warp_merge(L, rk, rr);
    if (lane < WDB_K) { out_k[lane] = rk; out_r[lane] = rr; }
  }
}

// the winners' SELECT values: thread i < WDB_K evaluates VAL at row best_r[i] (rows are global ids)
__device__ __forceinline__ void wdb_topk_emit_rows(const wdb_cols &C, const i64 row_base, const float *best_k, const i64 *best_r, const int offset,
                                                   float *__restrict__ out_vals, float *__restrict__ out_keys, i64 *__restrict__ out_count) {
  const int i = threadIdx.x;
  int valid = 0;
  if (i < WDB_K) {
    const i64 r = best_r[i];
    valid = (r != WDB_ROW_NONE) ? 1 : 0;
    if (valid && i >= offset) {
      wdb_rows R;
      wdb_load_row1(C, r - row_base, R, 0);
      if (out_vals) out_vals[i - offset] = WDB_VAL(R, 0);
      if (out_keys) out_keys[i - offset] = best_k[i];
    }
  }
  const int total = __syncthreads_count(valid);
  if (i == 0) *out_count = total > offset ? total - offset : 0;
}

// WDB_FUSED_TAIL: the last CTA to finish (a done-counter in global memory) selects among all CTAs' candidates
// and evaluates the SELECT expression at the winners, so one launch does the work of scan + final + emit:
// on a 1e9-row shard the two extra launches and their gaps were ~0.1 ms of a 0.7 ms step.
extern "C" __global__ void __launch_bounds__(WDB_BLOCK)
wdb_topk_scan(const wdb_cols C, const i64 n, const i64 row_base, float *__restrict__ cand_k, i64 *__restrict__ cand_r,
              const float *__restrict__ tau0   // optional: the K-th best key of a sample of the rows (a pre-pass of this kernel)
#if WDB_FUSED_TAIL
              , u32 *__restrict__ done, float *__restrict__ best_k, i64 *__restrict__ best_r, const int offset, float *__restrict__ out_vals,
              float *__restrict__ out_keys, i64 *__restrict__ out_count
#endif
) {
  wdb_list L;
  L.clear();
  // warp-uniform threshold: no row worse than it can reach the result.  It starts from the K-th best key of a sample
  // (K rows beat it already; the host runs this kernel over the first 2^20 rows first) and tightens with the lanes' lists.
  float tau = tau0 ? __ldg(tau0) : WDB_KEY_WORST;
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
      const bool have = full || v0 + (i64)u * WDB_BLOCK < nvec;       // warp-uniform except in the last tile
      const i64 row = (v0 + (i64)u * WDB_BLOCK) * WDB_VEC;
      // Evaluate the whole vector, then test its best key against this thread's current worst
      // once: after the first few tiles almost every vector is rejected by that single compare,
      // so the per-row cost is the key expression plus one max/min (the kernel must stay under
      // ~20 instructions per row to remain HBM-bound at 4 B/row).
      float key[WDB_VEC];
      float vbest = WDB_KEY_WORST;
      if (have) {
#pragma unroll
        for (int j = 0; j < WDB_VEC; ++j) {
          key[j] = WDB_KEY(R[u], j);
#if WDB_HAS_COND
          if (!WDB_COND(R[u], j)) key[j] = __int_as_float(0x7fc00000);   // NaN: never better than anything
#endif
#if WDB_DESC
          vbest = fmaxf(vbest, key[j]);
#else
          vbest = fminf(vbest, key[j]);
#endif
        }
      }
      // A row below the K-th best key of ANY lane of the warp cannot reach the result (K rows beat it already), so the
      // vector is tested against the best of the lanes' K-th keys (tau), not only against this thread's own list: a
      // thread sees too few rows to build a sharp threshold of its own (6 600 on a 1e9-row shard), and whenever
      // one lane of a warp passes the test the whole warp walks the insertion code -- with private thresholds that
      // was still every fifth vector at the end of such a shard (0.73 ms where 0.61 ms is the roofline).
#if WDB_DESC
      const bool pass = have && vbest >= fmaxf(L.k[WDB_K - 1], tau);
#else
      const bool pass = have && vbest <= fminf(L.k[WDB_K - 1], tau);
#endif
      if (!__any_sync(WDB_FULL_MASK, pass)) continue;                 // warp-uniform
      if (pass) {
#pragma unroll
        for (int j = 0; j < WDB_VEC; ++j) L.offer(key[j], row_base + row + j);
      }
      float b = L.k[WDB_K - 1];                                       // lists changed: refresh the warp's threshold
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(WDB_FULL_MASK, b, o);
#if WDB_DESC
        b = fmaxf(b, ob);
#else
        b = fminf(b, ob);
#endif
      }
#if WDB_DESC
      tau = fmaxf(tau, b);
#else
      tau = fminf(tau, b);
#endif
    }
  }
  if (blockIdx.x == 0) {
    const i64 row = nvec * WDB_VEC + threadIdx.x;
    if (row < n) {
      wdb_rows R;
      wdb_load_row1(C, row, R, 0);
#if WDB_HAS_COND
      if (WDB_COND(R, 0))
#endif
        L.offer(WDB_KEY(R, 0), row_base + row);
    }
  }
  wdb_block_topk(L, cand_k + (i64)blockIdx.x * WDB_K, cand_r + (i64)blockIdx.x * WDB_K);
#if WDB_FUSED_TAIL
  __shared__ u32 s_last;
  __threadfence();                                  // this CTA's candidates are visible before its ticket is
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(done, 1u) == gridDim.x - 1u) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  L.clear();
  const i64 m = (i64)gridDim.x * WDB_K;
  for (i64 i = threadIdx.x; i < m; i += WDB_BLOCK) {
    const i64 r = __ldcg(cand_r + i);               // written by other SMs: read at the L2
    if (r != WDB_ROW_NONE) L.offer(__ldcg(cand_k + i), r);
  }
  wdb_block_topk(L, best_k, best_r);
  __syncthreads();
  wdb_topk_emit_rows(C, row_base, best_k, best_r, offset, out_vals, out_keys, out_count);
#endif
}

extern "C" __global__ void __launch_bounds__(WDB_BLOCK)
wdb_topk_final(const float *__restrict__ cand_k, const i64 *__restrict__ cand_r, const i64 m, float *__restrict__ out_k,
               i64 *__restrict__ out_r) {
  wdb_list L;
  L.clear();
  for (i64 i = threadIdx.x; i < m; i += WDB_BLOCK) {
    const i64 r = cand_r[i];
    if (r != WDB_ROW_NONE) L.offer(cand_k[i], r);
  }
  wdb_block_topk(L, out_k, out_r);
}

// rows[i] are global row ids (row_base-relative columns): out_vals[i - offset] = VAL(row)
extern "C" __global__ void wdb_topk_emit(const wdb_cols C, const i64 row_base, const float *__restrict__ best_k,
                                         const i64 *__restrict__ best_r, const int offset, float *__restrict__ out_vals,
                                         float *__restrict__ out_keys, i64 *__restrict__ out_count) {
  wdb_topk_emit_rows(C, row_base, best_k, best_r, offset, out_vals, out_keys, out_count);
}

// best key of every tile of WDB_TILE_ROWS rows among rows passing the condition (WDB_KEY_WORST if none)
extern "C" __global__ void __launch_bounds__(WDB_BLOCK)
wdb_tile_best(const wdb_cols C, const i64 n, float *__restrict__ tile_best, const i64 ntiles) {
  __shared__ float s_w[WDB_BLOCK / 32];
  for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    float best = WDB_KEY_WORST;
    const i64 row0 = tile * ((i64)WDB_BLOCK * WDB_UNROLL * WDB_VEC);
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u) {
      const i64 row = row0 + ((i64)u * WDB_BLOCK + threadIdx.x) * WDB_VEC;
      if (row + WDB_VEC <= n) {
        wdb_rows R;
        wdb_load_rows(C, row, R);
#pragma unroll
        for (int j = 0; j < WDB_VEC; ++j) {
#if WDB_HAS_COND
          if (!WDB_COND(R, j)) continue;
#endif
          const float k = WDB_KEY(R, j);
#if WDB_DESC
          best = k > best ? k : best;
#else
          best = k < best ? k : best;
#endif
        }
      } else {
        for (int j = 0; j < WDB_VEC; ++j)
          if (row + j < n) {
            wdb_rows R;
            wdb_load_row1(C, row + j, R, 0);
#if WDB_HAS_COND
            if (!WDB_COND(R, 0)) continue;
#endif
            const float k = WDB_KEY(R, 0);
#if WDB_DESC
            best = k > best ? k : best;
#else
            best = k < best ? k : best;
#endif
          }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(WDB_FULL_MASK, best, o);
#if WDB_DESC
      best = ob > best ? ob : best;
#else
      best = ob < best ? ob : best;
#endif
    }
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < WDB_BLOCK / 32; ++w) {
#if WDB_DESC
        best = s_w[w] > best ? s_w[w] : best;
#else
        best = s_w[w] < best ? s_w[w] : best;
#endif
      }
      tile_best[tile] = best;
    }
    __syncthreads();
  }
}
