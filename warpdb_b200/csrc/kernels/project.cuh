// project.cuh -- fused filter + project kernel templates (replaces the `user_kernel` text of
// src/jit.cpp:55-61,81-83: one thread per row, scalar 32-bit LDG/STG, int index).
//
// Roofline: HBM.  Algorithmic bytes per row = sum(sizeof used columns) + 4*selectivity.
//
// Host-supplied macros (besides prelude.cuh's): WDB_BLOCK, WDB_UNROLL, WDB_MODE (0 dense
// untouched | 2 dense zero-fill), WDB_HAS_COND, WDB_TILE/WDB_STAGES (bulk variant),
// and the generated wdb_cols / wdb_rows / WDB_EXPR / WDB_COND.

__device__ __forceinline__ void wdb_emit_vec(float *__restrict__ out, i64 row, const wdb_rows &R) {
  float v[WDB_VEC];
#if WDB_HAS_COND
  u32 m = 0;
#pragma unroll
  for (int j = 0; j < WDB_VEC; ++j) m |= (WDB_COND(R, j) ? 1u : 0u) << j;
#if WDB_MODE == 2
#pragma unroll
  for (int j = 0; j < WDB_VEC; ++j) v[j] = ((m >> j) & 1u) ? WDB_EXPR(R, j) : 0.0f;
  wdb_store_vec(out, row, v);
#else
  // reference semantics: rows failing the condition leave their slot untouched
  if (m == (1u << WDB_VEC) - 1u) {
#pragma unroll
    for (int j = 0; j < WDB_VEC; ++j) v[j] = WDB_EXPR(R, j);
    wdb_store_vec(out, row, v);
  } else if (__popc(m) <= 2) {
    // one or two survivors (the common partial vector at low selectivity): predicated 4-byte stores
#pragma unroll
    for (int j = 0; j < WDB_VEC; ++j)
      if ((m >> j) & 1u) out[row + j] = WDB_EXPR(R, j);
  } else {
#if WDB_ALIGNED
    // partially passing vector: read-modify-write of the whole vector (one 256-bit load + store)
    // instead of up to WDB_VEC predicated 4-byte stores, which cost 2.6 ms per 1e9 rows at 50 %
    wdb_load_out_vec(out, row, v);
#pragma unroll
    for (int j = 0; j < WDB_VEC; ++j)
      if ((m >> j) & 1u) v[j] = WDB_EXPR(R, j);
    wdb_store_vec(out, row, v);
#else
#pragma unroll
    for (int j = 0; j < WDB_VEC; ++j)
      if ((m >> j) & 1u) out[row + j] = WDB_EXPR(R, j);
#endif
  }
#endif
#else
#pragma unroll
  for (int j = 0; j < WDB_VEC; ++j) v[j] = WDB_EXPR(R, j);
  wdb_store_vec(out, row, v);
#endif
}

__device__ __forceinline__ void wdb_emit_row(const wdb_cols &C, float *__restrict__ out, i64 row) {
  wdb_rows R;
  wdb_load_row1(C, row, R, 0);
#if WDB_HAS_COND
  if (WDB_COND(R, 0)) out[row] = WDB_EXPR(R, 0);
#if WDB_MODE == 2
  else out[row] = 0.0f;
#endif
#else
  out[row] = WDB_EXPR(R, 0);
#endif
}

// ---- variant 0/1: vectorised LDG/STG.  Each thread owns WDB_UNROLL vectors of WDB_VEC rows per
// tile, lanes interleaved so every warp instruction touches one contiguous 512 B / 1 KB span; all
// loads of a tile are issued before the first use (memory-level parallelism).  Launched either
// with one CTA per tile (variant 0) or as a persistent grid-stride loop (variant 1).
#if WDB_PRUNE
// Zone-map pruning: zmask[row >> zshift] == 0 means "no row of this zone can pass the condition"
// (decided on the host side of the ABI from per-zone min/max); such vectors are never loaded.
#define WDB_ZONE_ARGS , const unsigned char *__restrict__ zmask, const int zshift
#define WDB_ZONE_LIVE(row) (zmask[(row) >> zshift] != 0)
#else
#define WDB_ZONE_ARGS
#define WDB_ZONE_LIVE(row) true
#endif

__device__ __forceinline__ void wdb_emit_pruned(float *__restrict__ out, i64 row) {
#if WDB_MODE == 2
  float z[WDB_VEC];
#pragma unroll
  for (int j = 0; j < WDB_VEC; ++j) z[j] = 0.0f;
  wdb_store_vec(out, row, z);
#endif
}

extern "C" __global__ void __launch_bounds__(WDB_BLOCK)
wdb_project(const wdb_cols C, float *__restrict__ out, const i64 n WDB_ZONE_ARGS) {
  const i64 nvec = n / WDB_VEC;
  const i64 tile_vecs = (i64)WDB_BLOCK * WDB_UNROLL;
  const i64 ntiles = (nvec + tile_vecs - 1) / tile_vecs;
  for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const i64 v0 = tile * tile_vecs + threadIdx.x;
    wdb_rows R[WDB_UNROLL];
    if ((tile + 1) * tile_vecs <= nvec) {
      bool live[WDB_UNROLL];
#pragma unroll
      for (int u = 0; u < WDB_UNROLL; ++u) live[u] = WDB_ZONE_LIVE((v0 + (i64)u * WDB_BLOCK) * WDB_VEC);
#pragma unroll
      for (int u = 0; u < WDB_UNROLL; ++u)
        if (live[u]) wdb_load_rows(C, (v0 + (i64)u * WDB_BLOCK) * WDB_VEC, R[u]);
#pragma unroll
      for (int u = 0; u < WDB_UNROLL; ++u) {
        if (live[u]) wdb_emit_vec(out, (v0 + (i64)u * WDB_BLOCK) * WDB_VEC, R[u]);
        else wdb_emit_pruned(out, (v0 + (i64)u * WDB_BLOCK) * WDB_VEC);
      }
    } else {
#pragma unroll
      for (int u = 0; u < WDB_UNROLL; ++u) {
        const i64 v = v0 + (i64)u * WDB_BLOCK;
        if (v < nvec) {
          if (WDB_ZONE_LIVE(v * WDB_VEC)) { wdb_load_rows(C, v * WDB_VEC, R[u]); wdb_emit_vec(out, v * WDB_VEC, R[u]); }
          else wdb_emit_pruned(out, v * WDB_VEC);
        }
      }
    }
  }
  // ragged tail (n % WDB_VEC rows)
  if (blockIdx.x == 0) {
    const i64 row = nvec * WDB_VEC + threadIdx.x;
    if (row < n) wdb_emit_row(C, out, row);
  }
}

#if WDB_BULK
// ---- variant 2: bulk-async (TMA 1-D) pipeline.  One thread streams whole column tiles into shared
// memory with cp.async.bulk (UBLKCP) WDB_STAGES-1 tiles ahead; all threads compute from shared
// memory; results leave through a bulk shared->global store.  Only for outputs that are written
// densely (no condition, or zero-fill mode).
extern "C" __global__ void __launch_bounds__(WDB_BLOCK, 1)
wdb_project_bulk(const wdb_cols C, float *__restrict__ out, const i64 n) {
  extern __shared__ __align__(128) unsigned char wdb_smem[];
  u64 *full = reinterpret_cast<u64 *>(wdb_smem);                 // [WDB_STAGES]
  unsigned char *stage0 = wdb_smem + 128;                        // stage s at stage0 + s*WDB_STAGE_BYTES
  const i64 ntiles = n / WDB_TILE;                               // full tiles only; the rest below
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < WDB_STAGES; ++s) wdb_mbar_init(&full[s], 1);
    wdb_fence_barrier_init();
    wdb_fence_proxy_async();
  }
  __syncthreads();
  const i64 first = blockIdx.x, step = gridDim.x;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < WDB_STAGES - 1; ++s) {
      const i64 t = first + (i64)s * step;
      if (t < ntiles) {
        wdb_mbar_expect_tx(&full[s], WDB_IN_BYTES);
        wdb_bulk_load_tile(C, t * WDB_TILE, stage0 + (size_t)s * WDB_STAGE_BYTES, &full[s]);
      }
    }
  }
  int k = 0;
  for (i64 t = first; t < ntiles; t += step, ++k) {
    const int s = k % WDB_STAGES;
    unsigned char *sb = stage0 + (size_t)s * WDB_STAGE_BYTES;
    // keep the pipeline WDB_STAGES-1 deep: the buffer of tile k-1 was released by the barrier
    // at the end of the previous iteration
    if (threadIdx.x == 0) {
      const i64 tn = t + (i64)(WDB_STAGES - 1) * step;
      if (tn < ntiles) {
        const int sn = (k + WDB_STAGES - 1) % WDB_STAGES;
        wdb_mbar_expect_tx(&full[sn], WDB_IN_BYTES);
        wdb_bulk_load_tile(C, tn * WDB_TILE, stage0 + (size_t)sn * WDB_STAGE_BYTES, &full[sn]);
      }
    }
    wdb_mbar_wait(&full[s], (u32)((k / WDB_STAGES) & 1));
    float *so = reinterpret_cast<float *>(sb + WDB_IN_BYTES);
#pragma unroll
    for (int i = 0; i < WDB_TILE / 4 / WDB_BLOCK; ++i) {
      const int r = (threadIdx.x + i * WDB_BLOCK) * 4;
      wdb_rows4s R;
      wdb_lds_rows(sb, r, R);
      float4 v;
#if WDB_HAS_COND
      v.x = WDB_COND4(R, 0) ? WDB_EXPR4(R, 0) : 0.0f;
      v.y = WDB_COND4(R, 1) ? WDB_EXPR4(R, 1) : 0.0f;
      v.z = WDB_COND4(R, 2) ? WDB_EXPR4(R, 2) : 0.0f;
      v.w = WDB_COND4(R, 3) ? WDB_EXPR4(R, 3) : 0.0f;
#else
      v.x = WDB_EXPR4(R, 0); v.y = WDB_EXPR4(R, 1); v.z = WDB_EXPR4(R, 2); v.w = WDB_EXPR4(R, 3);
#endif
      *reinterpret_cast<float4 *>(so + r) = v;
    }
    wdb_fence_proxy_async();
    // the output buffer of this stage is rewritten WDB_STAGES iterations from now; the input
    // buffer of stage (k-1)%S... is refilled next iteration: both need every thread past this point
    if (threadIdx.x == 0) wdb_bulk_wait_read<WDB_STAGES - 2>();
    __syncthreads();
    if (threadIdx.x == 0) {
      wdb_bulk_s2g(out + t * WDB_TILE, so, WDB_TILE * 4);
      wdb_bulk_commit();
    }
  }
  if (threadIdx.x == 0) wdb_bulk_wait<0>();
  // rows past the last full tile: plain path, spread over the grid
  for (i64 row = ntiles * WDB_TILE + (i64)blockIdx.x * WDB_BLOCK + threadIdx.x; row < n; row += (i64)gridDim.x * WDB_BLOCK)
    wdb_emit_row(C, out, row);
}
#endif
