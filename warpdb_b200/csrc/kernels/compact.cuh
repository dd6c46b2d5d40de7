// compact.cuh -- fused filter + project + STABLE stream compaction in one pass.
//
// The reference never compacts on the GPU; row-order compaction exists only in the
// host half of query_sql (src/warpdb.cpp:336-344,457-459).  Here every CTA takes tiles in ticket
// order, evaluates cond/expr on WDB_VEC-wide vector loads, ranks survivors with warp ballots +
// popc (no shuffles), stages them in shared memory in row order and obtains its global offset with
// a decoupled look-back over 64-bit tile status words, so the packed output is written once, fully
// coalesced and in row order.
//
// Two kernels: wdb_compact (register-staged vector loads, tiles handed out by an atomic ticket) and
// wdb_compact_bulk (TMA bulk copies into a shared-memory ring, static tile assignment).
//
// Roofline: HBM.  Algorithmic bytes per row = sum(sizeof used columns) + 4*WDB_NOUT*selectivity.
//
// Host-supplied macros: WDB_BLOCK, WDB_UNROLL (slabs per warp per tile), WDB_VEC, WDB_NOUT (1|2),
// generated WDB_COND, WDB_EXPR and (NOUT==2) WDB_EXPR2.
#define WDB_NWARPS (WDB_BLOCK / 32)
#define WDB_SLAB_ROWS (32 * WDB_VEC)
#define WDB_WARP_ROWS (WDB_SLAB_ROWS * WDB_UNROLL)
#define WDB_TILE_ROWS (WDB_WARP_ROWS * WDB_NWARPS)

// Rows of a vector are lane-major (lane l holds rows 8l .. 8l+7), so the stable rank of a lane's
// first survivor is the number of survivors in lower lanes: an exclusive warp scan of the per-lane
// counts (6 shuffles per vector; the ballot-per-row formulation costs 8 votes + 16 popcounts).
__device__ __forceinline__ void wdb_warp_rank(const u32 c, const u32 lane, u32 &pre, u32 &tot) {
  u32 x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 y = __shfl_up_sync(WDB_FULL_MASK, x, o);
    if (lane >= (u32)o) x += y;
  }
  pre = x - c;
  tot = __shfl_sync(WDB_FULL_MASK, x, 31);
}

// WDB_THRESH: extra keep-test against the run-time threshold wdb_tau on the second expression (the
// ORDER BY key): 1 keeps key >= tau, 2 keeps key <= tau (negated compares, so NaN keys stay)
#if WDB_THRESH == 1
#define WDB_KEEP(R, j) (WDB_COND(R, j) && !(WDB_EXPR2(R, j) < wdb_tau))
#elif WDB_THRESH == 2
#define WDB_KEEP(R, j) (WDB_COND(R, j) && !(WDB_EXPR2(R, j) > wdb_tau))
#else
#define WDB_KEEP(R, j) WDB_COND(R, j)
#endif

// WDB_AUTO_SEL: the kernel belongs to one of two pipelines launched back to back; a sampled selectivity left
// on the device by wdb_sample_count (survivors, rows) decides -- without a host round trip -- which of them
// does the work: the staged two-pass kernels for selective filters, the L2-parked slabs otherwise.
#if WDB_AUTO_SEL
#define WDB_SEL_PARAM , const u64 *__restrict__ wdb_sel
#define WDB_SEL_STAGED (wdb_sel[0] * 1000ull <= wdb_sel[1] * (u64)WDB_STAGE_PERMILLE)
#else
#define WDB_SEL_PARAM
#endif

#define WDB_ST_AGG 1ull
#define WDB_ST_PREFIX 2ull
#define WDB_ST_MASK ((1ull << 62) - 1ull)

__device__ __forceinline__ u64 wdb_ld_status(const u64 *p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wdb_st_status(u64 *p, u64 v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 wdb_warp_sum64(u64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(WDB_FULL_MASK, v, o);
  return v;
}


// Decoupled look-back over a window of 32*WDB_LB predecessor tiles per step: lane l inspects the
// WDB_LB status words of tiles look-(l*WDB_LB+q).  Returns the exclusive prefix of `tile` (> 0).
#ifndef WDB_LB
#define WDB_LB 1
#endif
__device__ __forceinline__ i64 wdb_lookback(const u64 *__restrict__ status, const i64 tile, const u32 lane) {
  i64 excl = 0;
  i64 look = tile - 1;
  while (true) {
    u64 part = 0;          // sum of this lane's statuses up to and including its first PREFIX
    bool has_prefix = false;
#pragma unroll
    for (int q = 0; q < WDB_LB; ++q) {
      const i64 idx = look - ((i64)lane * WDB_LB + q);
      u64 st = (WDB_ST_PREFIX << 62);              // before the first tile: prefix 0
      if (idx >= 0) {
        do { st = wdb_ld_status(&status[idx]); } while ((st >> 62) == 0ull);
      }
      if (!has_prefix) {
        part += st & WDB_ST_MASK;
        has_prefix = (st >> 62) == WDB_ST_PREFIX;
      }
    }
    const u32 pm = __ballot_sync(WDB_FULL_MASK, has_prefix);
    if (pm) {
      const u32 first = (u32)__ffs((int)pm) - 1u;
      return excl + (i64)wdb_warp_sum64(lane <= first ? part : 0ull);
    }
    excl += (i64)wdb_warp_sum64(part);
    look -= 32 * WDB_LB;
  }
}

#if !WDB_BULK && !WDB_TWOPASS && !WDB_L2PASS
// ---- variant 0: register-staged vector loads, tiles handed out in ticket order
extern "C" __global__ void __launch_bounds__(WDB_BLOCK)
wdb_compact(const wdb_cols C, float *__restrict__ out, float *__restrict__ out2, const i64 n,
            u64 *__restrict__ status, u32 *__restrict__ ticket, i64 *__restrict__ out_count, const i64 ntiles,
            const float wdb_tau, const i64 out_cap) {
  __shared__ float s_stage[WDB_TILE_ROWS];
#if WDB_NOUT == 2
  __shared__ float s_stage2[WDB_TILE_ROWS];
#endif
  __shared__ u32 s_wcount[WDB_NWARPS];
  __shared__ i64 s_base;
  __shared__ u32 s_tile;
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const u32 lt = wdb_lanemask_lt();

  while (true) {
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();                                            // (1)
    const i64 tile = (i64)s_tile;
    if (tile >= ntiles) break;
    const i64 wrow0 = tile * WDB_TILE_ROWS + (i64)warp * WDB_WARP_ROWS + (i64)lane * WDB_VEC;

    u32 flags[WDB_UNROLL];
    float vals[WDB_UNROLL][WDB_VEC];
#if WDB_NOUT == 2
    float vals2[WDB_UNROLL][WDB_VEC];
#endif
    if ((tile + 1) * WDB_TILE_ROWS <= n) {
      wdb_rows R[WDB_UNROLL];
#pragma unroll
      for (int u = 0; u < WDB_UNROLL; ++u) wdb_load_rows(C, wrow0 + (i64)u * WDB_SLAB_ROWS, R[u]);
#pragma unroll
      for (int u = 0; u < WDB_UNROLL; ++u) {
        u32 m = 0;
#pragma unroll
        for (int j = 0; j < WDB_VEC; ++j) {
          m |= (WDB_KEEP(R[u], j) ? 1u : 0u) << j;
          vals[u][j] = WDB_EXPR(R[u], j);
#if WDB_NOUT == 2
          vals2[u][j] = WDB_EXPR2(R[u], j);
#endif
        }
        flags[u] = m;
      }
    } else {
#pragma unroll
      for (int u = 0; u < WDB_UNROLL; ++u) {
        u32 m = 0;
#pragma unroll
        for (int j = 0; j < WDB_VEC; ++j) {
          const i64 row = wrow0 + (i64)u * WDB_SLAB_ROWS + j;
          vals[u][j] = 0.0f;
#if WDB_NOUT == 2
          vals2[u][j] = 0.0f;
#endif
          if (row < n) {
            wdb_rows R;
            wdb_load_row1(C, row, R, 0);
            if (WDB_KEEP(R, 0)) {
              m |= 1u << j;
              vals[u][j] = WDB_EXPR(R, 0);
#if WDB_NOUT == 2
              vals2[u][j] = WDB_EXPR2(R, 0);
#endif
            }
          }
        }
        flags[u] = m;
      }
    }

    // rank of this lane's first survivor inside the warp's region: ballots give, per element slot
    // j, the set of lanes that keep it; rows are ordered (slab, lane, j)
    u32 rank[WDB_UNROLL], wtotal = 0;
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u) {
      u32 pre = 0, tot = 0;
#pragma unroll
      for (int j = 0; j < WDB_VEC; ++j) {
        const u32 b = __ballot_sync(WDB_FULL_MASK, (flags[u] >> j) & 1u);
        pre += __popc(b & lt);
        tot += __popc(b);
      }
      rank[u] = wtotal + pre;
      wtotal += tot;
    }
    if (lane == 0) s_wcount[warp] = wtotal;
    __syncthreads();                                            // (2)
    u32 woff = 0, ttotal = 0;
#pragma unroll
    for (int w = 0; w < WDB_NWARPS; ++w) {
      const u32 c = s_wcount[w];
      woff += (w < (int)warp) ? c : 0u;
      ttotal += c;
    }

    if (warp == 0) {  // decoupled look-back
      i64 excl = 0;
      if (tile == 0) {
        if (lane == 0) wdb_st_status(&status[0], (WDB_ST_PREFIX << 62) | (u64)ttotal);
      } else {
        if (lane == 0) wdb_st_status(&status[tile], (WDB_ST_AGG << 62) | (u64)ttotal);
        excl = wdb_lookback(status, tile, lane);
        if (lane == 0) wdb_st_status(&status[tile], (WDB_ST_PREFIX << 62) | (u64)(excl + (i64)ttotal));
      }
      if (lane == 0) {
        s_base = excl;
        if (tile == ntiles - 1) *out_count = excl + (i64)ttotal;
      }
    }

    // stage survivors in row order
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u) {
      u32 pos = woff + rank[u];
#pragma unroll
      for (int j = 0; j < WDB_VEC; ++j)
        if ((flags[u] >> j) & 1u) {
          s_stage[pos] = vals[u][j];
#if WDB_NOUT == 2
          s_stage2[pos] = vals2[u][j];
#endif
          ++pos;
        }
    }
    __syncthreads();                                            // (3)
    const i64 g0 = s_base;
    // copy out with warps writing 128-byte aligned spans of the destination
    const int mis = (int)(g0 & 31);
    for (int i = (int)threadIdx.x - mis; i < (int)ttotal; i += WDB_BLOCK)
      if (i >= 0 && g0 + i < out_cap) {   // the count stays exact when the output is too small
        out[g0 + i] = s_stage[i];
#if WDB_NOUT == 2
        out2[g0 + i] = s_stage2[i];
#endif
      }
    // the barriers (1) and (2) of the next iteration order these reads before the next staging
  }
}

#endif  // variant 0

#if WDB_BULK
// ---- variant 1: the same algorithm with the column tiles streamed into a double-buffered shared
// memory ring by bulk asynchronous copies (TMA 1-D, cp.async.bulk -> UBLKCP).  The tile after the
// current one is in flight for the whole time this CTA ranks, resolves its offset and copies out,
// so the serial latency chain of a tile (load -> rank -> look-back -> copy-out) no longer leaves
// HBM idle, and the loaded data no longer lives in registers across barriers.  Tiles are assigned
// statically (tile = blockIdx.x + k*gridDim.x) so the next tile is known without a ticket; the grid
// never exceeds the number of co-resident CTAs, hence every predecessor of a tile is running.
#ifndef WDB_MIN_CTAS
#define WDB_MIN_CTAS 1
#endif
extern "C" __global__ void __launch_bounds__(WDB_BLOCK, WDB_MIN_CTAS)
wdb_compact_bulk(const wdb_cols C, float *__restrict__ out, float *__restrict__ out2, const i64 n,
                 u64 *__restrict__ status, i64 *__restrict__ out_count, const i64 ntiles,
                 const float wdb_tau, const i64 out_cap) {
  extern __shared__ __align__(128) unsigned char wdb_smem[];
  u64 *full = reinterpret_cast<u64 *>(wdb_smem);                         // 2 mbarriers
  unsigned char *in0 = wdb_smem + 128;                                   // 2 x WDB_IN_BYTES
  float *s_stage = reinterpret_cast<float *>(in0 + 2 * (size_t)WDB_IN_BYTES);
#if WDB_NOUT == 2
  float *s_stage2 = s_stage + WDB_TILE_ROWS;
#endif
  __shared__ u32 s_wcount[WDB_NWARPS];
  __shared__ i64 s_base;
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const u32 lt = wdb_lanemask_lt();
  const i64 nfull = n / WDB_TILE_ROWS;                                   // tiles that can be bulk-loaded

  if (threadIdx.x == 0) {
    wdb_mbar_init(&full[0], 1);
    wdb_mbar_init(&full[1], 1);
    wdb_fence_barrier_init();
    wdb_fence_proxy_async();
  }
  __syncthreads();
  const i64 first = blockIdx.x, step = gridDim.x;
  if (threadIdx.x == 0 && first < nfull) {
    wdb_mbar_expect_tx(&full[0], WDB_IN_BYTES);
    wdb_bulk_load_tile(C, first * WDB_TILE_ROWS, in0, &full[0]);
  }

  u32 it = 0;
  for (i64 tile = first; tile < ntiles; tile += step, ++it) {
    const u32 buf = it & 1u;
    // the other buffer was last read before barrier (A) of the previous iteration: refill it now
    if (threadIdx.x == 0 && tile + step < nfull) {
      wdb_mbar_expect_tx(&full[buf ^ 1u], WDB_IN_BYTES);
      wdb_bulk_load_tile(C, (tile + step) * WDB_TILE_ROWS, in0 + (size_t)(buf ^ 1u) * WDB_IN_BYTES, &full[buf ^ 1u]);
    }
    const int lrow0 = (int)(warp * WDB_WARP_ROWS + lane * WDB_VEC);       // row inside the tile

    u32 flags[WDB_UNROLL];
    float vals[WDB_UNROLL][WDB_VEC];
#if WDB_NOUT == 2
    float vals2[WDB_UNROLL][WDB_VEC];
#endif
    if (tile < nfull) {
      wdb_mbar_wait(&full[buf], (it >> 1) & 1u);
      const unsigned char *sb = in0 + (size_t)buf * WDB_IN_BYTES;
#pragma unroll
      for (int u = 0; u < WDB_UNROLL; ++u) {
        u32 m = 0;
#pragma unroll
        for (int h = 0; h < WDB_VEC / 4; ++h) {
          wdb_rows4s R4;
          wdb_lds_rows(sb, lrow0 + u * WDB_SLAB_ROWS + 4 * h, R4);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            m |= (WDB_KEEP(R4, j) ? 1u : 0u) << (4 * h + j);
            vals[u][4 * h + j] = WDB_EXPR(R4, j);
#if WDB_NOUT == 2
            vals2[u][4 * h + j] = WDB_EXPR2(R4, j);
#endif
          }
        }
        flags[u] = m;
      }
    } else {  // ragged last tile: guarded direct loads
      const i64 wrow0 = tile * WDB_TILE_ROWS + lrow0;
#pragma unroll
      for (int u = 0; u < WDB_UNROLL; ++u) {
        u32 m = 0;
#pragma unroll
        for (int j = 0; j < WDB_VEC; ++j) {
          const i64 row = wrow0 + (i64)u * WDB_SLAB_ROWS + j;
          vals[u][j] = 0.0f;
#if WDB_NOUT == 2
          vals2[u][j] = 0.0f;
#endif
          if (row < n) {
            wdb_rows T1;
            wdb_load_row1(C, row, T1, 0);
            if (WDB_KEEP(T1, 0)) {
              m |= 1u << j;
              vals[u][j] = WDB_EXPR(T1, 0);
#if WDB_NOUT == 2
              vals2[u][j] = WDB_EXPR2(T1, 0);
#endif
            }
          }
        }
        flags[u] = m;
      }
    }

    u32 rank[WDB_UNROLL], wtotal = 0;
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u) {
      u32 pre = 0, tot = 0;
#pragma unroll
      for (int j = 0; j < WDB_VEC; ++j) {
        const u32 b = __ballot_sync(WDB_FULL_MASK, (flags[u] >> j) & 1u);
        pre += __popc(b & lt);
        tot += __popc(b);
      }
      rank[u] = wtotal + pre;
      wtotal += tot;
    }
    if (lane == 0) s_wcount[warp] = wtotal;
    __syncthreads();                                            // (A)
    u32 woff = 0, ttotal = 0;
#pragma unroll
    for (int w = 0; w < WDB_NWARPS; ++w) {
      const u32 c = s_wcount[w];
      woff += (w < (int)warp) ? c : 0u;
      ttotal += c;
    }
    if (threadIdx.x == 0)   // publish the aggregate before anything else
      wdb_st_status(&status[tile], ((tile == 0 ? WDB_ST_PREFIX : WDB_ST_AGG) << 62) | (u64)ttotal);
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u) {
      u32 pos = woff + rank[u];
#pragma unroll
      for (int j = 0; j < WDB_VEC; ++j)
        if ((flags[u] >> j) & 1u) {
          s_stage[pos] = vals[u][j];
#if WDB_NOUT == 2
          s_stage2[pos] = vals2[u][j];
#endif
          ++pos;
        }
    }
    if (warp == 0) {
      i64 excl = 0;
      if (tile > 0) {
        excl = wdb_lookback(status, tile, lane);
        if (lane == 0) wdb_st_status(&status[tile], (WDB_ST_PREFIX << 62) | (u64)(excl + (i64)ttotal));
      }
      if (lane == 0) {
        s_base = excl;
        if (tile == ntiles - 1) *out_count = excl + (i64)ttotal;
      }
    }
    __syncthreads();                                            // (B)
    const i64 g0 = s_base;
    const int mis = (int)(g0 & 31);
    for (int i = (int)threadIdx.x - mis; i < (int)ttotal; i += WDB_BLOCK)
      if (i >= 0 && g0 + i < out_cap) {
        out[g0 + i] = s_stage[i];
#if WDB_NOUT == 2
        out2[g0 + i] = s_stage2[i];
#endif
      }
    // barrier (A) of the next iteration orders these reads before the next staging writes
  }
}
#endif

#if WDB_TWOPASS
// ---- variant 2: two streaming passes without any inter-CTA dependency.
//   wdb_count    every warp counts the survivors of its chunk of WDB_WARP_ROWS rows (condition only)
//   (host)       exclusive scan of the chunk counts
//   wdb_scatter  every warp re-reads its chunk, ranks survivors with ballots, stages them in a
//                warp-private slice of shared memory and writes them at its scanned offset
// No tickets, no look-back, no block barriers: both kernels are plain streaming kernels.  The price
// is reading the columns the condition needs a second time: (cond bytes) + (all used bytes) +
// 4*NOUT*selectivity per row instead of (all used bytes) + 4*NOUT*selectivity.
#if WDB_PRUNE
// zone-map pruning: a chunk whose zone cannot contain a passing row is neither loaded nor counted
#define WDB_ZONE_ARGS , const unsigned char *__restrict__ zmask, const int zshift
#define WDB_CHUNK_DEAD(chunk) (zmask[((chunk) * WDB_WARP_ROWS) >> zshift] == 0)
#else
#define WDB_ZONE_ARGS
#define WDB_CHUNK_DEAD(chunk) false
#endif
extern "C" __global__ void __launch_bounds__(WDB_BLOCK)
wdb_count(const wdb_cols C, const i64 n, u32 *__restrict__ counts, const i64 nchunks, const float wdb_tau WDB_ZONE_ARGS) {
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const i64 chunk = (i64)blockIdx.x * WDB_NWARPS + warp;
  if (chunk >= nchunks) return;
  const i64 row0 = chunk * WDB_WARP_ROWS + (i64)lane * WDB_VEC;
  u32 cnt = 0;
  if (WDB_CHUNK_DEAD(chunk)) {
    if (lane == 0) counts[chunk] = 0u;
    return;
  }
  if ((chunk + 1) * WDB_WARP_ROWS <= n) {
    wdb_rows R[WDB_UNROLL];
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u) wdb_load_rows(C, row0 + (i64)u * WDB_SLAB_ROWS, R[u]);
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u)
#pragma unroll
      for (int j = 0; j < WDB_VEC; ++j) cnt += WDB_KEEP(R[u], j) ? 1u : 0u;
  } else {
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u)
      for (int j = 0; j < WDB_VEC; ++j) {
        const i64 row = row0 + (i64)u * WDB_SLAB_ROWS + j;
        if (row < n) {
          wdb_rows T1;
          wdb_load_row1(C, row, T1, 0);
          cnt += WDB_KEEP(T1, 0) ? 1u : 0u;
        }
      }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(WDB_FULL_MASK, cnt, o);
  if (lane == 0) counts[chunk] = cnt;
}

// flags (one bit per row of the lane's vectors) and values of one warp chunk
__device__ __forceinline__ void wdb_twopass_eval(const wdb_cols &C, const i64 n, const i64 chunk, const u32 lane, const float wdb_tau,
                                                 u32 (&flags)[WDB_UNROLL], float (&vals)[WDB_UNROLL][WDB_VEC]
#if WDB_NOUT == 2
                                                 , float (&vals2)[WDB_UNROLL][WDB_VEC]
#endif
) {
  const i64 row0 = chunk * WDB_WARP_ROWS + (i64)lane * WDB_VEC;
  if ((chunk + 1) * WDB_WARP_ROWS <= n) {
    wdb_rows R[WDB_UNROLL];
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u) wdb_load_rows(C, row0 + (i64)u * WDB_SLAB_ROWS, R[u]);
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u) {
      u32 m = 0;
#pragma unroll
      for (int j = 0; j < WDB_VEC; ++j) {
        m |= (WDB_KEEP(R[u], j) ? 1u : 0u) << j;
        vals[u][j] = WDB_EXPR(R[u], j);
#if WDB_NOUT == 2
        vals2[u][j] = WDB_EXPR2(R[u], j);
#endif
      }
      flags[u] = m;
    }
  } else {
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u) {
      u32 m = 0;
#pragma unroll
      for (int j = 0; j < WDB_VEC; ++j) {
        const i64 row = row0 + (i64)u * WDB_SLAB_ROWS + j;
        vals[u][j] = 0.0f;
#if WDB_NOUT == 2
        vals2[u][j] = 0.0f;
#endif
        if (row < n) {
          wdb_rows T1;
          wdb_load_row1(C, row, T1, 0);
          if (WDB_KEEP(T1, 0)) {
            m |= 1u << j;
            vals[u][j] = WDB_EXPR(T1, 0);
#if WDB_NOUT == 2
            vals2[u][j] = WDB_EXPR2(T1, 0);
#endif
          }
        }
      }
      flags[u] = m;
    }
  }
}

// rank the chunk's survivors, stage them in the warp's slice of shared memory in row order and write them at g0
__device__ __forceinline__ void wdb_twopass_write(const u32 (&flags)[WDB_UNROLL], const float (&vals)[WDB_UNROLL][WDB_VEC],
#if WDB_NOUT == 2
                                                  const float (&vals2)[WDB_UNROLL][WDB_VEC], float *st2, float *__restrict__ out2,
#endif
                                                  float *st, const u32 lane, const i64 g0, float *__restrict__ out, const i64 out_cap) {
  u32 total = 0;
#pragma unroll
  for (int u = 0; u < WDB_UNROLL; ++u) {
    u32 pre, tot;
    wdb_warp_rank(__popc(flags[u]), lane, pre, tot);
    u32 pos = total + pre;
#pragma unroll
    for (int j = 0; j < WDB_VEC; ++j)
      if ((flags[u] >> j) & 1u) {
        st[pos] = vals[u][j];
#if WDB_NOUT == 2
        st2[pos] = vals2[u][j];
#endif
        ++pos;
      }
    total += tot;
  }
  __syncwarp();
  const int mis = (int)(g0 & 31);
  for (int i = (int)lane - mis; i < (int)total; i += 32)
    if (i >= 0 && g0 + i < out_cap) {
      out[g0 + i] = st[i];
#if WDB_NOUT == 2
      out2[g0 + i] = st2[i];
#endif
    }
  __syncwarp();
}

extern "C" __global__ void __launch_bounds__(WDB_BLOCK)
wdb_scatter(const wdb_cols C, float *__restrict__ out, float *__restrict__ out2, const i64 n,
            const i64 *__restrict__ offsets, const i64 nchunks, const float wdb_tau, const i64 out_cap WDB_ZONE_ARGS) {
  __shared__ float s_stage[WDB_NWARPS][WDB_WARP_ROWS];
#if WDB_NOUT == 2
  __shared__ float s_stage2[WDB_NWARPS][WDB_WARP_ROWS];
#endif
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const i64 chunk = (i64)blockIdx.x * WDB_NWARPS + warp;
  if (chunk >= nchunks) return;
  if (WDB_CHUNK_DEAD(chunk)) return;
  u32 flags[WDB_UNROLL];
  float vals[WDB_UNROLL][WDB_VEC];
#if WDB_NOUT == 2
  float vals2[WDB_UNROLL][WDB_VEC];
  wdb_twopass_eval(C, n, chunk, lane, wdb_tau, flags, vals, vals2);
  wdb_twopass_write(flags, vals, vals2, s_stage2[warp], out2, s_stage[warp], lane, offsets[chunk], out, out_cap);
#else
  wdb_twopass_eval(C, n, chunk, lane, wdb_tau, flags, vals);
  wdb_twopass_write(flags, vals, s_stage[warp], lane, offsets[chunk], out, out_cap);
#endif
}

#if WDB_STAGE_CAP > 0 && WDB_NOUT == 1
// ---- variant 5: selective filters (the optimizer's sampled selectivity is below ~8 %).  Pass 1 is the count
// pass of variant 2 that ALSO evaluates the expression and parks each chunk's first WDB_STAGE_CAP survivors, in
// row order, in a per-chunk slot of a scratch array -- no global offset is needed for that, so the pass is
// a plain streaming kernel without tickets, look-back or block barriers.  After the scan of the counts,
// pass 2 only moves the parked survivors to their final place (a warp serves 32 chunks; their counts and
// offsets arrive in one coalesced load); a chunk that overflowed its slot is recomputed from the input
// like variant 2 does.  Traffic: 4 + 12 s bytes per row instead of 8 + 4 s (variant 2) -- 4.12 at s = 1 % against
// the 4.04 that are algorithmically necessary -- and none of the per-slab latency chain of variant 3.
// optimizer statistic: survivors among every `stride`-th whole chunk (the host samples ~256 chunks, 1 MB of a
// float column, to choose between this variant and the L2-parked slabs of variant 3)
extern "C" __global__ void __launch_bounds__(WDB_BLOCK)
wdb_sample_count(const wdb_cols C, const i64 n, u64 *__restrict__ totals, const i64 nchunks, const i64 stride, const float wdb_tau) {
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const i64 chunk = ((i64)blockIdx.x * WDB_NWARPS + warp) * stride;
  if (chunk >= nchunks || (chunk + 1) * WDB_WARP_ROWS > n) return;
  const i64 row0 = chunk * WDB_WARP_ROWS + (i64)lane * WDB_VEC;
  wdb_rows R[WDB_UNROLL];
#pragma unroll
  for (int u = 0; u < WDB_UNROLL; ++u) wdb_load_rows(C, row0 + (i64)u * WDB_SLAB_ROWS, R[u]);
  u32 cnt = 0;
#pragma unroll
  for (int u = 0; u < WDB_UNROLL; ++u)
#pragma unroll
    for (int j = 0; j < WDB_VEC; ++j) cnt += WDB_KEEP(R[u], j) ? 1u : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(WDB_FULL_MASK, cnt, o);
  if (lane == 0) { atomicAdd(&totals[0], (u64)cnt); atomicAdd(&totals[1], (u64)WDB_WARP_ROWS); }
}

// A warp serves WDB_STAGE_M consecutive chunks: the grid is that much smaller, which matters when this pipeline
// is the one that is NOT needed -- a quarter of a million CTAs that only return still cost 0.15 ms of CTA launches.
#ifndef WDB_STAGE_M
#define WDB_STAGE_M 8
#endif
extern "C" __global__ void __launch_bounds__(WDB_BLOCK)
wdb_count_stage(const wdb_cols C, const i64 n, u32 *__restrict__ counts, u64 *__restrict__ group_totals, float *__restrict__ scratch,
                const i64 nchunks, const float wdb_tau WDB_SEL_PARAM WDB_ZONE_ARGS) {
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const i64 chunk0 = ((i64)blockIdx.x * WDB_NWARPS + warp) * WDB_STAGE_M;
#if WDB_AUTO_SEL
  if (!WDB_SEL_STAGED) {                             // the other pipeline does the work: leave empty counts behind
    if (lane < WDB_STAGE_M && chunk0 + lane < nchunks) counts[chunk0 + lane] = 0u;
    return;
  }
#endif
#pragma unroll 1
  for (int m = 0; m < WDB_STAGE_M; ++m) {
    const i64 chunk = chunk0 + m;
    if (chunk >= nchunks) return;
    if (WDB_CHUNK_DEAD(chunk)) {
      if (lane == 0) counts[chunk] = 0u;
      continue;
    }
    u32 flags[WDB_UNROLL];
    float vals[WDB_UNROLL][WDB_VEC];
    wdb_twopass_eval(C, n, chunk, lane, wdb_tau, flags, vals);
    float *slot = scratch + chunk * WDB_STAGE_CAP;
    u32 total = 0;
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u) {
      u32 pre, tot;
      wdb_warp_rank(__popc(flags[u]), lane, pre, tot);
      u32 pos = total + pre;
      if (flags[u]) {                                // most lanes hold no survivor at low selectivity
#pragma unroll
        for (int j = 0; j < WDB_VEC; ++j)
          if ((flags[u] >> j) & 1u) {
            if (pos < (u32)WDB_STAGE_CAP) slot[pos] = vals[u][j];
            ++pos;
          }
      }
      total += tot;
    }
    if (lane == 0) {
      counts[chunk] = total;
      if (total) atomicAdd(&group_totals[chunk >> 5], (u64)total);     // 32 chunks = the unit one gather warp serves
    }
  }
}

// Pass 2.  A warp serves a group of 32 consecutive chunks: their counts arrive in one coalesced load, the group's
// global offset is the scanned group total, and the group's survivors form ONE contiguous output range.  Output
// element e of the group is located with a 5-step binary search over the lanes' exclusive prefixes (shuffles),
// so every load is independent of the others and the stores are coalesced -- a first version that walked the 32
// chunks one after the other spent 0.13 ms per 1e9 rows waiting for one dependent load per chunk.
extern "C" __global__ void __launch_bounds__(WDB_BLOCK)
wdb_gather_stage(const wdb_cols C, float *__restrict__ out, const i64 n, const u32 *__restrict__ counts, const i64 *__restrict__ group_offsets,
                 const float *__restrict__ scratch, const i64 nchunks, const float wdb_tau, const i64 out_cap WDB_SEL_PARAM WDB_ZONE_ARGS) {
  __shared__ float s_stage[WDB_NWARPS][WDB_WARP_ROWS];
#if WDB_AUTO_SEL
  if (!WDB_SEL_STAGED) return;
#endif
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const i64 group = (i64)blockIdx.x * WDB_NWARPS + warp;
  const i64 chunk0 = group * 32;
  if (chunk0 >= nchunks) return;
  const u32 my_cnt = chunk0 + lane < nchunks ? counts[chunk0 + lane] : 0u;
  u32 pre, total;
  wdb_warp_rank(my_cnt, lane, pre, total);
  if (total == 0u) return;
  const i64 g0 = group_offsets[group];
  if (__any_sync(WDB_FULL_MASK, my_cnt > (u32)WDB_STAGE_CAP)) {
    // a slot of this group overflowed: serve its chunks one by one, recomputing the overflowed ones from the input
#pragma unroll 1
    for (int k = 0; k < 32; ++k) {
      const u32 cnt = __shfl_sync(WDB_FULL_MASK, my_cnt, k);
      if (cnt == 0u) continue;                                           // warp-uniform
      const i64 gk = g0 + (i64)__shfl_sync(WDB_FULL_MASK, pre, k);
      const i64 chunk = chunk0 + k;
      if (cnt <= (u32)WDB_STAGE_CAP) {
        const float *slot = scratch + chunk * WDB_STAGE_CAP;
        for (u32 i = lane; i < cnt; i += 32)
          if (gk + i < out_cap) out[gk + i] = slot[i];
      } else {
        u32 flags[WDB_UNROLL];
        float vals[WDB_UNROLL][WDB_VEC];
        wdb_twopass_eval(C, n, chunk, lane, wdb_tau, flags, vals);
        wdb_twopass_write(flags, vals, s_stage[warp], lane, gk, out, out_cap);
      }
    }
    return;
  }
  const float *slots = scratch + chunk0 * WDB_STAGE_CAP;
  const int mis = (int)(g0 & 31);                                        // every store instruction covers one aligned 128-byte line
#pragma unroll 2
  for (int base = -mis; base < (int)total; base += 32) {                // warp-uniform trip count: the shuffles below need every lane
    const int e = base + (int)lane;
    const bool active = e >= 0 && e < (int)total;
    u32 k = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const u32 cand = k + step;                                         // <= 31
      const u32 pc = __shfl_sync(WDB_FULL_MASK, pre, cand);
      if (active && pc <= (u32)e) k = cand;                              // largest lane whose exclusive prefix is <= e
    }
    const u32 pk = __shfl_sync(WDB_FULL_MASK, pre, k);
    if (active && g0 + e < out_cap) out[g0 + e] = slots[(size_t)k * WDB_STAGE_CAP + ((u32)e - pk)];
  }
}
#endif
#endif

#if WDB_L2PASS
// ---- variant 3: single HBM pass with the tile parked in L2.  A CTA owns a slab of
// WDB_NWARPS * WDB_SLAB_M chunks (e.g. 32 K rows = 128 KB of a float column).  Phase 1 streams the
// slab from HBM and only counts survivors; the CTA then resolves its global offset with ONE
// decoupled look-back per slab; phase 2 re-reads the slab -- which is still in the 126 MB L2 as long
// as the slabs in flight (CTAs x slab bytes) fit -- ranks the survivors and writes them.  Compared
// with the register/shared-memory single-pass kernels the look-back latency is paid once per 128 KB
// instead of once per 32 KB and is covered by the other resident CTAs; compared with the two-pass
// variant the second read comes from L2, not HBM.
#ifndef WDB_SLAB_M
#define WDB_SLAB_M 4
#endif
#define WDB_SLAB_CHUNKS (WDB_NWARPS * WDB_SLAB_M)
// L2 residency hints (256-bit loads only): phase 1 parks the slab (evict_last), phase 2 releases it
// (evict_first), the output streams past it (evict_first policy): 7-12 % (ncu showed only ~10 % of
// the phase-2 sectors hitting the L2 without them).  Bulk-prefetching the NEXT slab into the L2
// (cp.async.bulk.prefetch.L2, so that phase 1 reads L2 too) was tried and is 2x slower at every
// slab size and prefetch granularity: two slabs per CTA no longer fit next to the output stream.
// Also tried: leaving the staged survivors through one bulk shared->global copy per chunk
// (cp.async.bulk, UBLKCP.G.S) instead of the LDS/STG loop: 8 % slower at 1 % and 50 % selectivity,
// equal at 99 % (the wait for the previous chunk's copy and the proxy fence cost more than the loop).
// And: staging the survivors already in phase 1 when a slab is sparse (so that phase 2, the second
// read, is skipped): only 3 % faster at 1 % selectivity, 14 % at 0.1 %, 3 % slower at 50 % -- even
// with phase 2 empty the kernel streams 4 GB in 0.84 ms (73 % of the copy peak): what bounds it at low
// selectivity is the count phase itself (4 vector loads in flight per warp, a barrier and a
// look-back per slab), not the re-read.
#if WDB_L2_HINTS && WDB_VEC == 8 && WDB_ALIGNED
#define WDB_L2_PARK 3
#define WDB_L2_DROP 2
#else
#define WDB_L2_PARK WDB_LD_HINT
#define WDB_L2_DROP WDB_LD_HINT
#endif

template <int H>
__device__ __forceinline__ u32 wdb_chunk_flags(const wdb_cols &C, const i64 n, const i64 chunk, const u32 lane, const float wdb_tau,
                                               u32 (&flags)[WDB_UNROLL], float (&vals)[WDB_UNROLL][WDB_VEC],
#if WDB_NOUT == 2
                                               float (&vals2)[WDB_UNROLL][WDB_VEC],
#endif
                                               const bool want_vals) {
  const i64 row0 = chunk * WDB_WARP_ROWS + (i64)lane * WDB_VEC;
  u32 cnt = 0;
  if ((chunk + 1) * WDB_WARP_ROWS <= n) {
    wdb_rows R[WDB_UNROLL];
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u) wdb_load_rows_h<H>(C, row0 + (i64)u * WDB_SLAB_ROWS, R[u]);
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u) {
      u32 m = 0;
#pragma unroll
      for (int j = 0; j < WDB_VEC; ++j) {
        m |= (WDB_KEEP(R[u], j) ? 1u : 0u) << j;
        if (want_vals) {
          vals[u][j] = WDB_EXPR(R[u], j);
#if WDB_NOUT == 2
          vals2[u][j] = WDB_EXPR2(R[u], j);
#endif
        }
      }
      flags[u] = m;
      cnt += __popc(m);
    }
  } else {
#pragma unroll
    for (int u = 0; u < WDB_UNROLL; ++u) {
      u32 m = 0;
#pragma unroll
      for (int j = 0; j < WDB_VEC; ++j) {
        const i64 row = row0 + (i64)u * WDB_SLAB_ROWS + j;
        vals[u][j] = 0.0f;
#if WDB_NOUT == 2
        vals2[u][j] = 0.0f;
#endif
        if (row < n) {
          wdb_rows T1;
          wdb_load_row1(C, row, T1, 0);
          if (WDB_KEEP(T1, 0)) {
            m |= 1u << j;
            vals[u][j] = WDB_EXPR(T1, 0);
#if WDB_NOUT == 2
            vals2[u][j] = WDB_EXPR2(T1, 0);
#endif
          }
        }
      }
      flags[u] = m;
      cnt += __popc(m);
    }
  }
  return cnt;   // this lane's survivors in the chunk
}

extern "C" __global__ void __launch_bounds__(WDB_BLOCK, WDB_MIN_CTAS)
wdb_compact_l2(const wdb_cols C, float *__restrict__ out, float *__restrict__ out2, const i64 n,
               u64 *__restrict__ status, u32 *__restrict__ ticket, i64 *__restrict__ out_count, const i64 nslabs,
               const i64 nchunks, const float wdb_tau, const i64 out_cap WDB_SEL_PARAM) {
#if WDB_AUTO_SEL
  if (WDB_SEL_STAGED) return;                        // selective filter: the staged two-pass pipeline does the work
#endif
  __shared__ float s_stage[WDB_NWARPS][WDB_WARP_ROWS];
#if WDB_NOUT == 2
  __shared__ float s_stage2[WDB_NWARPS][WDB_WARP_ROWS];
#endif
  __shared__ u32 s_wcount[WDB_NWARPS];
  __shared__ i64 s_base;
  __shared__ u32 s_slab;
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  while (true) {
    if (threadIdx.x == 0) s_slab = atomicAdd(ticket, 1u);
    __syncthreads();                                            // (1)
    const i64 slab = (i64)s_slab;
    if (slab >= nslabs) break;
    const i64 chunk0 = slab * WDB_SLAB_CHUNKS + (i64)warp * WDB_SLAB_M;
    u32 flags[WDB_UNROLL];
    float vals[WDB_UNROLL][WDB_VEC];
#if WDB_NOUT == 2
    float vals2[WDB_UNROLL][WDB_VEC];
#endif
    // phase 1: count (HBM -> L2)
    u32 ccount[WDB_SLAB_M];
    u32 wtotal = 0;
#pragma unroll
    for (int m = 0; m < WDB_SLAB_M; ++m) {
      u32 c = 0;
#if WDB_PF_NEXT
      // the warp's chunk m + WDB_PF_NEXT starts its way from HBM to the L2 now (no registers involved):
      // its loads later see L2 latency, which multiplies the bytes a warp keeps in flight during the count phase
#pragma unroll
      for (int a = (m == 0 ? 1 : WDB_PF_NEXT); a <= WDB_PF_NEXT; ++a)
        if (m + a < WDB_SLAB_M && (chunk0 + m + a + 1) * WDB_WARP_ROWS <= n) {
          const i64 prow = (chunk0 + m + a) * WDB_WARP_ROWS + (i64)lane * WDB_VEC;
#pragma unroll
          for (int u = 0; u < WDB_UNROLL; ++u) wdb_prefetch_rows(C, prow + (i64)u * WDB_SLAB_ROWS);
        }
#endif
      if (chunk0 + m < nchunks) {
#if WDB_NOUT == 2
        c = wdb_chunk_flags<WDB_L2_PARK>(C, n, chunk0 + m, lane, wdb_tau, flags, vals, vals2, false);
#else
        c = wdb_chunk_flags<WDB_L2_PARK>(C, n, chunk0 + m, lane, wdb_tau, flags, vals, false);
#endif
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(WDB_FULL_MASK, c, o);
      ccount[m] = c;
      wtotal += c;
    }
    if (lane == 0) s_wcount[warp] = wtotal;
    __syncthreads();                                            // (2)
    u32 woff = 0, ttotal = 0;
#pragma unroll
    for (int w = 0; w < WDB_NWARPS; ++w) {
      const u32 c = s_wcount[w];
      woff += (w < (int)warp) ? c : 0u;
      ttotal += c;
    }
    if (warp == 0) {
      i64 excl = 0;
      if (slab == 0) {
        if (lane == 0) wdb_st_status(&status[0], (WDB_ST_PREFIX << 62) | (u64)ttotal);
      } else {
        if (lane == 0) wdb_st_status(&status[slab], (WDB_ST_AGG << 62) | (u64)ttotal);
        excl = wdb_lookback(status, slab, lane);
        if (lane == 0) wdb_st_status(&status[slab], (WDB_ST_PREFIX << 62) | (u64)(excl + (i64)ttotal));
      }
      if (lane == 0) {
        s_base = excl;
        if (slab == nslabs - 1) *out_count = excl + (i64)ttotal;
      }
    }
    __syncthreads();                                            // (3)
    // phase 2: re-read (L2), rank, stage per warp, write
    i64 g0 = s_base + (i64)woff;
#pragma unroll
    for (int m = 0; m < WDB_SLAB_M; ++m) {
      if (chunk0 + m >= nchunks || ccount[m] == 0u) continue;   // warp-uniform
#if WDB_NOUT == 2
      wdb_chunk_flags<WDB_L2_DROP>(C, n, chunk0 + m, lane, wdb_tau, flags, vals, vals2, true);
#else
      wdb_chunk_flags<WDB_L2_DROP>(C, n, chunk0 + m, lane, wdb_tau, flags, vals, true);
#endif
      u32 total = 0;
#pragma unroll
      for (int u = 0; u < WDB_UNROLL; ++u) {
        u32 pre, tot;
        wdb_warp_rank(__popc(flags[u]), lane, pre, tot);
        u32 pos = total + pre;
#pragma unroll
        for (int j = 0; j < WDB_VEC; ++j)
          if ((flags[u] >> j) & 1u) {
            s_stage[warp][pos] = vals[u][j];
#if WDB_NOUT == 2
            s_stage2[warp][pos] = vals2[u][j];
#endif
            ++pos;
          }
        total += tot;
      }
      __syncwarp();
      const int mis = (int)(g0 & 31);
      for (int i = (int)lane - mis; i < (int)total; i += 32)
        if (i >= 0 && g0 + i < out_cap) {
#if WDB_L2_HINTS
          wdb_st_f32_stream(out + g0 + i, s_stage[warp][i]);
#if WDB_NOUT == 2
          wdb_st_f32_stream(out2 + g0 + i, s_stage2[warp][i]);
#endif
#else
          out[g0 + i] = s_stage[warp][i];
#if WDB_NOUT == 2
          out2[g0 + i] = s_stage2[warp][i];
#endif
#endif
        }
      __syncwarp();
      g0 += total;
    }
  }
}

// ---- variant 4: ONE read, software-pipelined look-back (wdb_compact_sp) -------------------------
// Variant 3 reads every slab twice (HBM, then L2) and pays three block barriers and a look-back on the
// critical path of every slab; ncu shows it issue- and barrier-bound, not bandwidth-bound.  Here a
// chunk's survivors are evaluated, ranked and staged in shared memory in the SAME pass that loads
// them, and nothing waits for the global offset: WDB_NWARPS worker warps keep streaming chunks
// while one extra warp (the scanner) does nothing but turn the workers' counts of round r into
// global offsets (publish aggregate, decoupled look-back, publish prefix).  A worker copies the
// staged survivors of round r - (S - 1) out after it has staged round r (S = WDB_SP_STAGES staging
// buffers per warp), so the look-back of a round has S - 1 rounds of loads to hide behind.
// Workers and scanner meet only through named barriers in the producer / consumer pattern of the
// PTX ISA (barrier.arrive on one side, barrier.sync on the other): no __syncthreads, no ticket.
// Slabs (one chunk per worker warp) are assigned round-robin, slab = round * gridDim.x + blockIdx.x,
// which is deadlock-free because the grid never exceeds what is co-resident (host: occupancy).
#ifndef WDB_SP_STAGES
#define WDB_SP_STAGES 2
#endif
#ifndef WDB_SP_SCAN
#define WDB_SP_SCAN 1
#endif
#define WDB_SP_THREADS (WDB_BLOCK + 32)
__device__ __forceinline__ void wdb_bar_arrive(const int id) { asm volatile("barrier.arrive %0, %1;" :: "r"(id), "n"(WDB_SP_THREADS) : "memory"); }
__device__ __forceinline__ void wdb_bar_sync(const int id) { asm volatile("barrier.sync %0, %1;" :: "r"(id), "n"(WDB_SP_THREADS) : "memory"); }
#define WDB_SP_BAR_CNT(b) (1 + (b))
#define WDB_SP_BAR_BASE(b) (1 + WDB_SP_STAGES + (b))

extern __shared__ __align__(16) unsigned char wdb_sp_smem[];
// layout: goff i64[S][NW] | wcount u32[S][NW] | stage f32[S][NW][WARP_ROWS] (| stage2 ...)
#define WDB_SP_GOFF(b, w) (reinterpret_cast<i64 *>(wdb_sp_smem)[(b) * WDB_NWARPS + (w)])
#define WDB_SP_WCNT(b, w) (reinterpret_cast<u32 *>(wdb_sp_smem + 8 * WDB_SP_STAGES * WDB_NWARPS)[(b) * WDB_NWARPS + (w)])
#define WDB_SP_HDR ((12 * WDB_SP_STAGES * WDB_NWARPS + 15) / 16 * 16)
#define WDB_SP_STAGE(b, w) (reinterpret_cast<float *>(wdb_sp_smem + WDB_SP_HDR) + ((size_t)(b) * WDB_NWARPS + (w)) * WDB_WARP_ROWS)
#define WDB_SP_STAGE2(b, w) (WDB_SP_STAGE(b, w) + (size_t)WDB_SP_STAGES * WDB_NWARPS * WDB_WARP_ROWS)

__device__ __forceinline__ void wdb_sp_copy_out(const int b, const u32 warp, const u32 lane, float *__restrict__ out, float *__restrict__ out2, const i64 out_cap) {
  const i64 g0 = WDB_SP_GOFF(b, warp);
  const int total = (int)WDB_SP_WCNT(b, warp);
  const float *st = WDB_SP_STAGE(b, warp);
#if WDB_NOUT == 2
  const float *st2 = WDB_SP_STAGE2(b, warp);
#endif
  const int mis = (int)(g0 & 31);             // every store instruction of the loop covers one aligned 128-byte line
  for (int i = (int)lane - mis; i < total; i += 32)
    if (i >= 0 && g0 + i < out_cap) {
      out[g0 + i] = st[i];
#if WDB_NOUT == 2
      out2[g0 + i] = st2[i];
#endif
    }
}

extern "C" __global__ void __launch_bounds__(WDB_SP_THREADS, WDB_MIN_CTAS)
wdb_compact_sp(const wdb_cols C, float *__restrict__ out, float *__restrict__ out2, const i64 n, u64 *__restrict__ status, u64 *__restrict__ rbase,
               u32 *__restrict__ ticket, i64 *__restrict__ out_count, const i64 nslabs, const i64 nchunks, const float wdb_tau, const i64 out_cap) {
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  // Slab of round r.  WDB_SP_SCAN == 1: static round-robin (round-synchronous prefix).  Otherwise the first
  // S rounds are static and every later one is a ticket the scanner drew S rounds earlier (slabs are
  // then handed out in time order: whatever a slab's look-back waits for was started before it, so a
  // late CTA delays nobody but itself) and parked in s_slab[r % 2S]; the workers pick it up behind the
  // BASE barrier of round r - S, which they pass before they start round r.
  __shared__ i64 s_slab[2 * WDB_SP_STAGES];
#if WDB_SP_SCAN == 1
#define WDB_SP_SLAB(r) ((i64)blockIdx.x + (r) * (i64)gridDim.x)
#else
#define WDB_SP_SLAB(r) ((r) < WDB_SP_STAGES ? (i64)blockIdx.x + (r) * (i64)gridDim.x : s_slab[(r) % (2 * WDB_SP_STAGES)])
#endif
  if (warp == WDB_NWARPS) {
    // ---- scanner warp: counts of round r -> global offsets of round r
    for (i64 r = 0;; ++r) {
      const int b = (int)(r % WDB_SP_STAGES);
      const i64 slab = WDB_SP_SLAB(r);
      if (slab >= nslabs) break;
#if WDB_SP_SCAN != 1
      if (lane == 0) s_slab[(r + WDB_SP_STAGES) % (2 * WDB_SP_STAGES)] = WDB_SP_STAGES * (i64)gridDim.x + (i64)atomicAdd(ticket, 1u);   // published by this round's BASE arrive
#endif
      wdb_bar_sync(WDB_SP_BAR_CNT(b));
      const u32 c = lane < WDB_NWARPS ? WDB_SP_WCNT(b, lane) : 0u;
      u32 pre, ttotal;
      wdb_warp_rank(c, lane, pre, ttotal);
      i64 excl = 0;
#if WDB_SP_SCAN == 1
      // Round-synchronous prefix: the offset of (r, c) is rbase[r] + the aggregates of (r, 0 .. c-1), all read
      // in parallel.  Measured SLOWER than the decoupled look-back: every round becomes a grid-wide
      // barrier in disguise, and with round-robin slabs one late CTA stalls all the others.
      {
        const i64 slab0 = r * gridDim.x;
        if (lane == 0) wdb_st_status(&status[slab], (WDB_ST_AGG << 62) | (u64)ttotal);
        u64 sum = 0;
        const i64 c0 = (i64)blockIdx.x;
        for (i64 i = lane; i < c0; i += 128) {
          u64 st[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) st[q] = (i + 32 * q < c0) ? wdb_ld_status(&status[slab0 + i + 32 * q]) : (WDB_ST_AGG << 62);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            while ((st[q] >> 62) == 0ull) st[q] = wdb_ld_status(&status[slab0 + i + 32 * q]);
            sum += st[q] & WDB_ST_MASK;
          }
        }
        u64 base = 0;
        if (r > 0) {
          if (lane == 0) {
            u64 st;
            do { st = wdb_ld_status(&rbase[r]); } while ((st >> 62) == 0ull);
            base = st & WDB_ST_MASK;
          }
          base = __shfl_sync(WDB_FULL_MASK, base, 0);
        }
        excl = (i64)(base + wdb_warp_sum64(sum));
        if (blockIdx.x == gridDim.x - 1 && lane == 0) wdb_st_status(&rbase[r + 1], (WDB_ST_AGG << 62) | (u64)(excl + (i64)ttotal));
      }
#else
      if (slab == 0) {
        if (lane == 0) wdb_st_status(&status[0], (WDB_ST_PREFIX << 62) | (u64)ttotal);
      } else {
        if (lane == 0) wdb_st_status(&status[slab], (WDB_ST_AGG << 62) | (u64)ttotal);
        excl = wdb_lookback(status, slab, lane);
        if (lane == 0) wdb_st_status(&status[slab], (WDB_ST_PREFIX << 62) | (u64)(excl + (i64)ttotal));
      }
#endif
      if (lane < WDB_NWARPS) WDB_SP_GOFF(b, lane) = excl + (i64)pre;
      if (lane == 0 && slab == nslabs - 1) *out_count = excl + (i64)ttotal;
      __syncwarp();
      wdb_bar_arrive(WDB_SP_BAR_BASE(b));
    }
    return;
  }
  // ---- worker warps
  i64 nr = 0;
  for (i64 r = 0;; ++r) {
    const int b = (int)(r % WDB_SP_STAGES);
    const i64 slab = WDB_SP_SLAB(r);
    if (slab >= nslabs) break;
    nr = r + 1;
    const i64 chunk = slab * WDB_NWARPS + warp;
    u32 total = 0;
    if (chunk < nchunks) {
      u32 flags[WDB_UNROLL];
      float vals[WDB_UNROLL][WDB_VEC];
#if WDB_NOUT == 2
      float vals2[WDB_UNROLL][WDB_VEC];
      wdb_chunk_flags<WDB_LD_HINT>(C, n, chunk, lane, wdb_tau, flags, vals, vals2, true);
      float *st2 = WDB_SP_STAGE2(b, warp);
#else
      wdb_chunk_flags<WDB_LD_HINT>(C, n, chunk, lane, wdb_tau, flags, vals, true);
#endif
      float *st = WDB_SP_STAGE(b, warp);
#pragma unroll
      for (int u = 0; u < WDB_UNROLL; ++u) {
        u32 pre, tot;
        wdb_warp_rank(__popc(flags[u]), lane, pre, tot);
        u32 pos = total + pre;
#pragma unroll
        for (int j = 0; j < WDB_VEC; ++j)
          if ((flags[u] >> j) & 1u) {
            st[pos] = vals[u][j];
#if WDB_NOUT == 2
            st2[pos] = vals2[u][j];
#endif
            ++pos;
          }
        total += tot;
      }
    }
    if (lane == 0) WDB_SP_WCNT(b, warp) = total;
    __syncwarp();
    wdb_bar_arrive(WDB_SP_BAR_CNT(b));
    if (r >= WDB_SP_STAGES - 1) {
      const int bo = (int)((r - (WDB_SP_STAGES - 1)) % WDB_SP_STAGES);
      wdb_bar_sync(WDB_SP_BAR_BASE(bo));
      wdb_sp_copy_out(bo, warp, lane, out, out2, out_cap);
      __syncwarp();
    }
  }
  for (i64 ro = nr > (WDB_SP_STAGES - 1) ? nr - (WDB_SP_STAGES - 1) : 0; ro < nr; ++ro) {   // drain
    const int bo = (int)(ro % WDB_SP_STAGES);
    wdb_bar_sync(WDB_SP_BAR_BASE(bo));
    wdb_sp_copy_out(bo, warp, lane, out, out2, out_cap);
    __syncwarp();
  }
}
#endif
