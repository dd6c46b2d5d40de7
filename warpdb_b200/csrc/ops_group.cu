// ops_group.cu -- hash GROUP BY: aggregation-table object, consume launch (kernels/group.cuh via
// NVRTC), merge of partial aggregates and ordered export.
//
// Replaces jit_group_sum (src/jit.cpp:179-246: one thread, O(N*G) linear search, float32 running
// sums) and the std::map<int,AggData> loop of src/warpdb.cpp:373-437 (fp64 sum/count/min/max, key
// ascending); jit_sort_pairs on the group arrays (src/warpdb.cpp:370-371) is the ordered export.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstring>

#include "core.hpp"
#include "agg.hpp"

namespace wdb {
template <class K>
int radix_sort(Device *d, cudaStream_t s, K *keys, K *tmp_keys, unsigned *pay, unsigned *tmp_pay, long long n, int key_bits);
}

using namespace wdb;

// ---- static kernels ------------------------------------------------------------------------------
__global__ void agg_init_kernel(wdb_table T, long long slots) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += (long long)gridDim.x * blockDim.x) {
    T.keys[i] = WDB_KEY_EMPTY;
    T.sums[i] = 0.0;
    T.counts[i] = 0ull;
    T.mins[i] = WDB_ENC_PLUS_INF;
    T.maxs[i] = WDB_ENC_MINUS_INF;
    T.first[i] = 0x7fffffffffffffffll;
    if (i < 8) T.meta[i] = 0u;
  }
}

template <int NEEDS>
__global__ void agg_merge_kernel(wdb_table T, const int *__restrict__ keys, const double *__restrict__ sums,
                                 const long long *__restrict__ counts, const double *__restrict__ mins,
                                 const double *__restrict__ maxs, const long long *__restrict__ first, long long m) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
    if ((NEEDS & WDB_NEED_FIRST_BIT) == 0) {   // a live direct-addressed side table owns the keys of its range
      const unsigned di = (unsigned)keys[i] - (unsigned)T.dlo;
      if (di < T.dspan) {
        wdb_dense_add<NEEDS & 7>(T, di, (NEEDS & WDB_NEED_SUM_BIT) ? sums[i] : 0.0, (NEEDS & WDB_NEED_CNT_BIT) ? (unsigned long long)counts[i] : 0ull,
                                 (NEEDS & WDB_NEED_MINMAX_BIT) ? wdb_f64_enc(mins[i]) : 0, (NEEDS & WDB_NEED_MINMAX_BIT) ? wdb_f64_enc(maxs[i]) : 0);
        continue;
      }
    }
    const long long s = wdb_table_slot(T, keys[i]);
    if (s < 0) continue;
    wdb_table_add<NEEDS>(T, s, (NEEDS & WDB_NEED_SUM_BIT) ? sums[i] : 0.0, (NEEDS & WDB_NEED_CNT_BIT) ? (unsigned long long)counts[i] : 0ull,
                         (NEEDS & WDB_NEED_MINMAX_BIT) ? wdb_f64_enc(mins[i]) : 0, (NEEDS & WDB_NEED_MINMAX_BIT) ? wdb_f64_enc(maxs[i]) : 0,
                         (NEEDS & WDB_NEED_FIRST_BIT) ? first[i] : 0);
  }
}

// occupied slots -> (sort key, slot) pairs.  order: 0 first appearance, 1 key asc, 2 key desc
__global__ void agg_collect_kernel(wdb_table T, long long slots, int order, unsigned long long *__restrict__ sort_keys,
                                   unsigned *__restrict__ sort_slots, unsigned long long *__restrict__ counter) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += (long long)gridDim.x * blockDim.x) {
    const bool special = i == slots - 1;
    const bool used = special ? (T.meta[2] != 0u) : (T.keys[i] != WDB_KEY_EMPTY);
    if (!used) continue;
    const int key = special ? WDB_KEY_EMPTY : T.keys[i];
    unsigned long long k;
    if (order == 0) k = (unsigned long long)T.first[i];
    else {
      k = (unsigned long long)((unsigned)key ^ 0x80000000u);
      if (order == 2) k = 0xffffffffull - k;
    }
    const unsigned long long pos = atomicAdd(counter, 1ull);
    sort_keys[pos] = k;
    sort_slots[pos] = (unsigned)i;
  }
}

__global__ void agg_emit_kernel(wdb_table T, long long slots, const unsigned *__restrict__ sorted_slots, long long g, int agg,
                                int *__restrict__ o_keys, float *__restrict__ o_vals, double *__restrict__ o_sums,
                                long long *__restrict__ o_counts, double *__restrict__ o_mins, double *__restrict__ o_maxs,
                                long long *__restrict__ o_first) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < g; i += (long long)gridDim.x * blockDim.x) {
    const unsigned s = sorted_slots[i];
    const int key = (s == slots - 1) ? WDB_KEY_EMPTY : T.keys[s];
    const double sum = T.sums[s];
    const double cnt = (double)T.counts[s];
    const double mn = wdb_f64_dec(T.mins[s]), mx = wdb_f64_dec(T.maxs[s]);
    if (o_keys) o_keys[i] = key;
    if (o_vals) {  // src/warpdb.cpp:429-435: double results narrowed to float
      float v;
      switch (agg) {
      case WDB_SUM: v = (float)sum; break;
      case WDB_AVG: v = (float)(sum / cnt); break;
      case WDB_COUNT: v = (float)cnt; break;
      case WDB_MIN: v = (float)mn; break;
      default: v = (float)mx; break;
      }
      o_vals[i] = v;
    }
    if (o_sums) o_sums[i] = sum;
    if (o_counts) o_counts[i] = (long long)T.counts[s];
    if (o_mins) o_mins[i] = mn;
    if (o_maxs) o_maxs[i] = mx;
    if (o_first) o_first[i] = T.first[s];
  }
}

// ---- direct-addressed side table -----------------------------------------------------------------
__global__ void dense_init_kernel(wdb_table T, int needs, long long span) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < span; i += (long long)gridDim.x * blockDim.x) {
    if (needs & WDB_NEED_SUM_BIT) reinterpret_cast<unsigned long long *>(T.dsums)[i] = WDB_DENSE_EMPTY;
    if (needs & WDB_NEED_CNT_BIT) T.dcnts[i] = 0ull;
    if (needs & WDB_NEED_MINMAX_BIT) { T.dmins[i] = WDB_ENC_PLUS_INF; T.dmaxs[i] = WDB_ENC_MINUS_INF; }
  }
}
constexpr int kDenseBlock = 256, kDenseItems = 8, kDenseTile = kDenseBlock * kDenseItems;
// pass 1: present entries per tile of 2048 indices
template <int NEEDS> __global__ void __launch_bounds__(kDenseBlock) dense_count_kernel(wdb_table T, unsigned *__restrict__ tile_counts) {
  const long long base = (long long)blockIdx.x * kDenseTile;
  unsigned c = 0;
#pragma unroll
  for (int k = 0; k < kDenseItems; ++k) {
    const long long i = base + (long long)k * kDenseBlock + threadIdx.x;
    if (i < (long long)T.dspan && wdb_dense_present<NEEDS>(T, i)) ++c;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ unsigned s[kDenseBlock / 32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < kDenseBlock / 32; ++w) t += s[w];
    tile_counts[blockIdx.x] = t;
  }
}
// pass 2 (one CTA): exclusive scan of the tile counts; total -> *total (and, as a signed 64-bit, -> *total_out)
__global__ void __launch_bounds__(1024) dense_scan_kernel(const unsigned *__restrict__ tile_counts, unsigned long long *__restrict__ tile_offsets,
                                                           long long ntiles, unsigned long long *__restrict__ total, long long *__restrict__ total_out) {
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_carry;
  if (threadIdx.x == 0) s_carry = 0ull;
  __syncthreads();
  for (long long base = 0; base < ntiles; base += 1024) {
    const long long i = base + threadIdx.x;
    const unsigned long long v = i < ntiles ? tile_counts[i] : 0ull;
    unsigned long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned long long w = s_warp[threadIdx.x], ww = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, ww, o);
        if (threadIdx.x >= o) ww += y;
      }
      s_warp[threadIdx.x] = ww - w;   // exclusive over warps
    }
    __syncthreads();
    const unsigned long long incl = s_carry + s_warp[threadIdx.x >> 5] + x;
    if (i < ntiles) tile_offsets[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *total = s_carry;
    if (total_out) *total_out = (long long)s_carry;
  }
}
// pass 3: ranked emit in key order (ascending, or descending when desc != 0); `to_table` folds the
// entries into the hash table instead (flush).  Groups beyond `cap` are dropped (the caller compares
// the total with cap).
template <int NEEDS> __global__ void __launch_bounds__(kDenseBlock)
dense_emit_kernel(wdb_table T, const unsigned long long *__restrict__ tile_offsets, const unsigned long long *__restrict__ total, int desc, int agg,
                  int to_table, long long cap, int *__restrict__ o_keys, float *__restrict__ o_vals, double *__restrict__ o_sums,
                  long long *__restrict__ o_counts, double *__restrict__ o_mins, double *__restrict__ o_maxs) {
  __shared__ unsigned s_wbase[kDenseBlock / 32];
  const long long base = (long long)blockIdx.x * kDenseTile;
  // thread t owns kDenseItems CONSECUTIVE indices so that ranks follow key order
  bool pres[kDenseItems];
  unsigned c = 0;
#pragma unroll
  for (int k = 0; k < kDenseItems; ++k) {
    const long long i = base + (long long)threadIdx.x * kDenseItems + k;
    pres[k] = i < (long long)T.dspan && wdb_dense_present<NEEDS>(T, i);
    c += pres[k] ? 1u : 0u;
  }
  unsigned x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) s_wbase[threadIdx.x >> 5] = x;
  __syncthreads();
  unsigned wbase = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += s_wbase[w];
  unsigned long long pos = tile_offsets[blockIdx.x] + wbase + (x - c);
  const unsigned long long g = *total;
#pragma unroll
  for (int k = 0; k < kDenseItems; ++k) {
    if (!pres[k]) continue;
    const long long i = base + (long long)threadIdx.x * kDenseItems + k;
    const int key = (int)((unsigned)T.dlo + (unsigned)i);
    const double sum = (NEEDS & WDB_NEED_SUM_BIT) ? T.dsums[i] : 0.0;
    const unsigned long long cnt = (NEEDS & WDB_NEED_CNT_BIT) ? T.dcnts[i] : 0ull;
    const long long mn = (NEEDS & WDB_NEED_MINMAX_BIT) ? T.dmins[i] : 0, mx = (NEEDS & WDB_NEED_MINMAX_BIT) ? T.dmaxs[i] : 0;
    if (to_table) {
      const long long sl = wdb_table_slot(T, key);
      if (sl >= 0) wdb_table_add<NEEDS>(T, sl, sum, cnt, mn, mx, 0);
    } else {
      const unsigned long long o = desc ? g - 1ull - pos : pos;
      if (o < (unsigned long long)cap) {
        if (o_keys) o_keys[o] = key;
        if (o_vals) {   // src/warpdb.cpp:429-435: double results narrowed to float
          float v;
          switch (agg) {
          case WDB_SUM: v = (float)sum; break;
          case WDB_AVG: v = (float)(sum / (double)((NEEDS & WDB_NEED_CNT_BIT) ? cnt : 1ull)); break;
          case WDB_COUNT: v = (float)(double)cnt; break;
          case WDB_MIN: v = (float)wdb_f64_dec(mn); break;
          default: v = (float)wdb_f64_dec(mx); break;
          }
          o_vals[o] = v;
        }
        if (o_sums) o_sums[o] = sum;
        if (o_counts) o_counts[o] = (long long)cnt;
        if (o_mins) o_mins[o] = wdb_f64_dec(mn);
        if (o_maxs) o_maxs[o] = wdb_f64_dec(mx);
      }
    }
    ++pos;
  }
}
// hash table -> side table (cross-GPU merge: every GPU's whole partial must sit in the side table);
// entries outside its range are counted in meta[4]
template <int NEEDS> __global__ void agg_hash_to_dense_kernel(wdb_table T, long long slots) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += (long long)gridDim.x * blockDim.x) {
    const bool special = i == slots - 1;
    const bool used = special ? (T.meta[2] != 0u) : (T.keys[i] != WDB_KEY_EMPTY);
    if (!used) continue;
    const int key = special ? WDB_KEY_EMPTY : T.keys[i];
    const unsigned di = (unsigned)key - (unsigned)T.dlo;
    if (di < T.dspan) wdb_dense_add<NEEDS>(T, di, (NEEDS & WDB_NEED_SUM_BIT) ? T.sums[i] : 0.0, (NEEDS & WDB_NEED_CNT_BIT) ? T.counts[i] : 0ull,
                                           (NEEDS & WDB_NEED_MINMAX_BIT) ? T.mins[i] : 0, (NEEDS & WDB_NEED_MINMAX_BIT) ? T.maxs[i] : 0);
    else atomicAdd(&T.meta[4], 1u);
  }
}

namespace wdb {

int needs_for_agg(int agg) {
  switch (agg) {
  case WDB_SUM: return WDB_NEED_SUM_BIT;
  case WDB_AVG: return WDB_NEED_SUM_BIT | WDB_NEED_CNT_BIT;
  case WDB_COUNT: return WDB_NEED_CNT_BIT;
  case WDB_MIN: case WDB_MAX: return WDB_NEED_MINMAX_BIT;
  }
  return 0;
}

static unsigned grid_for(Device *d, long long n) {
  return (unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)d->num_sms * 16));
}

// warp-private accumulators (wdb_group_wp): bytes per key and whether at least 4 warps per SM fit
constexpr int64_t kMaxDynSmem = 232448 - 64;   // sm_100: 227 KB per CTA
static int64_t wp_acc_bytes(int needs) {   // accumulators of one id (without the arbitration tag)
  return ((needs & WDB_NEED_SUM_BIT) ? 8 : 0) + ((needs & WDB_NEED_CNT_BIT) ? 4 : 0) + ((needs & WDB_NEED_MINMAX_BIT) ? 16 : 0) +
         ((needs & WDB_NEED_FIRST_BIT) ? 8 : 0);
}
// Accumulator scheme of the warp-private kernel for a key range of `span` ids:
//   2  lane-private copies (entry id * 32 + lane: conflict-free, nothing to arbitrate) while at least 8 warps of them fit an SM
//   0  one copy per warp, duplicates within a warp step arbitrated through a tag array
//   1  one copy per warp, duplicates found with MATCH.ANY (measured 2.6x slower than 0; kept as an option)
static int wp_mode_for(int needs, int64_t span) {
  if (span > 0 && opt("group.lane_private", 1) && kMaxDynSmem / (wp_acc_bytes(needs) * 32 * ((span + 7) / 8 * 8)) >= opt("group.lane_min_warps", 8)) return 2;
  return (int)opt("group.wp_mode", 0);
}
static int64_t wp_bytes_per_id(int needs, int mode) { return mode == 2 ? 32 * wp_acc_bytes(needs) : (mode == 1 ? 0 : 4) + wp_acc_bytes(needs); }
static bool wp_fits(int needs, int64_t span) {
  if (span <= 0 || span > opt("group.wp_max_span", 4096)) return false;
  return kMaxDynSmem / (wp_bytes_per_id(needs, wp_mode_for(needs, span)) * ((span + 7) / 8 * 8)) >= 4;   // fewer warps cannot hide the shared-memory latency
}

struct GroupPlan { GenSpec spec; int block, unroll, vec, smem_slots, wp_ids; size_t smem_bytes; const char *entry; };

static int plan_group(const wdb_col_t *cols, int ncols, const char *val, const char *key, const char *cond, int needs,
                      int64_t cap_hint, KeyRange range, bool dense, bool use_wp, bool check_alignment, GroupPlan *p) {
  const bool has_cond = cond && *cond;
  GenSpec &spec = p->spec;
  spec.kind = "group";
  spec.used = find_used_columns(cols, ncols, {val, key, has_cond ? cond : ""});
  for (const auto &u : spec.used)
    if (dtype_size(u.dtype) == 0) return fail("column %s has a non-numeric type and cannot be read on the GPU", u.name.c_str());
  p->block = (int)opt("group.block", 512);   // profiles/r01_diag_group_tuning_1e9.jsonl
  p->unroll = (int)opt("group.unroll", 1);
  p->vec = (int)opt("group.vec", 4);
  if (p->vec != 4 && p->vec != 8) return fail("group.vec must be 4 or 8");
  // Small cardinalities with a known key range (SUM/COUNT/AVG): warp-private, directly indexed
  // accumulators without shared-memory atomics (wdb_group_wp).  Otherwise shared pre-aggregation
  // with atomics pays while the distinct keys fit the CTA's table; beyond that every row misses it
  // and the probes are wasted work.
  const int64_t span = range.known ? range.hi - range.lo + 1 : -1;
  int64_t expected = cap_hint / 2;                           // cap_hint = table capacity = 2 x expected groups
  if (span > 0) expected = std::min(expected, span);          // an integer key cannot form more groups than its range holds
  const int64_t kMaxDyn = kMaxDynSmem;
  int64_t wp = 0;
  const int wp_mode = wp_mode_for(needs, span);
  const int64_t wp_per_id = wp_bytes_per_id(needs, wp_mode);
  if (use_wp && wp_fits(needs, span)) wp = (span + 7) / 8 * 8;
  int64_t slots = opt("group.smem_slots", -1);
  if (dense && !use_wp) {   // in-range rows go straight to the direct-addressed table: one RED each, nothing to pre-aggregate
    slots = 0;
    // profiles/r02_sweep_group10m.jsonl: few, small CTAs keep fewer REDs in flight per SM and the table slice
    // in the L2 (256 threads x 128-bit loads, 2 CTAs per SM: 6.25 ms per 1e9 rows at 10 M keys against 6.72 ms)
    p->block = (int)opt("group.dense_block", 256);
    p->unroll = (int)opt("group.dense_unroll", 1);
    p->vec = (int)opt("group.dense_vec", 4);
  }
  if (slots < 0) {
    slots = 0;
    if (expected <= 2048) { slots = 1024; while (slots < 8 * expected && slots < 8192) slots <<= 1; }   // low load factor: short probe chains
    else if (expected <= 4096) {
      // 2 K - 4 K keys of unknown range: one 16 K-slot table per SM and 1 024 threads still beat the
      // global table (184 vs 68 Grows/s at 3 K keys; from ~6 K keys on the global table wins)
      const int64_t per_slot_big = 8 + 4 + ((needs & WDB_NEED_CNT_BIT) ? 4 : 0) + ((needs & WDB_NEED_MINMAX_BIT) ? 16 : 0) + ((needs & WDB_NEED_FIRST_BIT) ? 8 : 0);
      if (per_slot_big * 16384 <= kMaxDyn) { slots = 16384; p->block = (int)opt("group.block", 1024); }
    }
  }
  if (slots & (slots - 1)) return fail("group.smem_slots must be a power of two");
  int wp_ilp = 1;
  if (wp > 0) {
    slots = 0;
    const int64_t per_id = wp_per_id;
    int warps = (int)std::min<int64_t>(opt("group.wp_warps", 16), kMaxDyn / (per_id * wp));
    if (warps < 1) return fail("key span too large for shared memory");
    p->block = 32 * warps;
    p->unroll = (int)opt("group.wp_unroll", 2);
    p->vec = (int)opt("group.wp_vec", 8);
    wp_ilp = (int)opt("group.wp_ilp", 1);
    if (wp_ilp != 1 && wp_ilp != 2 && wp_ilp != 4) return fail("group.wp_ilp must be 1, 2 or 4");
    p->smem_bytes = (size_t)(per_id * wp * warps);
    p->entry = "wdb_group_wp";
  } else {
    p->entry = "wdb_group";
  }
  p->wp_ids = (int)wp;
  p->smem_slots = (int)slots;
  int log2 = 0;
  while ((1ll << log2) < slots) ++log2;
  if (wp == 0) {
    size_t per_slot = 8 + 4 + ((needs & WDB_NEED_CNT_BIT) ? 4 : 0) + ((needs & WDB_NEED_MINMAX_BIT) ? 16 : 0) + ((needs & WDB_NEED_FIRST_BIT) ? 8 : 0);
    p->smem_bytes = per_slot * (size_t)slots;
  }
  const bool aligned = !check_alignment || all_aligned(spec.used, cols, nullptr, (size_t)p->vec * 4);
  // the side table lives in the L2: the column stream is loaded evict-first so that it does not displace it
  const int64_t ld_hint = dense && p->vec == 8 ? opt("group.dense_ld_hint", 2) : opt("group.ld_hint", 0);
  spec.defines = {{"WDB_VEC", p->vec}, {"WDB_ALIGNED", aligned ? 1 : 0}, {"WDB_LD_HINT", ld_hint}, {"WDB_ST_HINT", 0}, {"WDB_DENSE", dense ? 1 : 0},
                  {"WDB_BLOCK", p->block}, {"WDB_UNROLL", p->unroll}, {"WDB_NEEDS", needs}, {"WDB_SMEM_SLOTS", slots},
                  {"WDB_SMEM_LOG2", log2}, {"WDB_SMEM_PROBES", opt("group.smem_probes", 4)}, {"WDB_HAS_COND", has_cond ? 1 : 0},
                  {"WDB_WP_IDS", wp}, {"WDB_WP_ILP", wp_ilp}, {"WDB_WP_MODE", wp > 0 ? wp_mode : 0}};
  spec.fns.push_back({"val", "float", (needs & ~WDB_NEED_CNT_BIT & ~WDB_NEED_FIRST_BIT) ? val : "0.0f"});  // COUNT never evaluates its argument (src/warpdb.cpp:376)
  spec.fns.push_back({"key", "int", key});
  if (has_cond) spec.fns.push_back({"cond", "bool", cond});
  spec.bodies = {k_src_group_table, k_src_group};
  return 0;
}

int gen_group_source(const wdb_col_t *cols, int ncols, const char *val, const char *key, const char *cond, int agg,
                     std::string *src) {
  GroupPlan p;
  const int64_t span = opt("group.debug_span", 0);   // introspection only: pretend the key range [0, span) is known
  if (plan_group(cols, ncols, val, key, cond, needs_for_agg(agg), 2048, KeyRange{span > 0, 0, span - 1}, span > 0,
                 wp_fits(needs_for_agg(agg), span), false, &p)) return 1;
  *src = gen_source(p.spec);
  return 0;
}

// Optimizer statistics gathered on demand: min/max of the key EXPRESSION (as the kernels evaluate
// it) in one streaming pass over the columns it reads (kernels/keyrange.cuh; ~0.65 ms per 1e9 rows of a
// 4-byte column) before the aggregation when the caller supplied no range.  Not cached: a (pointer,
// length) pair says nothing about the contents (allocators reuse addresses), and a stale range sends
// every out-of-range row down the slow global path.
constexpr int kKeyRangeBlock = 512, kKeyRangeUnroll = 2, kKeyRangeVec = 8;
// returns false when there is nothing to scan (constant key, non-numeric or missing column)
static bool plan_keyrange(const wdb_col_t *cols, int ncols, const char *key_expr, bool check_alignment, GenSpec *spec) {
  spec->kind = "keyrange";
  spec->used = find_used_columns(cols, ncols, {key_expr});
  if (spec->used.empty()) return false;
  for (const auto &u : spec->used)
    if (dtype_size(u.dtype) == 0 || (check_alignment && !cols[u.table_index].dptr)) return false;
  const bool aligned = !check_alignment || all_aligned(spec->used, cols, nullptr, (size_t)kKeyRangeVec * 4);
  spec->defines = {{"WDB_VEC", kKeyRangeVec}, {"WDB_ALIGNED", aligned ? 1 : 0}, {"WDB_LD_HINT", 0}, {"WDB_ST_HINT", 0},
                   {"WDB_BLOCK", kKeyRangeBlock}, {"WDB_UNROLL", kKeyRangeUnroll}};
  spec->fns.push_back({"key", "int", key_expr});
  spec->bodies = {k_src_keyrange};
  return true;
}

int gen_keyrange_source(const wdb_col_t *cols, int ncols, const char *key, std::string *src) {
  GenSpec spec;
  if (!plan_keyrange(cols, ncols, key, false, &spec)) return fail("the key expression reads no numeric column");
  *src = gen_source(spec);
  return 0;
}

int auto_key_range(Device *d, cudaStream_t s, const wdb_col_t *cols, int ncols, const char *key_expr, int64_t n, KeyRange *out) {
  GenSpec spec;
  if (!plan_keyrange(cols, ncols, key_expr, true, &spec)) return 0;
  const int block = kKeyRangeBlock, unroll = kKeyRangeUnroll, vec = kKeyRangeVec;
  Kernel k;
  if (get_kernel(d, gen_source(spec), "wdb_keyrange.cu", "wdb_keyrange", &k)) return 1;
  Scratch scratch;
  WDB_CUDA(scratch.alloc(8, s));
  int *d_out = scratch.as<int>();
  const int init[2] = {INT32_MAX, INT32_MIN};
  WDB_CUDA(cudaMemcpyAsync(d_out, init, 8, cudaMemcpyHostToDevice, s));
  const int64_t tile_rows = (int64_t)block * unroll * vec;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + tile_rows - 1) / tile_rows, (int64_t)d->num_sms * 4));
  std::vector<const void *> ptrs;
  for (const auto &u : spec.used) ptrs.push_back(cols[u.table_index].dptr);
  long long nn = n;
  void *args[] = {ptrs.data(), &nn, &d_out};
  if (launch(k, grid, block, 0, s, args)) return 1;
  int h[2];
  WDB_CUDA(cudaMemcpyAsync(h, d_out, 8, cudaMemcpyDeviceToHost, s));
  WDB_CUDA(cudaStreamSynchronize(s));
  if (h[0] <= h[1]) *out = KeyRange{true, h[0], h[1]};
  return 0;
}

// ---- direct-addressed side table: host side -------------------------------------------------------
struct DenseScratch { Scratch mem; char *buf = nullptr; unsigned *tile_counts; unsigned long long *tile_offsets, *total; long long ntiles; };
static int dense_scratch(wdb_agg *t, cudaStream_t s, DenseScratch *sc) {
  sc->ntiles = ((long long)t->T.dspan + kDenseTile - 1) / kDenseTile;
  const size_t a = ((size_t)sc->ntiles * 4 + 15) & ~(size_t)15;
  WDB_CUDA(sc->mem.alloc(a + (size_t)sc->ntiles * 8 + 16, s));
  sc->buf = sc->mem.as<char>();
  sc->tile_counts = (unsigned *)sc->buf;
  sc->tile_offsets = (unsigned long long *)(sc->buf + a);
  sc->total = sc->tile_offsets + sc->ntiles;
  return 0;
}
#define WDB_DENSE_DISPATCH(needs, CALL)                                    \
  switch ((needs) & 7) {                                                   \
  case 1: { constexpr int N = 1; CALL; } break;                            \
  case 2: { constexpr int N = 2; CALL; } break;                            \
  case 3: { constexpr int N = 3; CALL; } break;                            \
  case 4: { constexpr int N = 4; CALL; } break;                            \
  case 5: { constexpr int N = 5; CALL; } break;                            \
  case 6: { constexpr int N = 6; CALL; } break;                            \
  default: { constexpr int N = 7; CALL; } break;                           \
  }
// count + scan: tile offsets and the number of present entries (left on the device in sc->total and, when given, *d_total)
static int dense_rank(wdb_agg *t, cudaStream_t s, DenseScratch *sc, long long *d_total = nullptr) {
  if (dense_scratch(t, s, sc)) return 1;
  WDB_DENSE_DISPATCH(t->needs, (dense_count_kernel<N><<<(unsigned)sc->ntiles, kDenseBlock, 0, s>>>(t->T, sc->tile_counts)));
  dense_scan_kernel<<<1, 1024, 0, s>>>(sc->tile_counts, sc->tile_offsets, sc->ntiles, sc->total, d_total);
  stats().launches += 2;
  WDB_CUDA(cudaGetLastError());
  return 0;
}
// fold the side table into the hash table and retire it
static int dense_flush(wdb_agg *t, cudaStream_t s) {
  if (!t->dense_live) return 0;
  DenseScratch sc;
  if (dense_rank(t, s, &sc)) return 1;
  WDB_DENSE_DISPATCH(t->needs, (dense_emit_kernel<N><<<(unsigned)sc.ntiles, kDenseBlock, 0, s>>>(t->T, sc.tile_offsets, sc.total, 0, 0, 1, 0, nullptr, nullptr,
                                                                                                    nullptr, nullptr, nullptr, nullptr)));
  stats().launches++;
  WDB_CUDA(cudaGetLastError());
  t->dense_live = false;
  t->T.dspan = 0;
  return 0;
}
static int64_t dense_bytes_per_key(int needs) {
  return ((needs & WDB_NEED_SUM_BIT) ? 8 : 0) + ((needs & WDB_NEED_CNT_BIT) ? 8 : 0) + ((needs & WDB_NEED_MINMAX_BIT) ? 16 : 0);
}
// make [lo, lo + span) the live side table (flushing a different live range first)
int dense_prepare(wdb_agg *t, cudaStream_t s, int64_t lo, int64_t span) {
  if (t->needs & WDB_NEED_FIRST_BIT) return fail("internal: a table that tracks first rows has no direct-addressed side table");
  if (t->dense_live && t->T.dlo == (int)lo && (int64_t)t->T.dspan == span) return 0;
  if (t->dense_live && lo >= (int64_t)t->T.dlo && lo + span <= (int64_t)t->T.dlo + (int64_t)t->T.dspan) return 0;   // covered by the live range
  if (dense_flush(t, s)) return 1;
  if (t->dense_cap < span) {
    if (t->dense_mem) { WDB_CUDA(cudaStreamSynchronize(s)); WDB_CUDA(cudaFree(t->dense_mem)); t->dense_mem = nullptr; t->dense_cap = 0; }
    cudaError_t e = cudaMalloc((void **)&t->dense_mem, (size_t)span * dense_bytes_per_key(t->needs) + 64);
    if (e != cudaSuccess) return fail("CUDA error: %s (direct-addressed aggregation table of %lld entries)", cudaGetErrorString(e), (long long)span);
    t->dense_cap = span;
  }
  char *p = t->dense_mem;
  t->T.dsums = (double *)p; p += (t->needs & WDB_NEED_SUM_BIT) ? (size_t)t->dense_cap * 8 : 0;
  t->T.dcnts = (unsigned long long *)p; p += (t->needs & WDB_NEED_CNT_BIT) ? (size_t)t->dense_cap * 8 : 0;
  t->T.dmins = (long long *)p; p += (t->needs & WDB_NEED_MINMAX_BIT) ? (size_t)t->dense_cap * 8 : 0;
  t->T.dmaxs = (long long *)p;
  t->T.dlo = (int)lo;
  t->T.dspan = (unsigned)span;
  dense_init_kernel<<<grid_for(t->dev, span), 256, 0, s>>>(t->T, t->needs, span);
  stats().launches++;
  WDB_CUDA(cudaGetLastError());
  t->dense_live = true;
  return 0;
}

int agg_hash_to_dense(wdb_agg *t, cudaStream_t s) {
  if (!t->dense_live) return fail("internal: no live side table");
  const long long slots = t->cap + 1;
  WDB_DENSE_DISPATCH(t->needs, (agg_hash_to_dense_kernel<N><<<grid_for(t->dev, slots), 256, 0, s>>>(t->T, slots)));
  stats().launches++;
  WDB_CUDA(cudaGetLastError());
  return 0;
}

int agg_export_dense_async(wdb_agg *t, cudaStream_t s, int agg, int order, int32_t *d_keys, float *d_vals, double *d_sums,
                           int64_t *d_counts, double *d_mins, double *d_maxs, int64_t cap, long long *d_groups) {
  if (!t->dense_live) return fail("internal: no live side table");
  if (order == WDB_ORDER_FIRST) return fail("internal: the side table cannot export in first-appearance order");
  DenseScratch sc;
  if (dense_rank(t, s, &sc, d_groups)) return 1;
  WDB_DENSE_DISPATCH(t->needs, (dense_emit_kernel<N><<<(unsigned)sc.ntiles, kDenseBlock, 0, s>>>(t->T, sc.tile_offsets, sc.total, order == WDB_ORDER_KEY_DESC, agg, 0,
                                                                                                    (long long)cap, d_keys, d_vals, d_sums, (long long *)d_counts, d_mins, d_maxs)));
  stats().launches++;
  WDB_CUDA(cudaGetLastError());
  return 0;
}

int agg_create_on(int device, int64_t expected_groups, int needs, cudaStream_t stream, wdb_agg **out) {
  Device *d;
  if (get_device(device, &d)) return 1;
  if (!out) return fail("null output");
  if (needs <= 0 || needs > 15) return fail("invalid needs mask %d", needs);
  // small tables: load factor <= 0.5; large ones (beyond the L2) trade probe length for footprint
  int64_t want = std::max<int64_t>(expected_groups, 512) * 2, cap = 1024;
  if (expected_groups > (1 << 20)) want = expected_groups + expected_groups / 2;
  while (cap < want) cap <<= 1;
  if (cap > (1ll << 31)) return fail("aggregation table of %lld slots is too large", (long long)cap);
  wdb_agg *t = new wdb_agg();
  t->dev = d;
  t->needs = needs;
  t->cap = cap;
  const size_t slots = (size_t)cap + 1;
  const size_t bytes = slots * (4 + 8 + 8 + 8 + 8 + 8) + 256 + 64;
  cudaError_t e = cudaMalloc((void **)&t->mem, bytes);
  if (e != cudaSuccess) { delete t; return fail("CUDA error: %s (aggregation table of %zu bytes)", cudaGetErrorString(e), bytes); }
  char *p = t->mem;
  t->T.sums = (double *)p; p += 8 * slots;
  t->T.counts = (unsigned long long *)p; p += 8 * slots;
  t->T.mins = (long long *)p; p += 8 * slots;
  t->T.maxs = (long long *)p; p += 8 * slots;
  t->T.first = (long long *)p; p += 8 * slots;
  t->T.keys = (int *)p; p += 4 * slots;
  p = (char *)(((uintptr_t)p + 63) & ~(uintptr_t)63);
  t->T.meta = (unsigned *)p;
  t->T.mask = (unsigned)(cap - 1);
  {
    unsigned lg = 0;
    while ((1ll << lg) < cap) ++lg;
    t->T.shift = 32u - lg;
  }
  if (wdb_agg_reset(t, stream)) { cudaFree(t->mem); delete t; return 1; }
  *out = t;
  return 0;
}

}  // namespace wdb

extern "C" {

int wdb_agg_create(int device, int64_t expected_groups, int needs, wdb_agg_t **out) {
  // The table is initialised on the legacy NULL stream; callers consume on streams of their own
  // (non-blocking streams have no implicit ordering with it), so wait for the initialisation here.
  if (wdb::agg_create_on(device, expected_groups, needs, nullptr, out)) return 1;
  WDB_CUDA(cudaStreamSynchronize(nullptr));
  return 0;
}

int wdb_agg_destroy(wdb_agg_t *t) {
  if (!t) return 0;
  cudaSetDevice(t->dev->id);
  cudaFree(t->mem);
  if (t->dense_mem) cudaFree(t->dense_mem);
  delete t;
  return 0;
}

int wdb_agg_set_key_range(wdb_agg_t *t, int known, int64_t lo, int64_t hi) {
  if (!t) return fail("null table");
  if (known && (lo > hi || lo < INT32_MIN || hi > INT32_MAX)) return fail("invalid key range [%lld, %lld]", (long long)lo, (long long)hi);
  t->have_range = known != 0;
  t->key_lo = lo;
  t->key_hi = hi;
  return 0;
}

int wdb_agg_reset(wdb_agg_t *t, void *stream) {
  if (!t) return fail("null table");
  WDB_CUDA(cudaSetDevice(t->dev->id));
  t->dense_live = false;   // the side table is re-initialised by the next consume that uses it
  t->T.dspan = 0;
  agg_init_kernel<<<grid_for(t->dev, t->cap + 1), 256, 0, (cudaStream_t)stream>>>(t->T, t->cap + 1);
  stats().launches++;
  WDB_CUDA(cudaGetLastError());
  return 0;
}

int wdb_agg_consume(wdb_agg_t *t, void *stream, const wdb_col_t *cols, int ncols, const char *val_expr, const char *key_expr,
                    const char *cond, int64_t n, int64_t row_base) {
  if (!t) return fail("null table");
  if (!key_expr || !*key_expr) return fail("empty GROUP BY key expression");
  if (!val_expr || !*val_expr) val_expr = "0.0f";
  if (n < 0) return fail("negative row count");
  Device *d = t->dev;
  WDB_CUDA(cudaSetDevice(d->id));
  KeyRange range{t->have_range, t->key_lo, t->key_hi};
  if (!range.known && n >= opt("group.auto_stats_min_rows", 1 << 20) &&
      auto_key_range(d, (cudaStream_t)stream, cols, ncols, key_expr, n, &range))
    return 1;
  // Integer keys with a known range: <= wp_max_span -> warp-private shared-memory accumulators;
  // <= dense_max_span -> direct-addressed table in HBM/L2 (no probe, no CAS, ordered export without
  // a sort); otherwise the hash table.
  const bool sumcnt = (t->needs & WDB_NEED_FIRST_BIT) == 0;   // what the side table can hold: sums, counts, extrema (not first rows)
  const int64_t span = range.known ? range.hi - range.lo + 1 : -1;
  const bool want_wp = wp_fits(t->needs, span);
  // (between the two, contention on few L2 addresses makes the hash table with its bigger footprint the faster one: measured 120 vs 88 Grows/s at 10 K keys)
  // ... and only where the table is not absurdly larger than the rows that will land in it (16 B per key of the range)
  const bool want_dense = sumcnt && !want_wp && span >= opt("group.dense_min_span", 32768) && span <= opt("group.dense_max_span", 1 << 26) &&
                          (span <= (1 << 20) || span <= 4 * n || (t->dense_live && (int64_t)t->T.dspan >= span));
  // the warp-private kernel folds its per-CTA totals into a (tiny) direct-addressed table as well:
  // the result is then in key order without a sort.  A small
  // side table is a bad target for row-by-row atomics though (few L2 lines take them all), so any other
  // kernel folds it into the hash table first.
  const bool wp_side = want_wp && sumcnt;
  if (t->dense_live && !wp_side && !want_dense && (int64_t)t->T.dspan < opt("group.dense_min_span", 32768) && wdb::dense_flush(t, (cudaStream_t)stream)) return 1;
  if ((want_dense || wp_side) && n > 0 && wdb::dense_prepare(t, (cudaStream_t)stream, range.lo, span)) return 1;
  GroupPlan p;
  if (plan_group(cols, ncols, val_expr, key_expr, cond, t->needs, t->cap, range, t->dense_live, want_wp && (t->dense_live || !sumcnt), true, &p)) return 1;
  Kernel k;
  if (get_kernel(d, gen_source(p.spec), "wdb_group.cu", p.entry, &k)) return 1;
  if (n == 0) return 0;
  if (p.smem_bytes > 48 * 1024) WDB_CUDA(cudaFuncSetAttribute((const void *)k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
  int nb = 0;
  WDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void *)k.fn, p.block, p.smem_bytes));
  nb = std::max(1, std::min<int>(nb, (int)opt("group.ctas_per_sm", (t->dense_live && p.wp_ids == 0) ? 2 : 8)));
  const int64_t tile_rows = (int64_t)p.block * p.unroll * p.vec;
  const int64_t ntiles = (n + tile_rows - 1) / tile_rows;
  // several waves of short CTAs instead of one persistent wave while a cross-GPU exchange of the previous table slice may
  // be waiting for CTA slots on a side stream (ops_comm.cu): a persistent grid would keep every SM until the kernel ends
  const int64_t waves = t->after_slice ? std::max<int64_t>(1, opt("group.dense_waves", 8)) : 1;
  unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)d->num_sms * nb * waves));
  std::vector<const void *> ptrs;
  for (const auto &u : p.spec.used) ptrs.push_back(cols[u.table_index].dptr);
  if (ptrs.empty()) ptrs.push_back(nullptr);
  long long nn = n, rb = row_base;
  // Tables larger than the L2 are filled in 2^pass_bits launches; launch p folds only the keys whose
  // hash prefix is p, which occupy one contiguous 1/2^pass_bits of the table, so the random
  // read-modify-writes of a launch hit the L2 instead of DRAM.  Each launch re-reads the columns.
  unsigned pass_bits = 0;
  if (p.wp_ids == 0 && p.smem_slots == 0 && !t->dense_live) {
    const double touched = (double)t->cap * (4.0 + ((t->needs & WDB_NEED_SUM_BIT) ? 8 : 0) + ((t->needs & WDB_NEED_CNT_BIT) ? 8 : 0) +
                                             ((t->needs & WDB_NEED_MINMAX_BIT) ? 16 : 0) + ((t->needs & WDB_NEED_FIRST_BIT) ? 8 : 0));
    const double budget = (double)opt("group.l2_budget_mb", 48) * 1048576.0;
    while (pass_bits < 6 && touched / (double)(1u << pass_bits) > budget) ++pass_bits;
    const int64_t forced = opt("group.pass_bits", -1);
    if (forced >= 0) pass_bits = (unsigned)forced;
  }
  if (p.wp_ids > 0) {
    int key_base = (int)range.lo;
    void *args[] = {ptrs.data(), &nn, &rb, &t->T, &key_base};
    return launch(k, grid, p.block, p.smem_bytes, (cudaStream_t)stream, args);
  }
  if (t->dense_live) {   // index slices of the direct-addressed table, each small enough to stay in the L2
    const double per = (double)dense_bytes_per_key(t->needs);
    const double budget = (double)opt("group.dense_l2_budget_mb", 60) * 1048576.0;
    int64_t passes = std::max<int64_t>(1, (int64_t)std::ceil((double)t->T.dspan * per / budget));
    const int64_t forced = opt("group.dense_passes", -1);
    if (forced > 0) passes = forced;
    const int64_t slice = ((int64_t)t->T.dspan + passes - 1) / passes;
    for (int64_t i = 0; i < passes; ++i) {
      unsigned lo = (unsigned)(i * slice), hi = (unsigned)std::min<int64_t>((i + 1) * slice, (int64_t)t->T.dspan);
      void *args[] = {ptrs.data(), &nn, &rb, &t->T, &lo, &hi};
      if (launch(k, grid, p.block, p.smem_bytes, (cudaStream_t)stream, args)) return 1;
      if (t->after_slice && t->after_slice(lo, hi)) return 1;
    }
    return 0;
  }
  for (unsigned pass = 0; pass < (1u << pass_bits); ++pass) {
    void *args[] = {ptrs.data(), &nn, &rb, &t->T, &pass_bits, &pass};
    if (launch(k, grid, p.block, p.smem_bytes, (cudaStream_t)stream, args)) return 1;
  }
  return 0;
}

int wdb_agg_merge(wdb_agg_t *t, void *stream, const int32_t *d_keys, const double *d_sums, const int64_t *d_counts,
                  const double *d_mins, const double *d_maxs, const int64_t *d_first, int64_t m) {
  if (!t) return fail("null table");
  if (m <= 0) return 0;
  const int needs = t->needs;
  if (!d_keys) return fail("merge needs the partial keys");
  if ((needs & WDB_NEED_SUM_BIT) && !d_sums) return fail("merge needs the partial sums");
  if ((needs & WDB_NEED_CNT_BIT) && !d_counts) return fail("merge needs the partial counts");
  if ((needs & WDB_NEED_MINMAX_BIT) && (!d_mins || !d_maxs)) return fail("merge needs the partial minima and maxima");
  if ((needs & WDB_NEED_FIRST_BIT) && !d_first) return fail("merge needs the partial first rows");
  WDB_CUDA(cudaSetDevice(t->dev->id));
  const unsigned g = grid_for(t->dev, m);
  cudaStream_t s = (cudaStream_t)stream;
  // partial keys inside the side table's range are added there (one home per key); statistics set on
  // this table (wdb_agg_set_key_range) open a side table for the partials as they would for rows
  if (!t->dense_live && t->have_range && (needs & WDB_NEED_FIRST_BIT) == 0) {
    const int64_t span = t->key_hi - t->key_lo + 1;
    if (span >= 1 && span <= opt("group.dense_max_span", 1 << 26) && (span <= (1 << 16) || span <= 16 * m) && wdb::dense_prepare(t, s, t->key_lo, span)) return 1;
  }
#define WDB_MERGE_CASE(N) case N: agg_merge_kernel<N><<<g, 256, 0, s>>>(t->T, d_keys, d_sums, (const long long *)d_counts, d_mins, d_maxs, (const long long *)d_first, m); break;
  switch (needs) {
    WDB_MERGE_CASE(1) WDB_MERGE_CASE(2) WDB_MERGE_CASE(3) WDB_MERGE_CASE(4) WDB_MERGE_CASE(5) WDB_MERGE_CASE(6) WDB_MERGE_CASE(7)
    WDB_MERGE_CASE(8) WDB_MERGE_CASE(9) WDB_MERGE_CASE(10) WDB_MERGE_CASE(11) WDB_MERGE_CASE(12) WDB_MERGE_CASE(13) WDB_MERGE_CASE(14) WDB_MERGE_CASE(15)
  }
#undef WDB_MERGE_CASE
  stats().launches++;
  WDB_CUDA(cudaGetLastError());
  return 0;
}

static int read_meta(wdb_agg_t *t, cudaStream_t s, unsigned meta[4]) {
  WDB_CUDA(cudaMemcpyAsync(meta, t->T.meta, 16, cudaMemcpyDeviceToHost, s));
  WDB_CUDA(cudaStreamSynchronize(s));
  if (meta[1]) return fail("aggregation table overflow: more than %lld distinct keys; recreate it with a larger expected_groups", (long long)(t->cap / 2));
  return 0;
}

int wdb_agg_size(wdb_agg_t *t, void *stream, int64_t *h_groups) {
  if (!t) return fail("null table");
  WDB_CUDA(cudaSetDevice(t->dev->id));
  unsigned meta[4];
  if (read_meta(t, (cudaStream_t)stream, meta)) return 1;
  // a key may sit in both tables (rows that arrived before the side table covered it): fold first, then count
  if (t->dense_live && (meta[0] || meta[2]) && (wdb::dense_flush(t, (cudaStream_t)stream) || read_meta(t, (cudaStream_t)stream, meta))) return 1;
  *h_groups = (int64_t)meta[0] + (meta[2] ? 1 : 0);
  if (t->dense_live) {
    cudaStream_t s = (cudaStream_t)stream;
    wdb::DenseScratch sc;
    if (wdb::dense_rank(t, s, &sc)) return 1;
    unsigned long long total = 0;
    WDB_CUDA(cudaMemcpyAsync(&total, sc.total, 8, cudaMemcpyDeviceToHost, s));
      WDB_CUDA(cudaStreamSynchronize(s));
    *h_groups += (int64_t)total;
  }
  return 0;
}

int wdb_agg_spilled(wdb_agg_t *t, void *stream, int64_t *h_rows) {
  if (!t || !h_rows) return fail("null argument");
  WDB_CUDA(cudaSetDevice(t->dev->id));
  unsigned meta[4];
  WDB_CUDA(cudaMemcpyAsync(meta, t->T.meta, 16, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  WDB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  *h_rows = meta[3];
  return 0;
}

int wdb_agg_export(wdb_agg_t *t, void *stream, int agg, int order, int32_t *d_keys, float *d_vals, double *d_sums,
                   int64_t *d_counts, double *d_mins, double *d_maxs, int64_t *d_first, int64_t cap, int64_t *h_groups) {
  if (!t) return fail("null table");
  if (agg < WDB_SUM || agg > WDB_MAX) return fail("invalid aggregation %d", agg);
  if (order < 0 || order > 2) return fail("invalid order %d", order);
  if (d_vals && (needs_for_agg(agg) & ~t->needs)) return fail("the table does not track what aggregation %d needs", agg);
  if (order == WDB_ORDER_FIRST && !(t->needs & WDB_NEED_FIRST_BIT)) return fail("first-appearance order needs a table created with the first-row bit");
  Device *d = t->dev;
  WDB_CUDA(cudaSetDevice(d->id));
  cudaStream_t s = (cudaStream_t)stream;
  unsigned meta[4];
  if (read_meta(t, s, meta)) return 1;
  if (t->dense_live) {
    const bool hash_empty = meta[0] == 0 && meta[2] == 0;
    if (hash_empty && order != WDB_ORDER_FIRST && !d_first) {
      // the whole result sits in the direct-addressed table, already in key order: rank + emit, no sort
      wdb::DenseScratch sc;
      if (wdb::dense_rank(t, s, &sc)) return 1;
      unsigned long long total = 0;
      WDB_CUDA(cudaMemcpyAsync(&total, sc.total, 8, cudaMemcpyDeviceToHost, s));
      WDB_CUDA(cudaStreamSynchronize(s));
      if (h_groups) *h_groups = (int64_t)total;
      if ((long long)total > cap) return fail("%lld groups exceed the output capacity %lld", (long long)total, (long long)cap);
      if (total) {
        WDB_DENSE_DISPATCH(t->needs, (dense_emit_kernel<N><<<(unsigned)sc.ntiles, kDenseBlock, 0, s>>>(t->T, sc.tile_offsets, sc.total, order == WDB_ORDER_KEY_DESC, agg, 0,
                                                                                                      (long long)cap, d_keys, d_vals, d_sums, (long long *)d_counts, d_mins, d_maxs)));
        stats().launches++;
        WDB_CUDA(cudaGetLastError());
      }
          WDB_CUDA(cudaStreamSynchronize(s));
      return 0;
    }
    if (wdb::dense_flush(t, s) || read_meta(t, s, meta)) return 1;
  }
  const long long g = (long long)meta[0] + (meta[2] ? 1 : 0);
  if (h_groups) *h_groups = g;
  if (g == 0) return 0;
  if (g > cap) return fail("%lld groups exceed the output capacity %lld", g, (long long)cap);
  const long long slots = t->cap + 1;
  const size_t kb = 8 * (size_t)g, pb = 4 * (size_t)g;
  Scratch scratch;
  WDB_CUDA(scratch.alloc(2 * kb + 2 * pb + 64, s));
  char *buf = scratch.as<char>();
  unsigned long long *sk = (unsigned long long *)buf, *skt = (unsigned long long *)(buf + kb);
  unsigned *sp = (unsigned *)(buf + 2 * kb), *spt = (unsigned *)(buf + 2 * kb + pb);
  unsigned long long *counter = (unsigned long long *)(buf + 2 * kb + 2 * pb);
  WDB_CUDA(cudaMemsetAsync(counter, 0, 8, s));
  agg_collect_kernel<<<grid_for(d, slots), 256, 0, s>>>(t->T, slots, order, sk, sp, counter);
  // collection order is arbitrary (atomic counter); sorting on the full key makes the export
  // deterministic: group keys are unique, first-appearance rows are unique
  if (radix_sort<unsigned long long>(d, s, sk, skt, sp, spt, g, order == WDB_ORDER_FIRST ? 64 : 32)) return 1;
  agg_emit_kernel<<<grid_for(d, g), 256, 0, s>>>(t->T, slots, sp, g, agg, d_keys, d_vals, d_sums, (long long *)d_counts, d_mins, d_maxs,
                                                  (long long *)d_first);
  stats().launches += 2;
  WDB_CUDA(cudaGetLastError());
  WDB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int wdb_group_agg(int device, void *stream, const wdb_col_t *cols, int ncols, const char *val_expr, const char *key_expr,
                  const char *cond, int agg, int order, int64_t n, int64_t expected_groups, int32_t *d_keys, float *d_vals,
                  int64_t cap, int64_t *h_groups) {
  if (agg < WDB_SUM || agg > WDB_MAX) return fail("invalid aggregation %d", agg);
  int needs = needs_for_agg(agg) | (order == WDB_ORDER_FIRST ? WDB_NEED_FIRST_BIT : 0);
  int64_t expect = expected_groups > 0 ? expected_groups : 1 << 16;
  for (int attempt = 0; attempt < 6; ++attempt) {
    wdb_agg_t *t = nullptr;
    if (wdb::agg_create_on(device, expect, needs, (cudaStream_t)stream, &t)) return 1;
    int rc = wdb_agg_consume(t, stream, cols, ncols, val_expr, key_expr, cond, n, 0);
    int64_t g = 0;
    if (!rc) rc = wdb_agg_export(t, stream, agg, order, d_keys, d_vals, nullptr, nullptr, nullptr, nullptr, nullptr, cap, &g);
    const bool overflow = rc && strstr(wdb_last_error(), "table overflow") != nullptr;
    wdb_agg_destroy(t);
    if (!rc) { if (h_groups) *h_groups = g; return 0; }
    if (!overflow) return 1;
    expect *= 16;  // unknown cardinality: grow and run again
    if (expect > (1ll << 30)) break;
  }
  return fail("aggregation table overflow: too many distinct keys");
}
}
