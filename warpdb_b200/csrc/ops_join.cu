// ops_join.cu -- inner equi-join on integer keys (the last one).
//
// Reference: `JOIN <table> ON <expr>` is parsed into QueryAST::joins (src/expression.cpp:375-401,
// include/expression.hpp:123-135) and its condition is validated (src/warpdb.cpp:321-323); nothing
// in the reference executes it ("Currently JOIN loads the same table for demonstration purposes",
// include/warpdb.hpp:22).  The semantics here are SQL's: every pair (i, j) with
// probe_key[i] == build_key[j], in nested-loop order -- ascending i, and ascending j within one i --
// which is what the oracle's orc_join_pairs enumerates.
//
// Sort-based, so the result order needs no post-pass and equal keys on either side are handled:
//   build   (key, row) pairs of the build column, stable LSD radix sort by key (ops_sort.cu); equal
//           keys keep their row order.  When the key range is dense (ids of a dimension table) every key
//           value of [min, max] gets an 8-byte entry {matches, build row or first sorted position} as
//           well: a probe is then one load instead of a binary search plus the row lookup
//   probe   two streaming passes over the probe column.  Pass 1 counts the matches of every 1 024-row
//           tile (lower / upper bound in the sorted keys, which stay in L2 for dimension-sized build
//           sides), one block scan turns the tile counts into offsets, pass 2 repeats the searches
//           and writes the pairs at tile offset + in-tile prefix.  Nothing per probe row is kept
//           between the passes: HBM traffic is 2 x the probe keys + 16 B per emitted pair.
//   gather  dst[i] = src[rows[i]]: late materialisation of the columns a query reads, after which
//           every other operator of the core (filter / project / GROUP BY / top-k) runs unchanged
//           on the joined columns.
#include <algorithm>
#include <mutex>

#include "core.hpp"

namespace wdb {

template <class K>
int radix_sort(Device *d, cudaStream_t s, K *keys, K *tmp_keys, unsigned *pay, unsigned *tmp_pay, long long n, int key_bits);

constexpr int kJoinBlock = 256;
constexpr int kJoinWarps = kJoinBlock / 32;
constexpr int kJoinRounds = 4;
constexpr int kJoinTile = kJoinBlock * kJoinRounds;

// order-preserving signed -> unsigned, and the payload (= build row) alongside
template <class S, class U>
__global__ void join_encode_kernel(const S *__restrict__ in, U *__restrict__ out, unsigned *__restrict__ rows, long long n) {
  constexpr U kSign = (U)1 << (8 * sizeof(U) - 1);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    out[i] = (U)in[i] ^ kSign;
    rows[i] = (unsigned)i;
  }
}
template <class S, class U>
__global__ void join_decode_kernel(U *keys, long long n) {   // in place: same width
  constexpr U kSign = (U)1 << (8 * sizeof(U) - 1);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const S v = (S)(keys[i] ^ kSign);
    reinterpret_cast<S *>(keys)[i] = v;
  }
}

// [lb, ub) = positions of `v` in the ascending keys[0..m)
template <class B>
__device__ __forceinline__ void equal_range(const B *__restrict__ keys, unsigned m, long long v, unsigned *lb, unsigned *ub) {
  unsigned lo = 0, hi = m;
  while (lo < hi) {
    const unsigned mid = lo + ((hi - lo) >> 1);
    if ((long long)keys[mid] < v) lo = mid + 1; else hi = mid;
  }
  *lb = lo;
  unsigned e = lo;
  if (lo < m && (long long)keys[lo] == v) {
    e = lo + 1;                                     // unique build keys (the usual case) stop here
    if (e < m && (long long)keys[e] == v) {
      unsigned l2 = e + 1, h2 = m;
      while (l2 < h2) {
        const unsigned mid = l2 + ((h2 - l2) >> 1);
        if ((long long)keys[mid] <= v) l2 = mid + 1; else h2 = mid;
      }
      e = l2;
    }
  }
  *ub = e;
}

// The probe's view of the index.  `tab` (optional) is the direct-addressed form of the same sorted
// keys, one 8-byte entry per key value of [lo, lo + span): {matches, the build row itself when there
// is exactly one match (distinct ids: the usual case), else the position of the first match in the
// sorted order}.  A probe is then ONE load instead of ~log2(m) dependent ones plus the row lookup.
template <class B> struct JoinView {
  const B *keys;
  const uint2 *tab;
  long long lo, span;
  unsigned m;
};
// number of matches of v; *ref = the build row (*is_row) or the sorted position of the first match
template <class B>
__device__ __forceinline__ unsigned join_lookup(const JoinView<B> &ix, long long v, unsigned *ref, bool *is_row) {
  if (ix.tab) {
    const unsigned long long k = (unsigned long long)v - (unsigned long long)ix.lo;   // v < lo wraps to a huge value
    if (k < (unsigned long long)ix.span) {
      const uint2 e = ix.tab[k];
      *ref = e.y;
      *is_row = e.x == 1u;
      return e.x;
    }
    *ref = 0;
    *is_row = false;
    return 0;
  }
  unsigned lb, ub;
  equal_range<B>(ix.keys, ix.m, v, &lb, &ub);
  *ref = lb;
  *is_row = false;
  return ub - lb;
}
// first[k] for k in [0, span]: one binary search per key value of the range, once per index
template <class B>
__global__ void join_first_kernel(const B *__restrict__ keys, unsigned m, long long lo, long long span, unsigned *__restrict__ first) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k <= span; k += (long long)gridDim.x * blockDim.x) {
    const long long v = lo + k;
    unsigned a = 0, b = m;
    while (a < b) {
      const unsigned mid = a + ((b - a) >> 1);
      if ((long long)keys[mid] < v) a = mid + 1; else b = mid;
    }
    first[k] = a;
  }
}

__global__ void join_tab_kernel(const unsigned *__restrict__ first, const unsigned *__restrict__ rows, long long span, uint2 *__restrict__ tab) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < span; k += (long long)gridDim.x * blockDim.x) {
    const unsigned lb = first[k], cnt = first[k + 1] - lb;
    tab[k] = make_uint2(cnt, cnt == 1u ? rows[lb] : lb);
  }
}

template <class B, class P>
__global__ void __launch_bounds__(kJoinBlock) join_count_kernel(const JoinView<B> ix, const P *__restrict__ pkeys, long long n,
                                                                unsigned long long *__restrict__ tile_counts) {
  __shared__ unsigned long long s_w[kJoinWarps];
  const long long base = (long long)blockIdx.x * kJoinTile;
  unsigned long long c = 0;
#pragma unroll
  for (int k = 0; k < kJoinRounds; ++k) {
    const long long i = base + k * kJoinBlock + threadIdx.x;
    if (i < n) {
      unsigned ref;
      bool is_row;
      c += join_lookup<B>(ix, (long long)pkeys[i], &ref, &is_row);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < kJoinWarps; ++w) t += s_w[w];
    tile_counts[blockIdx.x] = t;
  }
}

// exclusive scan of the tile counts, one block; *total = number of pairs
__global__ void __launch_bounds__(1024) join_scan_kernel(const unsigned long long *__restrict__ counts, unsigned long long *__restrict__ offsets,
                                                         long long ntiles, unsigned long long *__restrict__ total) {
  __shared__ unsigned long long s_part[1024];
  const long long per = (ntiles + 1023) / 1024;
  const long long b = min((long long)threadIdx.x * per, ntiles), e = min(b + per, ntiles);
  unsigned long long sum = 0;
  for (long long i = b; i < e; ++i) sum += counts[i];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < 1024; ++i) { const unsigned long long t = s_part[i]; s_part[i] = run; run += t; }
    *total = run;
  }
  __syncthreads();
  unsigned long long run = s_part[threadIdx.x];
  for (long long i = b; i < e; ++i) { offsets[i] = run; run += counts[i]; }
}

template <class B, class P>
__global__ void __launch_bounds__(kJoinBlock) join_emit_kernel(const JoinView<B> ix, const unsigned *__restrict__ brows,
                                                               const P *__restrict__ pkeys, long long n,
                                                               const unsigned long long *__restrict__ tile_offsets,
                                                               long long *__restrict__ out_probe, long long *__restrict__ out_build,
                                                               unsigned long long cap) {
  __shared__ unsigned long long s_w[kJoinWarps];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const long long base = (long long)blockIdx.x * kJoinTile;
  unsigned long long run = tile_offsets[blockIdx.x];
  // all lookups of the tile first (independent loads in flight), then one block scan per round
  unsigned refs[kJoinRounds], cs[kJoinRounds];
  bool is_rows[kJoinRounds];
#pragma unroll
  for (int k = 0; k < kJoinRounds; ++k) {
    const long long i = base + k * kJoinBlock + threadIdx.x;
    refs[k] = 0;
    cs[k] = 0;
    is_rows[k] = false;
    if (i < n) cs[k] = join_lookup<B>(ix, (long long)pkeys[i], &refs[k], &is_rows[k]);
  }
#pragma unroll
  for (int k = 0; k < kJoinRounds; ++k) {            // rows of one round are consecutive over the threads: output order = row order
    const long long i = base + k * kJoinBlock + threadIdx.x;
    const unsigned ref = refs[k], c = cs[k];
    const bool is_row = is_rows[k];
    const unsigned long long cnt = c;
    unsigned long long incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (unsigned)o) incl += y;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    unsigned long long wpre = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kJoinWarps; ++w) {
      const unsigned long long t = s_w[w];
      if ((unsigned)w < warp) wpre += t;
      total += t;
    }
    unsigned long long pos = run + wpre + incl - cnt;
    if (is_row) {                                      // cap: never outside the caller's arrays, whatever the offsets say
      if (pos < cap) {
        if (out_probe) out_probe[pos] = i;
        if (out_build) out_build[pos] = (long long)ref;
      }
    } else {
      for (unsigned q = 0; q < c && pos < cap; ++q, ++pos) {
        if (out_probe) out_probe[pos] = i;
        if (out_build) out_build[pos] = (long long)brows[ref + q];
      }
    }
    run += total;
    __syncthreads();
  }
}

template <class T>
__global__ void gather_rows_kernel(const T *__restrict__ src, const long long *__restrict__ rows, T *__restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[rows[i]];
}

static unsigned join_grid(Device *d, long long n) {
  return (unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)d->num_sms * 16));
}

}  // namespace wdb

using namespace wdb;

struct wdb_join {
  Device *dev = nullptr;
  int key_dtype = WDB_INT32;
  int64_t m = 0;
  void *keys = nullptr;        // ascending signed keys (int or long long)
  unsigned *rows = nullptr;    // build row of every sorted position
  uint2 *tab = nullptr;        // direct-addressed {matches, row or first position} over [lo, lo + span), when the key range is dense enough
  long long lo = 0, span = 0;
  // pass 1 of the last count-only probe, kept for the emitting call that normally follows it
  std::mutex mu;
  struct { void *buf = nullptr; const void *probe = nullptr; long long n = 0; int dtype = 0; bool direct = false; unsigned long long pairs = 0; } counted;
};

template <class B, class P>
static int join_probe_typed(wdb_join *j, cudaStream_t s, const void *probe_keys, long long n, long long ntiles, int64_t *d_probe_rows,
                            int64_t *d_build_rows, int64_t cap, int64_t *h_pairs) {
  const P *pkeys = static_cast<const P *>(probe_keys);
  const bool direct = opt("join.direct", 1) != 0 && j->tab != nullptr;
  const JoinView<B> ix{static_cast<const B *>(j->keys), direct ? j->tab : nullptr, j->lo, j->span, (unsigned)j->m};
  const int pdtype = sizeof(P) == 4 ? WDB_INT32 : WDB_INT64;
  const bool count_only = !d_probe_rows && !d_build_rows;
  Scratch scratch;
  unsigned long long pairs = 0;
  bool counted = false;
  {  // a count of exactly this probe column, left by the previous call: its tile offsets are pass 1
    std::lock_guard<std::mutex> lock(j->mu);
    if (j->counted.buf) {
      if (!count_only && j->counted.probe == probe_keys && j->counted.n == n && j->counted.dtype == pdtype && j->counted.direct == direct) {
        scratch.p = j->counted.buf;
        scratch.s = s;
        pairs = j->counted.pairs;
        counted = true;
      } else {
        cudaFreeAsync(j->counted.buf, s);
      }
      j->counted.buf = nullptr;
    }
  }
  if (!counted) WDB_CUDA(scratch.alloc(8 * (2 * (size_t)ntiles + 1), s));
  unsigned long long *counts = scratch.as<unsigned long long>(), *offsets = counts + ntiles, *total = offsets + ntiles;
  if (!counted) {
    join_count_kernel<B, P><<<(unsigned)ntiles, kJoinBlock, 0, s>>>(ix, pkeys, n, counts);
    join_scan_kernel<<<1, 1024, 0, s>>>(counts, offsets, ntiles, total);
    stats().launches += 2;
    WDB_CUDA(cudaGetLastError());
    WDB_CUDA(cudaMemcpyAsync(&pairs, total, 8, cudaMemcpyDeviceToHost, s));
    WDB_CUDA(cudaStreamSynchronize(s));
  }
  if (h_pairs) *h_pairs = (int64_t)pairs;
  if (count_only) {                                   // the caller sizes its arrays from *h_pairs and calls again
    std::lock_guard<std::mutex> lock(j->mu);
    j->counted.buf = scratch.p;
    scratch.p = nullptr;
    j->counted.probe = probe_keys;
    j->counted.n = n;
    j->counted.dtype = pdtype;
    j->counted.direct = direct;
    j->counted.pairs = pairs;
    return 0;
  }
  if ((long long)pairs > cap) return fail("%lld joined rows exceed the output capacity %lld", (long long)pairs, (long long)cap);
  if (pairs == 0) return 0;
  join_emit_kernel<B, P><<<(unsigned)ntiles, kJoinBlock, 0, s>>>(ix, j->rows, pkeys, n, offsets, (long long *)d_probe_rows, (long long *)d_build_rows,
                                                                 (unsigned long long)cap);
  stats().launches++;
  WDB_CUDA(cudaGetLastError());
  WDB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

extern "C" {

int wdb_join_build(int device, void *stream, const wdb_col_t *build_key, wdb_join_t **out) {
  Device *d;
  if (get_device(device, &d)) return 1;
  if (!build_key || !out) return fail("null argument");
  if (build_key->dtype != WDB_INT32 && build_key->dtype != WDB_INT64)
    return fail("JOIN needs integer key columns (%s is not one)", build_key->name ? build_key->name : "?");
  const long long m = build_key->len;
  if (m < 0) return fail("negative row count");
  if (m >= (1ll << 32)) return fail("the build side of a JOIN is limited to 2^32 - 1 rows (%lld given)", m);
  cudaStream_t s = (cudaStream_t)stream;
  wdb_join *j = new wdb_join;
  j->dev = d;
  j->key_dtype = build_key->dtype;
  j->m = m;
  *out = j;
  if (m == 0) return 0;
  const size_t ksz = build_key->dtype == WDB_INT32 ? 4 : 8;
  auto bail = [&](int rc) { wdb_join_destroy(j); *out = nullptr; return rc; };
  if (cudaMalloc(&j->keys, ksz * (size_t)m) != cudaSuccess || cudaMalloc((void **)&j->rows, 4 * (size_t)m) != cudaSuccess) {
    cudaGetLastError();
    return bail(fail("CUDA error: out of memory (join index of %lld rows)", m));
  }
  Scratch tmp;
  if (tmp.alloc((ksz + 4) * (size_t)m, s) != cudaSuccess) { cudaGetLastError(); return bail(fail("CUDA error: out of memory (join sort scratch)")); }
  const unsigned g = join_grid(d, m);
  int rc = 0;
  if (build_key->dtype == WDB_INT32) {
    unsigned *k = (unsigned *)j->keys, *kt = tmp.as<unsigned>(), *rt = kt + m;
    join_encode_kernel<int, unsigned><<<g, 256, 0, s>>>((const int *)build_key->dptr, k, j->rows, m);
    rc = radix_sort<unsigned>(d, s, k, kt, j->rows, rt, m, 32);
    if (!rc) join_decode_kernel<int, unsigned><<<g, 256, 0, s>>>(k, m);
  } else {
    unsigned long long *k = (unsigned long long *)j->keys, *kt = tmp.as<unsigned long long>();
    unsigned *rt = (unsigned *)(kt + m);
    join_encode_kernel<long long, unsigned long long><<<g, 256, 0, s>>>((const long long *)build_key->dptr, k, j->rows, m);
    rc = radix_sort<unsigned long long>(d, s, k, kt, j->rows, rt, m, 64);
    if (!rc) join_decode_kernel<long long, unsigned long long><<<g, 256, 0, s>>>(k, m);
  }
  stats().launches += 2;
  if (rc) return bail(1);
  if (cudaGetLastError() != cudaSuccess) return bail(fail("CUDA error: join build launch failed"));
  // Dense key range (ids of a dimension table): tabulate every key value of [min, max] as well --
  // 8 B per value, at most join.direct_factor (8) values per build row
  long long lo = 0, hi = -1;
  if (build_key->dtype == WDB_INT32) {
    int ends[2];
    if (cudaMemcpyAsync(&ends[0], j->keys, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(&ends[1], (const int *)j->keys + (m - 1), 4, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
      return bail(fail("CUDA error: join build failed (%s)", cudaGetErrorString(cudaGetLastError())));
    lo = ends[0]; hi = ends[1];
  } else {
    long long ends[2];
    if (cudaMemcpyAsync(&ends[0], j->keys, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(&ends[1], (const long long *)j->keys + (m - 1), 8, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
      return bail(fail("CUDA error: join build failed (%s)", cudaGetErrorString(cudaGetLastError())));
    lo = ends[0]; hi = ends[1];
  }
  const unsigned long long span = (unsigned long long)hi - (unsigned long long)lo + 1ull;   // 0 on the full int64 range
  const unsigned long long limit = std::max<unsigned long long>((unsigned long long)opt("join.direct_factor", 8) * (unsigned long long)m, 1ull << 16);
  if (opt("join.direct", 1) && span != 0 && span <= limit && span < (1ull << 28)) {
    Scratch first;
    if (first.alloc(4 * (size_t)(span + 1), s) != cudaSuccess) { cudaGetLastError(); return 0; }                              // no room: probes binary-search
    if (cudaMalloc((void **)&j->tab, 8 * (size_t)span) != cudaSuccess) { cudaGetLastError(); j->tab = nullptr; return 0; }
    j->lo = lo;
    j->span = (long long)span;
    const unsigned gf = join_grid(d, (long long)span + 1);
    if (build_key->dtype == WDB_INT32) join_first_kernel<int><<<gf, 256, 0, s>>>((const int *)j->keys, (unsigned)m, lo, (long long)span, first.as<unsigned>());
    else join_first_kernel<long long><<<gf, 256, 0, s>>>((const long long *)j->keys, (unsigned)m, lo, (long long)span, first.as<unsigned>());
    join_tab_kernel<<<gf, 256, 0, s>>>(first.as<unsigned>(), j->rows, (long long)span, j->tab);
    stats().launches += 2;
    if (cudaGetLastError() != cudaSuccess) return bail(fail("CUDA error: join build launch failed"));
  }
  return 0;
}

int wdb_join_destroy(wdb_join_t *j) {
  if (!j) return 0;
  if (j->dev) cudaSetDevice(j->dev->id);
  if (j->keys || j->rows || j->counted.buf) cudaDeviceSynchronize();
  if (j->keys) cudaFree(j->keys);
  if (j->rows) cudaFree(j->rows);
  if (j->tab) cudaFree(j->tab);
  if (j->counted.buf) cudaFree(j->counted.buf);
  delete j;
  return 0;
}

int wdb_join_info(const wdb_join_t *j, int64_t *build_rows, int *key_dtype, int64_t *direct_span) {
  if (!j) return fail("null join index");
  if (build_rows) *build_rows = j->m;
  if (key_dtype) *key_dtype = j->key_dtype;
  if (direct_span) *direct_span = j->tab ? j->span : 0;
  return 0;
}

int wdb_join_probe(wdb_join_t *j, void *stream, const wdb_col_t *probe_key, int64_t *d_probe_rows, int64_t *d_build_rows, int64_t cap,
                   int64_t *h_pairs) {
  if (!j || !probe_key) return fail("null argument");
  if (probe_key->dtype != WDB_INT32 && probe_key->dtype != WDB_INT64)
    return fail("JOIN needs integer key columns (%s is not one)", probe_key->name ? probe_key->name : "?");
  Device *d = j->dev;
  WDB_CUDA(cudaSetDevice(d->id));
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = probe_key->len;
  if (n < 0) return fail("negative row count");
  if (h_pairs) *h_pairs = 0;
  if (n == 0 || j->m == 0) return 0;
  const long long ntiles = (n + kJoinTile - 1) / kJoinTile;
  if (ntiles > 0x7fffffffll) return fail("the probe side of a JOIN is limited to 2^41 rows");
  const bool b32 = j->key_dtype == WDB_INT32, p32 = probe_key->dtype == WDB_INT32;
  if (b32 && p32) return join_probe_typed<int, int>(j, s, probe_key->dptr, n, ntiles, d_probe_rows, d_build_rows, cap, h_pairs);
  if (b32) return join_probe_typed<int, long long>(j, s, probe_key->dptr, n, ntiles, d_probe_rows, d_build_rows, cap, h_pairs);
  if (p32) return join_probe_typed<long long, int>(j, s, probe_key->dptr, n, ntiles, d_probe_rows, d_build_rows, cap, h_pairs);
  return join_probe_typed<long long, long long>(j, s, probe_key->dptr, n, ntiles, d_probe_rows, d_build_rows, cap, h_pairs);
}

int wdb_gather(int device, void *stream, const wdb_col_t *src, const int64_t *d_rows, int64_t count, void *d_dst) {
  Device *d;
  if (get_device(device, &d)) return 1;
  if (count < 0) return fail("negative count");
  if (count == 0) return 0;
  if (!src || !d_dst) return fail("null argument");
  cudaStream_t s = (cudaStream_t)stream;
  size_t esz = 0;
  switch (src->dtype) {
  case WDB_INT32: case WDB_FLOAT32: esz = 4; break;
  case WDB_INT64: case WDB_FLOAT64: esz = 8; break;
  default: return fail("column %s has a non-numeric type", src->name ? src->name : "?");
  }
  if (!d_rows) {                                    // identity row map
    if (count > src->len) return fail("gather of %lld rows from a column of %lld", (long long)count, (long long)src->len);
    WDB_CUDA(cudaMemcpyAsync(d_dst, src->dptr, esz * (size_t)count, cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  const unsigned g = join_grid(d, count);
  if (esz == 4) gather_rows_kernel<unsigned><<<g, 256, 0, s>>>((const unsigned *)src->dptr, (const long long *)d_rows, (unsigned *)d_dst, count);
  else gather_rows_kernel<unsigned long long><<<g, 256, 0, s>>>((const unsigned long long *)src->dptr, (const long long *)d_rows, (unsigned long long *)d_dst, count);
  stats().launches++;
  WDB_CUDA(cudaGetLastError());
  return 0;
}
}
