// ops_zonemap.cu -- zone maps (per-zone min/max of a column) and zone-pruned filter/project.
//
// The reference's optimizer is a stub: analyze_condition ignores its inputs and TableStats is never
// filled (src/optimizer.cpp:13-17,35).  Here a zone map is built once per column (one
// streaming pass) and a WHERE clause made of `col <op> const` terms joined by AND is turned into a
// per-zone live mask; the filter kernels never load zones no row of which can pass.  On clustered
// or sorted columns this removes most of the HBM traffic; on uniformly random data every zone stays
// live and the pruned kernels cost the same as the plain ones.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "agg.hpp"

struct wdb_zonemap {
  wdb::Device *dev = nullptr;
  int dtype = 0;
  int64_t n = 0;
  int zshift = 12;
  int64_t nzones = 0;
  double *mins = nullptr, *maxs = nullptr;   // device
};

namespace wdb {
int run_project(Device *d, cudaStream_t stream, const wdb_col_t *cols, int ncols, const char *expr, const char *cond,
                float *d_out, int64_t n, int mode, const unsigned char *zmask = nullptr, int zshift = 0);
int run_compact_ex(Device *d, cudaStream_t stream, const wdb_col_t *cols, int ncols, const char *expr, const char *expr2,
                   const char *cond, float *d_out, float *d_out2, int64_t n, int64_t *d_count, int64_t *h_count,
                   int thresh, float tau, int64_t out_cap, const unsigned char *zmask = nullptr, int zshift = 0);

// value as the filter kernel sees it when the column is compared with a float literal:
// integers are converted to float first (usual arithmetic conversions), doubles stay double
template <class T> __device__ __forceinline__ double as_compared(T x) { return (double)(float)x; }
template <> __device__ __forceinline__ double as_compared<float>(float x) { return (double)x; }
template <> __device__ __forceinline__ double as_compared<double>(double x) { return x; }

// One WARP per zone (8 zones per CTA: a CTA per 16 KB zone is launch-bound).  4-byte columns whose
// zone is complete and 16-byte aligned are read with 128-bit loads, 8 in flight per lane; anything
// else takes the scalar loop.  Reduction with shuffles only: no shared memory, no block barrier.
template <class T>
__global__ void __launch_bounds__(256) zonemap_build_kernel(const T *__restrict__ v, long long n, int zshift, long long nzones,
                                                            double *__restrict__ mins, double *__restrict__ maxs) {
  const unsigned lane = threadIdx.x & 31u;
  const long long zone = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (zone >= nzones) return;                       // warp-uniform
  const long long b = zone << zshift, e = min(b + (1ll << zshift), n);
  double lo = 1.0 / 0.0, hi = -1.0 / 0.0;
  bool nan = false;
  auto take = [&](T raw) {
    const double x = as_compared<T>(raw);
    nan |= (x != x);
    lo = x < lo ? x : lo;
    hi = x > hi ? x : hi;
  };
  bool vectorised = false;
  if constexpr (sizeof(T) == 4) {
    if (((e - b) & 1023) == 0 && (reinterpret_cast<unsigned long long>(v + b) & 15ull) == 0) {   // whole rounds of 8 x 32 x 4 rows
      vectorised = true;
      const uint4 *p = reinterpret_cast<const uint4 *>(v + b);
      const long long nvec = (e - b) >> 2;
      for (long long k0 = 0; k0 < nvec; k0 += 256) {
        uint4 q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = __ldg(p + k0 + u * 32 + lane);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          take(*reinterpret_cast<const T *>(&q[u].x));
          take(*reinterpret_cast<const T *>(&q[u].y));
          take(*reinterpret_cast<const T *>(&q[u].z));
          take(*reinterpret_cast<const T *>(&q[u].w));
        }
      }
    }
  }
  if (!vectorised)
    for (long long i = b + lane; i < e; i += 32) take(v[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  nan = __any_sync(0xffffffffu, nan);
  if (lane == 0) {   // a zone holding a NaN is never pruned
    mins[zone] = nan ? -1.0 / 0.0 : lo;
    maxs[zone] = nan ? 1.0 / 0.0 : hi;
  }
}

struct DevPred { const double *mins, *maxs; int op; double value; };
constexpr int kMaxPreds = 8;
struct DevPreds { DevPred p[kMaxPreds]; int n; };

__global__ void zone_mask_kernel(DevPreds P, long long nzones, unsigned char *__restrict__ mask, unsigned long long *__restrict__ live) {
  unsigned long long mine = 0;
  for (long long z = (long long)blockIdx.x * blockDim.x + threadIdx.x; z < nzones; z += (long long)gridDim.x * blockDim.x) {
    bool keep = true;
    for (int i = 0; i < P.n; ++i) {
      const double lo = P.p[i].mins[z], hi = P.p[i].maxs[z], c = P.p[i].value;
      bool may;
      switch (P.p[i].op) {
      case 0: may = hi > c; break;              // >
      case 1: may = hi >= c; break;             // >=
      case 2: may = lo < c; break;              // <
      case 3: may = lo <= c; break;             // <=
      case 4: may = lo <= c && c <= hi; break;  // ==
      default: may = !(lo == c && hi == c); break;  // !=
      }
      keep = keep && may;
    }
    mask[z] = keep ? 1 : 0;
    mine += keep ? 1ull : 0ull;
  }
  if (mine) atomicAdd(live, mine);
}

// min/max of zones [z0, z1) of the column at device address `base` (row 0 of the column)
static int zonemap_launch(wdb_zonemap *z, cudaStream_t s, const void *base, long long z0, long long z1) {
  if (z1 <= z0) return 0;
  const long long nz = z1 - z0, row0 = z0 << z->zshift, n = z->n - row0;
  const unsigned g = (unsigned)((nz + 7) / 8);
  double *mins = z->mins + z0, *maxs = z->maxs + z0;
  switch (z->dtype) {
  case WDB_INT32: zonemap_build_kernel<int><<<g, 256, 0, s>>>((const int *)base + row0, n, z->zshift, nz, mins, maxs); break;
  case WDB_INT64: zonemap_build_kernel<long long><<<g, 256, 0, s>>>((const long long *)base + row0, n, z->zshift, nz, mins, maxs); break;
  case WDB_FLOAT32: zonemap_build_kernel<float><<<g, 256, 0, s>>>((const float *)base + row0, n, z->zshift, nz, mins, maxs); break;
  default: zonemap_build_kernel<double><<<g, 256, 0, s>>>((const double *)base + row0, n, z->zshift, nz, mins, maxs); break;
  }
  stats().launches++;
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// which zones can hold a row passing every predicate -> maximal runs of live zones as row ranges (host);
// synchronises `s` (the mask is read back: nzones bytes)
int zone_live_ranges(Device *d, cudaStream_t s, const wdb_prune_t *preds, int npreds, int64_t n, std::vector<std::pair<int64_t, int64_t>> *ranges,
                     int64_t *live, int64_t *nzones) {
  if (npreds > kMaxPreds) npreds = kMaxPreds;
  DevPreds P;
  P.n = npreds;
  const wdb_zonemap *z0 = preds[0].zonemap;
  for (int i = 0; i < npreds; ++i) {
    const wdb_zonemap *z = preds[i].zonemap;
    if (!z || z->n != n || z->zshift != z0->zshift) return fail("zone maps must cover the table's %lld rows with one zone size", (long long)n);
    if (preds[i].op < 0 || preds[i].op > 5) return fail("invalid pruning operator %d", preds[i].op);
    P.p[i] = DevPred{z->mins, z->maxs, preds[i].op, preds[i].value};
  }
  const int64_t nz = z0->nzones;
  *nzones = nz;
  *live = 0;
  ranges->clear();
  if (nz == 0) return 0;
  Scratch scratch;
  WDB_CUDA(scratch.alloc((size_t)nz + 16, s));
  char *buf = scratch.as<char>();
  WDB_CUDA(cudaMemsetAsync(buf, 0, 8, s));
  zone_mask_kernel<<<(unsigned)std::min<int64_t>((nz + 255) / 256, 4096), 256, 0, s>>>(P, nz, (unsigned char *)(buf + 16), (unsigned long long *)buf);
  stats().launches++;
  WDB_CUDA(cudaGetLastError());
  std::vector<unsigned char> h((size_t)nz);
  WDB_CUDA(cudaMemcpyAsync(h.data(), buf + 16, (size_t)nz, cudaMemcpyDeviceToHost, s));
  WDB_CUDA(cudaStreamSynchronize(s));
  const int64_t zr = 1ll << z0->zshift;
  for (int64_t z = 0; z < nz;) {
    if (!h[z]) { ++z; continue; }
    int64_t e = z;
    while (e < nz && h[e]) ++e;
    ranges->push_back({z * zr, std::min<int64_t>(e * zr, n)});
    *live += e - z;
    z = e;
  }
  return 0;
}
int topk_candidates(Device *d, cudaStream_t s, const wdb_col_t *cols, int ncols, const char *key, const char *val, const char *cond,
                    bool desc, int K, int64_t n, int64_t row_base, char *cand);
int topk_merge_launch(cudaStream_t s, const char *gathered, int nparts, int K, bool desc, int offset, float *out_vals, float *out_keys,
                      long long *out_n);
std::string order_key(const char *key_expr, bool desc);

// columns of the row range [start, end): pointers advanced, lengths cut
static std::vector<wdb_col_t> slice_cols(const wdb_col_t *cols, int ncols, int64_t start, int64_t end) {
  std::vector<wdb_col_t> out(cols, cols + ncols);
  for (auto &c : out) {
    const size_t sz = (size_t)dtype_size(c.dtype);
    if (c.dptr && sz) c.dptr = (const char *)c.dptr + (size_t)start * sz;
    c.len = end - start;
  }
  return out;
}

// pinned staging ring for uploads from pageable host memory: two slots per device, kept across calls
struct StagingRing {
  std::mutex mu;
  char *host = nullptr;
  size_t slot_bytes = 0;
  cudaEvent_t done[2] = {nullptr, nullptr};
};
static StagingRing g_ring[64];

}  // namespace wdb

using namespace wdb;

extern "C" {

int wdb_upload_column(int device, void *stream, const wdb_col_t *h_col, void *d_dst, int64_t zone_rows, wdb_zonemap_t **out_zm,
                      double *h_min, double *h_max) {
  Device *d;
  if (get_device(device, &d)) return 1;
  if (!h_col || (!d_dst && h_col->len > 0)) return fail("null argument");
  const size_t esz = (size_t)dtype_size(h_col->dtype);
  if (esz == 0) return fail("column %s has a non-numeric type", h_col->name ? h_col->name : "?");
  const int64_t n = h_col->len;
  cudaStream_t s = (cudaStream_t)stream;
  wdb_zonemap *z = nullptr;
  if (out_zm && zone_rows >= 0) {
    // the map's arrays exist before the first chunk lands; zones are filled chunk by chunk below
    if (zone_rows == 0) zone_rows = 4096;
    if (zone_rows < 2048 || (zone_rows & (zone_rows - 1))) return fail("zone_rows must be a power of two >= 2048");
    z = new wdb_zonemap();
    z->dev = d;
    z->dtype = h_col->dtype;
    z->n = n;
    z->zshift = 0;
    while ((1ll << z->zshift) < zone_rows) ++z->zshift;
    z->nzones = (n + zone_rows - 1) / zone_rows;
    const size_t bytes = sizeof(double) * (size_t)std::max<int64_t>(z->nzones, 1);
    if (cudaMalloc((void **)&z->mins, 2 * bytes) != cudaSuccess) { delete z; return fail("CUDA error: out of memory (zone map)"); }
    z->maxs = z->mins + std::max<int64_t>(z->nzones, 1);
  }
  auto bail = [&](const char *what) {
    if (z) { cudaFree(z->mins); delete z; }
    return fail("CUDA error: %s (%s)", cudaGetErrorString(cudaGetLastError()), what);
  };
  // chunks of 64 MB (a whole number of zones): copy(i+1) overlaps the statistics kernel of chunk i
  const int64_t chunk_rows = std::max<int64_t>(1 << 20, opt("upload.chunk_bytes", 64 << 20) / (int64_t)esz) & ~(int64_t)65535;
  cudaPointerAttributes at{};
  const bool pinned = cudaPointerGetAttributes(&at, h_col->dptr) == cudaSuccess && (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged);
  cudaGetLastError();
  StagingRing &ring = g_ring[device];
  std::unique_lock<std::mutex> lock(ring.mu, std::defer_lock);
  if (!pinned && n > 0) {
    lock.lock();
    const size_t need = (size_t)chunk_rows * esz;
    if (ring.slot_bytes < need) {
      if (ring.host) cudaFreeHost(ring.host);
      ring.host = nullptr;
      ring.slot_bytes = 0;
      if (cudaHostAlloc((void **)&ring.host, 2 * need, cudaHostAllocDefault) != cudaSuccess) return bail("cudaHostAlloc staging ring");
      ring.slot_bytes = need;
      for (auto &e : ring.done)
        if (!e && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return bail("cudaEventCreate");
    }
  }
  int c = 0;
  for (int64_t row = 0; row < n; row += chunk_rows, ++c) {
    const int64_t m = std::min<int64_t>(chunk_rows, n - row);
    const char *src = (const char *)h_col->dptr + (size_t)row * esz;
    if (!pinned) {   // pageable source: CPU copy into the pinned slot while the previous slot is on the wire
      char *slot = ring.host + (size_t)(c & 1) * ring.slot_bytes;
      if (c >= 2 && cudaEventSynchronize(ring.done[c & 1]) != cudaSuccess) return bail("staging slot wait");
      memcpy(slot, src, (size_t)m * esz);
      src = slot;
    }
    if (cudaMemcpyAsync((char *)d_dst + (size_t)row * esz, src, (size_t)m * esz, cudaMemcpyHostToDevice, s) != cudaSuccess) return bail("H2D");
    if (!pinned && cudaEventRecord(ring.done[c & 1], s) != cudaSuccess) return bail("event record");
    if (z && zonemap_launch(z, s, d_dst, row >> z->zshift, (row + m + (1ll << z->zshift) - 1) >> z->zshift)) return bail("zone map kernel");
  }
  if (!pinned && n > 0 && cudaStreamSynchronize(s) != cudaSuccess) return bail("stream sync");   // the ring is free again when we leave
  if (h_min && h_max) {   // exact min/max (TableStats): one more pass over the resident column, 0.6 ms per 1e9 rows
    wdb_col_t dc = *h_col;
    dc.dptr = d_dst;
    if (wdb_column_minmax(device, stream, &dc, h_min, h_max)) { if (z) { cudaFree(z->mins); delete z; } return 1; }
  }
  if (out_zm) *out_zm = z;
  return 0;
}

int wdb_zonemap_build(int device, void *stream, const wdb_col_t *col, int64_t zone_rows, wdb_zonemap_t **out) {
  Device *d;
  if (get_device(device, &d)) return 1;
  if (!col || !out) return fail("null argument");
  if (zone_rows == 0) zone_rows = 4096;
  if (zone_rows < 2048 || (zone_rows & (zone_rows - 1))) return fail("zone_rows must be a power of two >= 2048");
  if (dtype_size(col->dtype) == 0) return fail("column %s has a non-numeric type", col->name ? col->name : "?");
  wdb_zonemap *z = new wdb_zonemap();
  z->dev = d;
  z->dtype = col->dtype;
  z->n = col->len;
  z->zshift = 0;
  while ((1ll << z->zshift) < zone_rows) ++z->zshift;
  z->nzones = (col->len + zone_rows - 1) / zone_rows;
  const size_t bytes = sizeof(double) * (size_t)std::max<int64_t>(z->nzones, 1);
  if (cudaMalloc((void **)&z->mins, 2 * bytes) != cudaSuccess) { delete z; return fail("CUDA error: out of memory (zone map)"); }
  z->maxs = z->mins + std::max<int64_t>(z->nzones, 1);
  cudaStream_t s = (cudaStream_t)stream;
  if (z->nzones > 0 && zonemap_launch(z, s, col->dptr, 0, z->nzones)) { cudaFree(z->mins); delete z; return fail("CUDA error: zone map build failed"); }
  *out = z;
  return 0;
}

int wdb_zonemap_destroy(wdb_zonemap_t *z) {
  if (!z) return 0;
  cudaSetDevice(z->dev->id);
  cudaFree(z->mins);
  delete z;
  return 0;
}

int wdb_zonemap_info(const wdb_zonemap_t *z, int64_t *zone_rows, int64_t *nzones) {
  if (!z) return fail("null zone map");
  if (zone_rows) *zone_rows = 1ll << z->zshift;
  if (nzones) *nzones = z->nzones;
  return 0;
}

int wdb_project_filter_pruned(int device, void *stream, const wdb_col_t *cols, int ncols, const char *expr, const char *cond,
                              float *d_out, int64_t n, int mode, int64_t *d_count, int64_t *h_count, const wdb_prune_t *preds,
                              int npreds, int64_t *h_zones_live) {
  if (npreds <= 0 || !preds) return wdb_project_filter(device, stream, cols, ncols, expr, cond, d_out, n, mode, d_count, h_count);
  if (!expr || !*expr) return fail("empty expression");
  if (!cond || !*cond) return fail("zone-map pruning needs a condition");
  if (npreds > kMaxPreds) npreds = kMaxPreds;   // any subset of a conjunction is a valid (weaker) pruning test
  Device *d;
  if (get_device(device, &d)) return 1;
  cudaStream_t s = (cudaStream_t)stream;
  DevPreds P;
  P.n = npreds;
  const wdb_zonemap *z0 = preds[0].zonemap;
  for (int i = 0; i < npreds; ++i) {
    const wdb_zonemap *z = preds[i].zonemap;
    if (!z || z->n != n || z->zshift != z0->zshift) return fail("zone maps must cover the table's %lld rows with one zone size", (long long)n);
    if (preds[i].op < 0 || preds[i].op > 5) return fail("invalid pruning operator %d", preds[i].op);
    P.p[i] = DevPred{z->mins, z->maxs, preds[i].op, preds[i].value};
  }
  const int64_t nz = z0->nzones;
  Scratch scratch;
  WDB_CUDA(scratch.alloc((size_t)std::max<int64_t>(nz, 1) + 16, s));
  char *buf = scratch.as<char>();
  unsigned long long *d_live = (unsigned long long *)buf;
  unsigned char *mask = (unsigned char *)(buf + 16);
  WDB_CUDA(cudaMemsetAsync(d_live, 0, 8, s));
  if (nz > 0) {
    zone_mask_kernel<<<(unsigned)std::min<int64_t>((nz + 255) / 256, 4096), 256, 0, s>>>(P, nz, mask, d_live);
    stats().launches++;
    WDB_CUDA(cudaGetLastError());
  }
  // No host round trip on the way: the pruned kernels consult the mask themselves (a blocking read
  // of the live count to choose a plan cost ~0.2 ms per query, more than the mask lookups it saved).
  int rc;
  if (mode == WDB_COMPACT)
    rc = run_compact_ex(d, s, cols, ncols, expr, nullptr, cond, d_out, nullptr, n, d_count, h_count, 0, 0.0f, n, mask, z0->zshift);
  else {
    rc = run_project(d, s, cols, ncols, expr, cond, d_out, n, mode, mask, z0->zshift);
    long long nn = n;
    if (!rc && d_count) WDB_CUDA(cudaMemcpyAsync(d_count, &nn, sizeof nn, cudaMemcpyHostToDevice, s));
    if (!rc && h_count) *h_count = n;
  }
  if (!rc && h_zones_live) {
    unsigned long long live = 0;
    WDB_CUDA(cudaMemcpyAsync(&live, d_live, 8, cudaMemcpyDeviceToHost, s));
    WDB_CUDA(cudaStreamSynchronize(s));
    *h_zones_live = (int64_t)live;
  } else if (!rc && (h_count || d_count))
    WDB_CUDA(cudaStreamSynchronize(s));
  return rc;
}

// GROUP BY / ORDER BY ... LIMIT with a WHERE clause over a zone-mapped table: the kernels run on the
// maximal runs of live zones only (zones are multiples of 2048 rows, so every run keeps the columns'
// vector alignment).  Worth it when the runs are few and cover under half of the table -- sorted or
// clustered columns; on a random layout everything is live and the plain call is taken.
int wdb_agg_consume_pruned(wdb_agg_t *t, void *stream, const wdb_col_t *cols, int ncols, const char *val_expr, const char *key_expr,
                           const char *cond, int64_t n, int64_t row_base, const wdb_prune_t *preds, int npreds, int64_t *h_zones_live) {
  if (npreds <= 0 || !preds || !cond || !*cond) return wdb_agg_consume(t, stream, cols, ncols, val_expr, key_expr, cond, n, row_base);
  if (!t) return fail("null table");
  Device *d = t->dev;
  WDB_CUDA(cudaSetDevice(d->id));
  std::vector<std::pair<int64_t, int64_t>> ranges;
  int64_t live = 0, nz = 0;
  if (zone_live_ranges(d, (cudaStream_t)stream, preds, npreds, n, &ranges, &live, &nz)) return 1;
  if (h_zones_live) *h_zones_live = live;
  if ((int64_t)ranges.size() > opt("prune.max_ranges", 64) || 2 * live > nz)
    return wdb_agg_consume(t, stream, cols, ncols, val_expr, key_expr, cond, n, row_base);
  for (const auto &r : ranges) {
    const std::vector<wdb_col_t> rc = slice_cols(cols, ncols, r.first, r.second);
    if (wdb_agg_consume(t, stream, rc.data(), ncols, val_expr, key_expr, cond, r.second - r.first, row_base + r.first)) return 1;
  }
  return 0;
}

int wdb_topk_pruned(int device, void *stream, const wdb_col_t *cols, int ncols, const char *key_expr, const char *val_expr, const char *cond,
                    int descending, int64_t k, int64_t offset, int64_t n, float *d_out_vals, float *d_out_keys, int64_t *h_n,
                    const wdb_prune_t *preds, int npreds, int64_t *h_zones_live) {
  const int64_t K = k + offset;
  if (npreds <= 0 || !preds || !cond || !*cond || k <= 0 || K > opt("topk.reg_max", 16))
    return wdb_topk(device, stream, cols, ncols, key_expr, val_expr, cond, descending, k, offset, n, d_out_vals, d_out_keys, h_n);
  if (!key_expr || !*key_expr) return fail("empty ORDER BY expression");
  if (!val_expr || !*val_expr) val_expr = key_expr;
  Device *d;
  if (get_device(device, &d)) return 1;
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<std::pair<int64_t, int64_t>> ranges;
  int64_t live = 0, nz = 0;
  if (zone_live_ranges(d, s, preds, npreds, n, &ranges, &live, &nz)) return 1;
  if (h_zones_live) *h_zones_live = live;
  if (ranges.empty() || (int64_t)ranges.size() > std::min<int64_t>(opt("prune.max_ranges", 64), 2048 / K) || 2 * live > nz)
    return wdb_topk(device, stream, cols, ncols, key_expr, val_expr, cond, descending, k, offset, n, d_out_vals, d_out_keys, h_n);
  // every run of live zones yields its K best (key, global row) pairs; the one-warp selection of the sharded
  // ORDER BY merges them (runs are ascending row ranges, so ties keep row order)
  const size_t per = (size_t)K * 16;
  Scratch scratch;
  WDB_CUDA(scratch.alloc(per * ranges.size() + 64, s));
  char *buf = scratch.as<char>();
  long long *d_cnt = (long long *)(buf + per * ranges.size());
  const std::string okey = order_key(key_expr, descending != 0);
  int rc = 0;
  for (size_t i = 0; i < ranges.size() && !rc; ++i) {
    const std::vector<wdb_col_t> cr = slice_cols(cols, ncols, ranges[i].first, ranges[i].second);
    rc = topk_candidates(d, s, cr.data(), ncols, okey.c_str(), val_expr, cond, descending != 0, (int)K, ranges[i].second - ranges[i].first, ranges[i].first,
                         buf + per * i);
  }
  if (!rc) rc = topk_merge_launch(s, buf, (int)ranges.size(), (int)K, descending != 0, (int)offset, d_out_vals, d_out_keys, d_cnt);
  long long cnt = 0;
  if (!rc && h_n) {
    if (cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) rc = fail("CUDA error: top-k count read-back");
    *h_n = cnt;
  }
  return rc;
}
}
