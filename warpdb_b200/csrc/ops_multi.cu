// ops_multi.cu -- host-buffer / multi-GPU entry point and column statistics.
//
// wdb_multi_project_filter_host replaces run_multi_gpu_jit_host (src/multi_gpu_utils.cpp:5-63):
// the same contiguous shards chunk = ceil(n/ndev) (:24-31) and the same host-vector-in /
// host-vector-out contract, but every device runs concurrently from its own host thread, and each
// shard is streamed through a ring of device buffers on separate streams so the H2D copy of chunk
// i+1, the kernel of chunk i and the D2H copy of chunk i-1 overlap (the reference uploads, compiles,
// launches and downloads one device after the other with synchronous copies).
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>

#include "core.hpp"

namespace wdb {
int run_project(Device *d, cudaStream_t stream, const wdb_col_t *cols, int ncols, const char *expr, const char *cond,
                float *d_out, int64_t n, int mode, const unsigned char *zmask = nullptr, int zshift = 0);
int run_compact(Device *d, cudaStream_t stream, const wdb_col_t *cols, int ncols, const char *expr, const char *expr2,
                const char *cond, float *d_out, float *d_out2, int64_t n, int64_t *d_count, int64_t *h_count);

struct ShardJob {
  int dev = 0;
  int64_t start = 0, end = 0;
  int64_t count = 0;     // survivors (COMPACT) or rows
  int rc = 0;
  std::string err;
};

static int run_shard(ShardJob *job, const wdb_col_t *h_cols, int ncols, const char *expr, const char *cond, float *h_out, int mode) {
  Device *d;
  if (get_device(job->dev, &d)) return 1;
  const bool has_cond = cond && *cond;
  const int64_t rows = job->end - job->start;
  if (rows <= 0) return 0;
  std::vector<UsedCol> used = find_used_columns(h_cols, ncols, {expr, has_cond ? cond : ""});
  for (const auto &u : used)
    if (dtype_size(u.dtype) == 0) return fail("column %s has a non-numeric type and cannot be read on the GPU", u.name.c_str());
  const int64_t chunk = std::min<int64_t>(rows, std::max<int64_t>(1 << 16, opt("multi.chunk_rows", 1 << 27)));
  const int nslots = (int)std::max<int64_t>(1, std::min<int64_t>(opt("multi.slots", 2), (rows + chunk - 1) / chunk));
  size_t row_bytes = 4;
  for (const auto &u : used) row_bytes += dtype_size(u.dtype);
  // the ring of device buffers is kept per device across calls (cudaMalloc/cudaFree of a multi-GB
  // ring cost 50-250 ms per call and made the end-to-end time erratic); one host-buffer call per
  // device at a time
  static std::mutex ring_mu[64];
  static char *ring_ptr[64];
  static size_t ring_bytes[64];
  std::lock_guard<std::mutex> ring_lock(ring_mu[job->dev]);
  const size_t need = (size_t)nslots * (size_t)chunk * row_bytes + 256 * (used.size() + 2) * nslots;
  if (need > ring_bytes[job->dev]) {
    if (ring_ptr[job->dev]) cudaFree(ring_ptr[job->dev]);
    ring_ptr[job->dev] = nullptr;
    ring_bytes[job->dev] = 0;
    WDB_CUDA(cudaMalloc((void **)&ring_ptr[job->dev], need));
    ring_bytes[job->dev] = need;
  }
  char *pool = ring_ptr[job->dev];
  struct StreamSet {   // destroyed on every exit path
    std::vector<cudaStream_t> v;
    ~StreamSet() { for (auto s : v) if (s) cudaStreamDestroy(s); }
  } stream_set;
  stream_set.v.assign(nslots, nullptr);
  std::vector<cudaStream_t> &streams = stream_set.v;
  for (auto &s : streams) WDB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  // carve: per slot, one buffer per used column (256-byte aligned) + the output
  std::vector<std::vector<char *>> in(nslots, std::vector<char *>(used.size()));
  std::vector<float *> out(nslots);
  char *p = pool;
  auto bump = [&](size_t bytes) { char *r = p; p += (bytes + 255) & ~(size_t)255; return r; };
  for (int s = 0; s < nslots; ++s) {
    for (size_t k = 0; k < used.size(); ++k) in[s][k] = bump((size_t)chunk * dtype_size(used[k].dtype));
    out[s] = (float *)bump((size_t)chunk * 4);
  }
  int rc = 0;
  int64_t written = 0;
  int c = 0;
  for (int64_t off = 0; off < rows && !rc; off += chunk, ++c) {
    const int s = c % nslots;
    const int64_t m = std::min<int64_t>(chunk, rows - off);
    std::vector<wdb_col_t> dcols(h_cols, h_cols + ncols);
    for (auto &dc : dcols) { dc.dptr = nullptr; dc.len = m; }
    for (size_t k = 0; k < used.size() && !rc; ++k) {
      const wdb_col_t &hc = h_cols[used[k].table_index];
      const size_t sz = dtype_size(used[k].dtype);
      cudaError_t e = cudaMemcpyAsync(in[s][k], (const char *)hc.dptr + (size_t)(job->start + off) * sz, (size_t)m * sz,
                                      cudaMemcpyHostToDevice, streams[s]);
      if (e != cudaSuccess) rc = fail("CUDA error: %s (H2D)", cudaGetErrorString(e));
      dcols[used[k].table_index].dptr = in[s][k];
    }
    if (rc) break;
    if (mode == WDB_COMPACT && has_cond) {
      int64_t cnt = 0;
      rc = run_compact(d, streams[s], dcols.data(), ncols, expr, nullptr, cond, out[s], nullptr, m, nullptr, &cnt);
      if (!rc && cnt > 0) {
        cudaError_t e = cudaMemcpyAsync(h_out + job->start + written, out[s], (size_t)cnt * 4, cudaMemcpyDeviceToHost, streams[s]);
        if (e != cudaSuccess) rc = fail("CUDA error: %s (D2H)", cudaGetErrorString(e));
      }
      written += cnt;
    } else {
      rc = run_project(d, streams[s], dcols.data(), ncols, expr, cond, out[s], m, mode == WDB_COMPACT ? WDB_DENSE : mode);
      if (!rc) {
        cudaError_t e = cudaMemcpyAsync(h_out + job->start + off, out[s], (size_t)m * 4, cudaMemcpyDeviceToHost, streams[s]);
        if (e != cudaSuccess) rc = fail("CUDA error: %s (D2H)", cudaGetErrorString(e));
      }
      written += m;
    }
  }
  for (auto &s : streams) {
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess && !rc) rc = fail("CUDA error: %s (stream sync)", cudaGetErrorString(e));
  }
  job->count = written;
  return rc;
}

// ---- column min/max (TableStats of include/csv_loader.hpp:22-37; feeds the optimizer) ----------
template <class T>
__global__ void __launch_bounds__(256) minmax_kernel(const T *__restrict__ v, long long n, double *__restrict__ out /* [2*grid] */) {
  double lo = 1.0 / 0.0, hi = -1.0 / 0.0;
  auto take = [&](T raw) {
    const double x = (double)raw;
    lo = x < lo ? x : lo;
    hi = x > hi ? x : hi;
  };
  long long done = 0;
  if constexpr (sizeof(T) == 4) {   // 128-bit loads over the aligned body
    if ((reinterpret_cast<unsigned long long>(v) & 15ull) == 0) {
      const uint4 *p = reinterpret_cast<const uint4 *>(v);
      const long long nvec = n >> 2;
      for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nvec; k += (long long)gridDim.x * blockDim.x) {
        const uint4 q = __ldg(p + k);
        take(*reinterpret_cast<const T *>(&q.x));
        take(*reinterpret_cast<const T *>(&q.y));
        take(*reinterpret_cast<const T *>(&q.z));
        take(*reinterpret_cast<const T *>(&q.w));
      }
      done = nvec << 2;
    }
  }
  for (long long i = done + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) take(v[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ double s_lo[8], s_hi[8];
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < 8; ++w) { lo = fmin(lo, s_lo[w]); hi = fmax(hi, s_hi[w]); }
    out[2 * blockIdx.x] = lo;
    out[2 * blockIdx.x + 1] = hi;
  }
}
}  // namespace wdb

using namespace wdb;

extern "C" {

int wdb_multi_project_filter_host(int ndev, const int *devices, const wdb_col_t *h_cols, int ncols, const char *expr,
                                  const char *cond, float *h_out, int64_t n, int mode, int64_t *h_count) {
  if (!expr || !*expr) return fail("empty expression");
  if (n < 0) return fail("negative row count");
  if (mode != WDB_DENSE && mode != WDB_COMPACT && mode != WDB_DENSE_ZERO) return fail("invalid mode %d", mode);
  int avail = 0;
  wdb_device_count(&avail);
  if (avail == 0) return fail("CUDA error: no CUDA device available; warpcore has no CPU fallback");
  if (ndev <= 0) ndev = avail;                       // 0 = all devices (cudaGetDeviceCount: src/multi_gpu_utils.cpp:8-9)
  std::vector<ShardJob> jobs(ndev);
  for (int i = 0; i < ndev; ++i) {
    jobs[i].dev = devices ? devices[i] : i;
    if (jobs[i].dev < 0 || jobs[i].dev >= avail) return fail("invalid device id %d", jobs[i].dev);
    wdb_shard_range(n, ndev, i, &jobs[i].start, &jobs[i].end);
  }
  std::vector<std::thread> th;
  for (int i = 0; i < ndev; ++i)
    th.emplace_back([&, i]() {
      jobs[i].rc = run_shard(&jobs[i], h_cols, ncols, expr, cond, h_out, mode);
      if (jobs[i].rc) jobs[i].err = wdb_last_error();
    });
  for (auto &t : th) t.join();
  for (auto &j : jobs)
    if (j.rc) { set_error(j.err); return 1; }
  int64_t total = 0;
  if (mode == WDB_COMPACT && cond && *cond) {      // pack the per-device survivor runs (row order is preserved)
    for (auto &j : jobs) {
      if (j.count > 0 && total != j.start) memmove(h_out + total, h_out + j.start, (size_t)j.count * 4);
      total += j.count;
    }
  } else
    total = n;
  if (h_count) *h_count = total;
  return 0;
}

int wdb_column_minmax(int device, void *stream, const wdb_col_t *col, double *h_min, double *h_max) {
  Device *d;
  if (get_device(device, &d)) return 1;
  if (!col || !h_min || !h_max) return fail("null argument");
  if (col->len <= 0) { *h_min = 0; *h_max = 0; return 0; }
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)std::min<int64_t>((col->len + 255) / 256, (int64_t)d->num_sms * 8);
  Scratch scratch;
  WDB_CUDA(scratch.alloc(sizeof(double) * 2 * grid, s));
  double *part = scratch.as<double>();
  switch (col->dtype) {
  case WDB_INT32: minmax_kernel<int><<<grid, 256, 0, s>>>((const int *)col->dptr, col->len, part); break;
  case WDB_INT64: minmax_kernel<long long><<<grid, 256, 0, s>>>((const long long *)col->dptr, col->len, part); break;
  case WDB_FLOAT32: minmax_kernel<float><<<grid, 256, 0, s>>>((const float *)col->dptr, col->len, part); break;
  case WDB_FLOAT64: minmax_kernel<double><<<grid, 256, 0, s>>>((const double *)col->dptr, col->len, part); break;
  default: return fail("column %s has a non-numeric type", col->name ? col->name : "?");
  }
  stats().launches++;
  WDB_CUDA(cudaGetLastError());
  std::vector<double> h(2 * grid);
  WDB_CUDA(cudaMemcpyAsync(h.data(), part, sizeof(double) * 2 * grid, cudaMemcpyDeviceToHost, s));
  WDB_CUDA(cudaStreamSynchronize(s));
  double lo = h[0], hi = h[1];
  for (unsigned i = 1; i < grid; ++i) { lo = std::min(lo, h[2 * i]); hi = std::max(hi, h[2 * i + 1]); }
  *h_min = lo;
  *h_max = hi;
  return 0;
}
}
