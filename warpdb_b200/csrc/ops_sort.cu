// ops_sort.cu -- stable LSD radix sort on the device (8-bit digits), the building block behind
//   wdb_sort_float / wdb_sort_pairs   (replace the <<<1,1>>> bubble sorts of src/jit.cpp:248-307)
//   key-ordered export of GROUP BY tables and the ORDER BY ... LIMIT candidate sort.
// Stability matters: the reference's bubble sorts only swap strictly out-of-order neighbours, so
// equal keys keep their original order in both directions; descending order is obtained by
// complementing the key, which keeps the LSD passes stable.
//
// Three kernels per digit pass (memory traffic: 2 reads + 1 write of the keys and payloads):
//   sort_hist_kernel     every CTA owns a contiguous chunk of tiles and counts its digits
//   sort_scan_kernel     exclusive scan of the digit-major (digit, CTA) counters -> global offsets
//   sort_scatter_kernel  the CTA walks its chunk tile by tile (4 096 keys): every warp ranks its 512
//                        keys digit by digit with ballots (stable: rounds and lanes in memory order,
//                        a warp-private digit counter carried from round to round, no atomics), the
//                        tile is permuted into digit order in shared memory and copied out so that
//                        equal digits leave as contiguous runs (coalesced writes instead of one
//                        32-byte sector per 4-byte key).
// Measured on a B200, 2^28 float keys: see DESIGN.md section 3.5 (the first version of this file -- one
// key per thread per round, scattered stores, a single-block scan over 16 M counters -- took 132 ms).
#include <algorithm>

#include "core.hpp"

namespace wdb {

constexpr int kSortBlock = 256;
constexpr int kSortWarps = kSortBlock / 32;
constexpr int kSortTileBytes = 16384;                 // keys of one tile in shared memory

template <class K> __device__ __forceinline__ unsigned digit_of(K key, int shift) { return (unsigned)(key >> shift) & 255u; }
template <class K> struct SortTile { static constexpr int kItems = kSortTileBytes / (int)sizeof(K) / kSortBlock; static constexpr int kTile = kItems * kSortBlock; };

template <class K>
__global__ void __launch_bounds__(kSortBlock) sort_hist_kernel(const K *__restrict__ keys, long long n, long long chunk, int shift,
                                                               unsigned *__restrict__ hist /* [256][nctas] */, long long nctas) {
  __shared__ unsigned s_hist[kSortWarps][256];       // warp-private histograms: native 32-bit shared atomics, no inter-warp contention
  for (int w = 0; w < kSortWarps; ++w) s_hist[w][threadIdx.x] = 0;
  __syncthreads();
  const unsigned warp = threadIdx.x >> 5;
  const long long begin = (long long)blockIdx.x * chunk, end = min(begin + chunk, n);
  for (long long i = begin + threadIdx.x; i < end; i += kSortBlock) atomicAdd(&s_hist[warp][digit_of(keys[i], shift)], 1u);
  __syncthreads();
  unsigned t = 0;
  for (int w = 0; w < kSortWarps; ++w) t += s_hist[w][threadIdx.x];
  hist[(long long)threadIdx.x * nctas + blockIdx.x] = t;
}

// exclusive scan of `m` counters into 64-bit offsets, one block
__global__ void __launch_bounds__(1024) sort_scan_kernel(const unsigned *__restrict__ hist, long long *__restrict__ offs, long long m) {
  __shared__ long long s_part[1024];
  const long long per = (m + 1023) / 1024;
  const long long b = (long long)threadIdx.x * per, e = min(b + per, m);
  long long sum = 0;
  for (long long i = b; i < e; ++i) sum += hist[i];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long run = 0;
    for (int i = 0; i < 1024; ++i) { long long t = s_part[i]; s_part[i] = run; run += t; }
  }
  __syncthreads();
  long long run = s_part[threadIdx.x];
  for (long long i = b; i < e; ++i) { offs[i] = run; run += hist[i]; }
}

template <class K, bool HAS_PAYLOAD>
__global__ void __launch_bounds__(kSortBlock) sort_scatter_kernel(const K *__restrict__ keys_in, K *__restrict__ keys_out,
                                                                  const unsigned *__restrict__ pay_in, unsigned *__restrict__ pay_out,
                                                                  long long n, long long chunk, int shift,
                                                                  const long long *__restrict__ offs, long long nctas) {
  constexpr int ITEMS = SortTile<K>::kItems, TILE = SortTile<K>::kTile;
  __shared__ K s_keys[TILE];
  __shared__ unsigned s_pay[HAS_PAYLOAD ? TILE : 1];
  __shared__ unsigned s_wcnt[kSortWarps][256];      // per-warp digit counts of the tile, then exclusive prefixes over warps
  __shared__ unsigned s_dstart[256];                // first position of a digit in the tile's digit order
  __shared__ long long s_gbase[256];                // running global offset of a digit for this CTA
  __shared__ long long s_gpos[256];                 // global position of tile-order index i with digit d: s_gpos[d] + i
  __shared__ unsigned s_wtot[kSortWarps];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  s_gbase[threadIdx.x] = offs[(long long)threadIdx.x * nctas + blockIdx.x];
  const long long begin = (long long)blockIdx.x * chunk, end = min(begin + chunk, n);
  for (long long t0 = begin; t0 < end; t0 += TILE) {
    const int count = (int)min((long long)TILE, end - t0);
    for (int w = 0; w < kSortWarps; ++w) s_wcnt[w][threadIdx.x] = 0;
    __syncthreads();
    // warp-striped load: item r of lane l is element warp*32*ITEMS + r*32 + l of the tile (memory order = (warp, r, lane))
    K key[ITEMS];
    unsigned pay[HAS_PAYLOAD ? ITEMS : 1], rank[ITEMS];
    const int wbase = (int)warp * 32 * ITEMS;
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
      const int i = wbase + r * 32 + (int)lane;
      key[r] = 0;
      if (HAS_PAYLOAD) pay[r] = 0;
      if (i < count) { key[r] = keys_in[t0 + i]; if (HAS_PAYLOAD) pay[r] = pay_in[t0 + i]; }
    }
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
      const bool valid = wbase + r * 32 + (int)lane < count;
      const unsigned d = digit_of(key[r], shift);
      // lanes of this warp holding the same digit in this round (8 ballots); invalid lanes form their own class
      unsigned peers = __ballot_sync(0xffffffffu, valid);
      if (!valid) peers = ~peers;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const unsigned bal = __ballot_sync(0xffffffffu, (d >> b) & 1u);
        peers &= ((d >> b) & 1u) ? bal : ~bal;
      }
      const int leader = __ffs(peers) - 1;
      unsigned old = 0;
      if (valid && (int)lane == leader) { old = s_wcnt[warp][d]; s_wcnt[warp][d] = old + __popc(peers); }   // one writer per digit per round
      old = __shfl_sync(0xffffffffu, old, leader);
      rank[r] = old + __popc(peers & lt);
      __syncwarp();
    }
    __syncthreads();
    // digit `threadIdx.x`: exclusive prefix over the warps, tile total, then the digit's start in tile order
    unsigned run = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) { const unsigned c = s_wcnt[w][threadIdx.x]; s_wcnt[w][threadIdx.x] = run; run += c; }
    unsigned incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (unsigned)o) incl += y;
    }
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    unsigned wpre = 0;
    for (unsigned w = 0; w < warp; ++w) wpre += s_wtot[w];
    const unsigned dstart = wpre + incl - run;
    s_dstart[threadIdx.x] = dstart;
    s_gpos[threadIdx.x] = s_gbase[threadIdx.x] - (long long)dstart;
    s_gbase[threadIdx.x] += run;
    __syncthreads();
    // permute the tile into digit order (stable) in shared memory
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
      if (wbase + r * 32 + (int)lane < count) {
        const unsigned d = digit_of(key[r], shift);
        const unsigned pos = s_dstart[d] + s_wcnt[warp][d] + rank[r];
        s_keys[pos] = key[r];
        if (HAS_PAYLOAD) s_pay[pos] = pay[r];
      }
    }
    __syncthreads();
    // copy out: neighbours in tile order with the same digit are neighbours in the output
    for (int i = threadIdx.x; i < count; i += kSortBlock) {
      const K k = s_keys[i];
      const long long pos = s_gpos[digit_of(k, shift)] + i;
      keys_out[pos] = k;
      if (HAS_PAYLOAD) pay_out[pos] = s_pay[i];
    }
    __syncthreads();
  }
}

// ---- key transforms ----------------------------------------------------------------------------
// float -> u32 that sorts like the float (NaN-free inputs; -0 sorts before +0)
__device__ __forceinline__ unsigned f32_to_ordered(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_f32(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__global__ void f32_encode_kernel(const float *__restrict__ in, unsigned *__restrict__ out, long long n, int descending) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned k = f32_to_ordered(in[i]);
    out[i] = descending ? ~k : k;
  }
}
__global__ void f32_decode_kernel(const unsigned *__restrict__ in, float *__restrict__ out, long long n, int descending) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned k = in[i];
    out[i] = ordered_to_f32(descending ? ~k : k);
  }
}
__global__ void i32_encode_kernel(const int *__restrict__ in, unsigned *__restrict__ out, long long n, int descending) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned k = (unsigned)in[i] ^ 0x80000000u;
    out[i] = descending ? ~k : k;
  }
}
__global__ void i32_decode_kernel(const unsigned *__restrict__ in, int *__restrict__ out, long long n, int descending) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned k = in[i];
    out[i] = (int)((descending ? ~k : k) ^ 0x80000000u);
  }
}
__global__ void iota_kernel(unsigned *__restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = (unsigned)i;
}
__global__ void gather_f32_kernel(const float *__restrict__ in, const unsigned *__restrict__ idx, float *__restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = in[idx[i]];
}

// ---- small sorts: one CTA, one launch ---------------------------------------------------------------
// The callers of jit_sort_float / jit_sort_pairs sort a handful of groups or a LIMITed result
// (src/warpdb.cpp:370-371,453-455); the radix sort needs 14 launches whatever the size.  Up to 4 096
// elements: encode, stable bitonic sort of (key << 32 | index) in shared memory, decode, payload
// permuted alongside -- the index in the low word makes the network stable, like the bubble sorts it replaces.
constexpr int kSmallSortMax = 4096;
template <int IS_FLOAT>
__global__ void __launch_bounds__(1024) small_sort_kernel(void *__restrict__ keys, unsigned *__restrict__ payload, int n, int m, int descending) {
  extern __shared__ unsigned long long s_comp[];                      // m composites, then n payload words
  unsigned *s_pay = reinterpret_cast<unsigned *>(s_comp + m);
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    unsigned long long c = ~0ull;                                     // padding sorts last
    if (i < n) {
      unsigned k = IS_FLOAT ? f32_to_ordered(reinterpret_cast<const float *>(keys)[i]) : ((unsigned)reinterpret_cast<const int *>(keys)[i] ^ 0x80000000u);
      if (descending) k = ~k;
      c = ((unsigned long long)k << 32) | (unsigned)i;
      if (payload) s_pay[i] = payload[i];
    }
    s_comp[i] = c;
  }
  __syncthreads();
  for (int k = 2; k <= m; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int x = i ^ j;
        if (x > i) {
          const unsigned long long a = s_comp[i], b = s_comp[x];
          if ((a > b) == ((i & k) == 0)) { s_comp[i] = b; s_comp[x] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long c = s_comp[i];
    unsigned k = (unsigned)(c >> 32);
    if (descending) k = ~k;
    if (IS_FLOAT) reinterpret_cast<float *>(keys)[i] = ordered_to_f32(k);
    else reinterpret_cast<int *>(keys)[i] = (int)(k ^ 0x80000000u);
    if (payload) payload[i] = s_pay[(unsigned)c];
  }
}
// returns true when the small path handled the sort
static bool small_sort(cudaStream_t s, void *keys, unsigned *payload, long long n, bool is_float, bool ascending) {
  if (n > opt("sort.small_max", kSmallSortMax) || n > kSmallSortMax) return false;
  int m = 2;
  while (m < n) m <<= 1;
  const size_t smem = (size_t)m * 8 + (payload ? (size_t)n * 4 : 0);
  const unsigned block = (unsigned)std::min(1024, std::max(32, m / 2));
  if (is_float) small_sort_kernel<1><<<1, block, smem, s>>>(keys, payload, (int)n, m, ascending ? 0 : 1);
  else small_sort_kernel<0><<<1, block, smem, s>>>(keys, payload, (int)n, m, ascending ? 0 : 1);
  stats().launches++;
  return true;
}

static unsigned grid_for(Device *d, long long n) {
  return (unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)d->num_sms * 16));
}

// Stable ascending sort of n keys (and optional 32-bit payloads).  keys/pay are sorted in place;
// tmp_keys/tmp_pay are same-sized scratch.  key_bits selects the number of 8-bit passes.
template <class K>
int radix_sort(Device *d, cudaStream_t s, K *keys, K *tmp_keys, unsigned *pay, unsigned *tmp_pay, long long n, int key_bits) {
  if (n <= 1) return 0;
  // every CTA owns a contiguous chunk of whole tiles; ~8 CTAs per SM keep the scan small (256 x nctas counters)
  const long long tile = SortTile<K>::kTile;
  const long long ntiles = (n + tile - 1) / tile;
  const long long nctas = std::max<long long>(1, std::min<long long>(ntiles, (long long)d->num_sms * 8));
  const long long chunk = (ntiles + nctas - 1) / nctas * tile;
  const long long nchunks = (n + chunk - 1) / chunk;
  const size_t hist_bytes = sizeof(unsigned) * 256 * (size_t)nchunks, offs_bytes = sizeof(long long) * 256 * (size_t)nchunks;
  void *hist = nullptr;
  Scratch hist_scratch;
  WDB_CUDA(hist_scratch.alloc(hist_bytes + offs_bytes, s));
  hist = hist_scratch.p;
  unsigned *d_hist = (unsigned *)hist;
  long long *d_offs = (long long *)((char *)hist + hist_bytes);
  K *src = keys, *dst = tmp_keys;
  unsigned *psrc = pay, *pdst = tmp_pay;
  const int passes = key_bits / 8;
  for (int p = 0; p < passes; ++p) {
    const int shift = 8 * p;
    sort_hist_kernel<K><<<(unsigned)nchunks, kSortBlock, 0, s>>>(src, n, chunk, shift, d_hist, nchunks);
    sort_scan_kernel<<<1, 1024, 0, s>>>(d_hist, d_offs, 256 * nchunks);
    if (pay) sort_scatter_kernel<K, true><<<(unsigned)nchunks, kSortBlock, 0, s>>>(src, dst, psrc, pdst, n, chunk, shift, d_offs, nchunks);
    else sort_scatter_kernel<K, false><<<(unsigned)nchunks, kSortBlock, 0, s>>>(src, dst, nullptr, nullptr, n, chunk, shift, d_offs, nchunks);
    stats().launches += 3;
    std::swap(src, dst);
    std::swap(psrc, pdst);
  }
  WDB_CUDA(cudaGetLastError());
  if (src != keys) {  // odd number of passes: result sits in the scratch buffers
    WDB_CUDA(cudaMemcpyAsync(keys, src, sizeof(K) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    if (pay) WDB_CUDA(cudaMemcpyAsync(pay, psrc, sizeof(unsigned) * (size_t)n, cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}
template int radix_sort<unsigned>(Device *, cudaStream_t, unsigned *, unsigned *, unsigned *, unsigned *, long long, int);
template int radix_sort<unsigned long long>(Device *, cudaStream_t, unsigned long long *, unsigned long long *, unsigned *, unsigned *, long long, int);

// sort floats in place (optionally carrying a float payload that is permuted alongside)
int sort_f32(Device *d, cudaStream_t s, float *d_keys, float *d_payload, long long n, bool ascending) {
  if (n <= 1) return 0;
  if (small_sort(s, d_keys, reinterpret_cast<unsigned *>(d_payload), n, true, ascending)) { WDB_CUDA(cudaGetLastError()); return 0; }
  if (d_payload && n >= (1ll << 32)) return fail("ORDER BY with a separate SELECT expression is limited to 2^32 surviving rows (%lld given)", n);
  const size_t nb = sizeof(unsigned) * (size_t)n;
  Scratch scratch;
  WDB_CUDA(scratch.alloc(nb * (d_payload ? 5 : 2), s));
  char *buf = scratch.as<char>();
  unsigned *k = (unsigned *)buf, *kt = (unsigned *)(buf + nb);
  unsigned *p = d_payload ? (unsigned *)(buf + 2 * nb) : nullptr, *pt = d_payload ? (unsigned *)(buf + 3 * nb) : nullptr;
  float *pv = d_payload ? (float *)(buf + 4 * nb) : nullptr;
  const unsigned g = grid_for(d, n);
  f32_encode_kernel<<<g, 256, 0, s>>>(d_keys, k, n, ascending ? 0 : 1);
  if (d_payload) iota_kernel<<<g, 256, 0, s>>>(p, n);
  if (radix_sort<unsigned>(d, s, k, kt, p, pt, n, 32)) return 1;
  f32_decode_kernel<<<g, 256, 0, s>>>(k, d_keys, n, ascending ? 0 : 1);
  if (d_payload) {
    gather_f32_kernel<<<g, 256, 0, s>>>(d_payload, p, pv, n);
    WDB_CUDA(cudaMemcpyAsync(d_payload, pv, nb, cudaMemcpyDeviceToDevice, s));
  }
  stats().launches += d_payload ? 4 : 2;
  WDB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace wdb

using namespace wdb;

extern "C" {

// jit_sort_float (include/jit.hpp:26-27, src/jit.cpp:283-307)
int wdb_sort_float(int device, void *stream, float *d_vals, int64_t count, int ascending) {
  Device *d;
  if (get_device(device, &d)) return 1;
  if (count < 0) return fail("negative count");
  return sort_f32(d, (cudaStream_t)stream, d_vals, nullptr, count, ascending != 0);
}

// jit_sort_pairs (include/jit.hpp:22-23, src/jit.cpp:248-281): stable sort of (key,val) by int key
int wdb_sort_pairs(int device, void *stream, int32_t *d_keys, float *d_vals, int64_t count, int ascending) {
  Device *d;
  if (get_device(device, &d)) return 1;
  if (count < 0) return fail("negative count");
  if (count <= 1) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = count;
  if (small_sort(s, d_keys, reinterpret_cast<unsigned *>(d_vals), n, false, ascending != 0)) { WDB_CUDA(cudaGetLastError()); return 0; }
  const size_t nb = 4 * (size_t)n;
  Scratch scratch;
  WDB_CUDA(scratch.alloc(nb * 3, s));
  char *buf = scratch.as<char>();
  unsigned *k = (unsigned *)buf, *kt = (unsigned *)(buf + nb), *pt = (unsigned *)(buf + 2 * nb);
  const unsigned g = grid_for(d, n);
  i32_encode_kernel<<<g, 256, 0, s>>>(d_keys, k, n, ascending ? 0 : 1);
  // the float values ride along as raw 32-bit payloads
  if (radix_sort<unsigned>(d, s, k, kt, (unsigned *)d_vals, pt, n, 32)) return 1;
  i32_decode_kernel<<<g, 256, 0, s>>>(k, d_keys, n, ascending ? 0 : 1);
  stats().launches += 2;
  WDB_CUDA(cudaGetLastError());
  return 0;
}
}
