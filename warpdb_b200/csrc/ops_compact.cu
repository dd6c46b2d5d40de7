// ops_compact.cu -- host side of the single-pass stable compaction (kernels/compact.cuh).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <mutex>
#include <unordered_map>

#include "core.hpp"

namespace wdb {

// ---- exclusive scan of u32 chunk counts into i64 offsets (two-pass compaction) -----------------
constexpr int kScanBlock = 1024, kScanItems = 16;
__global__ void __launch_bounds__(kScanBlock) scan_local_kernel(const unsigned *__restrict__ in, long long *__restrict__ out,
                                                                long long *__restrict__ block_sums, long long m) {
  __shared__ long long s_warp[kScanBlock / 32];
  const long long base = ((long long)blockIdx.x * kScanBlock + threadIdx.x) * kScanItems;
  unsigned v[kScanItems];
  long long sum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) { v[i] = base + i < m ? in[base + i] : 0u; sum += v[i]; }
  long long incl = sum;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += t; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    long long w = s_warp[lane];
    long long wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= (unsigned)o) wi += t; }
    s_warp[lane] = wi - w;
    if (lane == 31) block_sums[blockIdx.x] = wi;
  }
  __syncthreads();
  long long run = s_warp[warp] + incl - sum;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) { if (base + i < m) out[base + i] = run; run += v[i]; }
}
__global__ void __launch_bounds__(1024) scan_sums_kernel(long long *__restrict__ block_sums, long long nb, long long *__restrict__ total) {
  // single block: exclusive scan of nb values in place.  Every thread owns a contiguous run; the 1024 run totals
  // are scanned with two levels of warp shuffles (a serial loop of thread 0 over them cost 30 us).
  __shared__ long long s_warp[32];
  const long long per = (nb + 1023) / 1024;
  const long long b = (long long)threadIdx.x * per, e = min(b + per, nb);
  long long sum = 0;
  for (long long i = b; i < e; ++i) sum += block_sums[i];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  long long incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += t; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const long long w = s_warp[lane];
    long long wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= (unsigned)o) wi += t; }
    s_warp[lane] = wi - w;
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  long long run = s_warp[warp] + incl - sum;
  for (long long i = b; i < e; ++i) { const long long t = block_sums[i]; block_sums[i] = run; run += t; }
}
__global__ void __launch_bounds__(kScanBlock) scan_add_kernel(long long *__restrict__ out, const long long *__restrict__ block_sums, long long m) {
  const long long base = ((long long)blockIdx.x * kScanBlock + threadIdx.x) * kScanItems;
  const long long add = block_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < m) out[base + i] += add;
}

// Selectivity feedback: one pinned 8-byte slot (the survivor count; ~0 = nothing yet) per (query shape, table size,
// device), written by an asynchronous device-to-host copy at the end of every call and read by the next call of that shape.
static std::mutex g_fb_mu;
static std::unordered_map<size_t, unsigned long long *> g_fb_slots;
static unsigned long long *g_fb_pool = nullptr;
static size_t g_fb_used = 0;
constexpr size_t kFbSlots = 4096;
static unsigned long long *sel_feedback_slot(Device *d, size_t key) {
  key ^= (size_t)d->id * 0x100000001b3ull;
  std::lock_guard<std::mutex> l(g_fb_mu);
  auto it = g_fb_slots.find(key);
  if (it != g_fb_slots.end()) return it->second;
  if (!g_fb_pool) {
    if (cudaHostAlloc((void **)&g_fb_pool, kFbSlots * 8, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    memset(g_fb_pool, 0xff, kFbSlots * 8);
  }
  if (g_fb_used == kFbSlots) { g_fb_slots.clear(); g_fb_used = 0; memset(g_fb_pool, 0xff, kFbSlots * 8); }
  unsigned long long *slot = g_fb_pool + g_fb_used++;
  g_fb_slots[key] = slot;
  return slot;
}

struct CompactPlan { GenSpec spec; int block, unroll, vec, variant, stage_cap, stage_m; int64_t tile_rows; bool two; size_t smem; };

static int plan_compact(const wdb_col_t *cols, int ncols, const char *expr, const char *expr2, const char *cond,
                        bool check_alignment, int thresh, CompactPlan *p, bool prune = false, int force_variant = -1, bool auto_sel = false,
                        int stage_cap_hint = 0) {
  GenSpec &spec = p->spec;
  spec.kind = "compact";
  if (!cond || !*cond) cond = "true";
  const bool two = expr2 && *expr2;
  p->two = two;
  spec.used = find_used_columns(cols, ncols, {expr, two ? expr2 : "", cond});
  for (const auto &u : spec.used)
    if (dtype_size(u.dtype) == 0) return fail("column %s has a non-numeric type and cannot be read on the GPU", u.name.c_str());
  int block = (int)opt("compact.block", 256);
  const int vec = (int)opt("compact.vec", 8);
  int unroll = (int)opt("compact.unroll", 4);
  int variant = force_variant >= 0 ? force_variant : (int)opt("compact.variant", two ? 2 : 3);   // defaults from profiles/r01_sweep_compact_*.jsonl   // 0 ticket + register loads, 1 TMA bulk ring, 2 two-pass count/scatter
  if (vec != 4 && vec != 8) return fail("compact.vec must be 4 or 8");
  if (block < 32 || block > 1024 || (block & 31)) return fail("compact.block must be a multiple of 32 in [32,1024]");
  if (unroll < 1 || unroll > 8) return fail("compact.unroll must be in [1,8]");
  const bool aligned = !check_alignment || all_aligned(spec.used, cols, nullptr, (size_t)vec * 4);
  if (!aligned && variant == 1) variant = 2;       // bulk copies need 16-byte aligned columns
  // slot of a chunk's parked survivors: the smaller it is, the closer the slots lie in DRAM for the gather pass
  // (the optimizer sizes it from the selectivity it has seen: expected survivors per chunk + 6 sigma, 32 / 64 / 128;
  // a chunk that overflows its slot anyway is recomputed from the input by the gather pass)
  const int stage_cap = (int)opt("compact.stage_cap", stage_cap_hint > 0 ? stage_cap_hint : 128);
  // chunks per warp: 1 is fastest when the kernels run (profiles/r02_sweep_compact_staged.jsonl); the device-selected
  // twin pipelines use 8 so that the pipeline that is not needed costs few CTA launches
  const int stage_m = (int)std::max<int64_t>(1, std::min<int64_t>(32, opt("compact.stage_m", auto_sel ? 8 : 1)));
  if (variant == 5 && (two || stage_cap < 32 || stage_cap > 1024 || (stage_cap & 31))) variant = two ? 2 : 3;
  if (variant == 5 && opt("compact.block", -1) < 0) block = 128;
  if (prune) {                                     // zone-map pruning is wired into the two-pass kernels
    if (variant != 5) variant = 2;
    while (unroll & (unroll - 1)) --unroll;        // a warp's chunk must not straddle zones: power-of-two chunks divide the (power-of-two) zone size
  }
  size_t row_bytes = 0;
  for (const auto &u : spec.used) row_bytes += dtype_size(u.dtype);
  if (variant != 1)
    while (unroll > 1 && (int64_t)block * vec * unroll * 4 * (two ? 2 : 1) > 46 * 1024) unroll /= 2;   // static shared memory
  else
    while (unroll > 1 && 128 + (size_t)block * vec * unroll * (2 * row_bytes + 4 * (two ? 2 : 1)) > 200 * 1024) unroll /= 2;
  p->block = block; p->unroll = unroll; p->vec = vec; p->variant = variant; p->stage_cap = stage_cap; p->stage_m = stage_m;
  p->tile_rows = (int64_t)block * vec * unroll;
  p->smem = variant == 1 ? 128 + (size_t)p->tile_rows * (2 * row_bytes + 4 * (two ? 2 : 1)) : 0;
  const int sp_stages = (int)opt("compact.sp_stages", 2);
  if (variant == 4) {   // single pass, pipelined look-back: S staging buffers per worker warp in dynamic shared memory
    if (sp_stages < 2 || sp_stages > 7) return fail("compact.sp_stages must be in [2,7]");
    const int nw = block / 32;
    p->smem = (size_t)((12 * sp_stages * nw + 15) / 16 * 16) + (size_t)sp_stages * p->tile_rows * 4 * (two ? 2 : 1);
    if (p->smem > 227 * 1024) return fail("compact.sp_stages x compact.block needs %zu B of shared memory", p->smem);
  }
  if (variant != 1 && variant != 4 && p->tile_rows * 4 * (two ? 2 : 1) > 46 * 1024) return fail("compact tile of %lld rows does not fit static shared memory", (long long)p->tile_rows);
  spec.defines = {{"WDB_VEC", vec}, {"WDB_ALIGNED", aligned ? 1 : 0}, {"WDB_LD_HINT", opt("compact.ld_hint", 0)},
                  {"WDB_ST_HINT", 0}, {"WDB_BLOCK", block}, {"WDB_UNROLL", unroll}, {"WDB_NOUT", two ? 2 : 1},
                  {"WDB_THRESH", two ? thresh : 0}, {"WDB_MIN_CTAS", opt("compact.min_ctas", variant == 3 ? 4 : (variant == 4 ? 3 : 1))},
                  {"WDB_LB", opt("compact.lookback", variant == 1 ? 4 : 1)}, {"WDB_BULK", variant == 1 ? 1 : 0}, {"WDB_TWOPASS", (variant == 2 || variant == 5) ? 1 : 0},
                  {"WDB_STAGE_CAP", variant == 5 ? stage_cap : 0}, {"WDB_AUTO_SEL", auto_sel ? 1 : 0},
                  {"WDB_STAGE_PERMILLE", opt("compact.stage_max_sel_permille", 45)}, {"WDB_STAGE_M", stage_m}, {"WDB_PRUNE", prune ? 1 : 0},
                  {"WDB_L2PASS", (variant == 3 || variant == 4) ? 1 : 0}, {"WDB_SP_STAGES", sp_stages}, {"WDB_SP_SCAN", opt("compact.sp_scan", 0)}, {"WDB_SLAB_M", opt("compact.slab_m", 4)}, {"WDB_L2_HINTS", opt("compact.l2_hints", 1)}, {"WDB_PF_NEXT", opt("compact.pf_next", 1)}};
  if (variant == 1) spec.defines.push_back({"WDB_TILE", p->tile_rows});
  spec.fns.push_back({"expr", "float", expr});
  if (two) spec.fns.push_back({"expr2", "float", expr2});
  spec.fns.push_back({"cond", "bool", cond});
  spec.bodies = {k_src_compact};
  return 0;
}

int gen_compact_source(const wdb_col_t *cols, int ncols, const char *expr, const char *expr2, const char *cond,
                       bool assume_aligned, std::string *src) {
  CompactPlan p;
  if (plan_compact(cols, ncols, expr, expr2, cond, !assume_aligned, 0, &p)) return 1;
  *src = gen_source(p.spec);
  return 0;
}

// launches the kernels of plan p; *count_ptr points at the survivor count inside `scratch` (kept by the caller
// until it has copied the count out).  sel != nullptr: the plan was built with auto_sel and its kernels
// consult the sampled selectivity at `sel` to decide whether they or the sibling pipeline do the work.
static int launch_compact_plan(Device *d, cudaStream_t stream, const CompactPlan &p, const wdb_col_t *cols, float *d_out, float *d_out2,
                               int64_t n, float tau, int64_t out_cap, const unsigned char *zmask, int zshift,
                               const unsigned long long *sel, Scratch &scratch, long long **count_ptr) {
  const GenSpec &spec = p.spec;
  const int block = p.block;
  const int64_t tile_rows = p.tile_rows;
  std::vector<const void *> ptrs;
  for (const auto &u : spec.used) ptrs.push_back(cols[u.table_index].dptr);
  if (ptrs.empty()) ptrs.push_back(nullptr);
  const void *ptrs_data = ptrs.data();   // kernels take the pointer block by value
  (void)ptrs_data;
  if (p.variant == 2 || p.variant == 5) {  // two streaming passes: count per warp chunk, scan, scatter (5: survivors parked in pass 1)
    const bool staged = p.variant == 5;
    const std::string src = gen_source(spec);
    Kernel kc, ks;
    if (get_kernel(d, src, "wdb_compact.cu", staged ? "wdb_count_stage" : "wdb_count", &kc) ||
        get_kernel(d, src, "wdb_compact.cu", staged ? "wdb_gather_stage" : "wdb_scatter", &ks))
      return 1;
    const int nwarps = block / 32;
    const int64_t chunk_rows = tile_rows / nwarps;
    const int64_t nchunks = (n + chunk_rows - 1) / chunk_rows;
    const int64_t nsb = (nchunks + kScanBlock * kScanItems - 1) / (kScanBlock * kScanItems);
    const int64_t stage_cap = staged ? p.stage_cap : 0;
    const int64_t ngroups = (nchunks + 31) / 32;   // staged: one gather warp serves 32 chunks; only the group totals are scanned
    // scratch: [0,8) total, [64, ...) offsets i64[nchunks] (staged: i64[ngroups] group totals -> offsets), block sums i64[nsb],
    // counts u32[nchunks], (staged) slots f32[nchunks][cap]
    const size_t off_bytes = 8 * (size_t)std::max<int64_t>(staged ? ngroups : nchunks, 1), sum_bytes = 8 * (size_t)std::max<int64_t>(nsb, 1);
    const size_t cnt_bytes = (4 * (size_t)std::max<int64_t>(nchunks, 1) + 255) & ~(size_t)255;
    const size_t need = 64 + off_bytes + sum_bytes + cnt_bytes + (size_t)nchunks * (size_t)stage_cap * 4;
    // stream-ordered scratch: concurrent calls on different streams of one device do not share it
    WDB_CUDA(scratch.alloc(need, stream));
    char *sc = scratch.as<char>();
    long long *d_total = (long long *)sc;
    long long *d_offs = (long long *)(sc + 64);
    long long *d_sums = (long long *)(sc + 64 + off_bytes);
    unsigned *d_counts = (unsigned *)(sc + 64 + off_bytes + sum_bytes);
    float *d_slots = (float *)(sc + 64 + off_bytes + sum_bytes + cnt_bytes);
    *count_ptr = d_total;
    WDB_CUDA(cudaMemsetAsync(sc, 0, staged ? 64 + off_bytes : 64, stream));
    if (nchunks > 0) {
      long long nn = n, nc = nchunks, cap = out_cap;
      const unsigned grid = (unsigned)((nchunks + nwarps - 1) / nwarps);
      {
        const int64_t per_cta = (int64_t)nwarps * (staged ? p.stage_m : 1);
        const unsigned cgrid = (unsigned)((nchunks + per_cta - 1) / per_cta);
        std::vector<void *> args = {ptrs.data(), &nn, &d_counts};
        if (staged) { args.push_back(&d_offs); args.push_back(&d_slots); }
        args.push_back(&nc);
        args.push_back(&tau);
        if (sel) args.push_back(&sel);
        args.push_back(&zmask);
        args.push_back(&zshift);
        if (launch(kc, cgrid, block, 0, stream, args.data())) return 1;
      }
      if (staged) {
        scan_sums_kernel<<<1, 1024, 0, stream>>>(d_offs, ngroups, d_total);   // group totals -> exclusive group offsets, in place
        stats().launches++;
        WDB_CUDA(cudaGetLastError());
        const unsigned ggrid = (unsigned)((ngroups + nwarps - 1) / nwarps);
        std::vector<void *> args = {ptrs.data(), &d_out, &nn, &d_counts, &d_offs, &d_slots, &nc, &tau, &cap};
        if (sel) args.push_back(&sel);
        args.push_back(&zmask);
        args.push_back(&zshift);
        if (launch(ks, ggrid, block, 0, stream, args.data())) return 1;
      } else {
        scan_local_kernel<<<(unsigned)nsb, kScanBlock, 0, stream>>>(d_counts, d_offs, d_sums, nchunks);
        scan_sums_kernel<<<1, 1024, 0, stream>>>(d_sums, nsb, d_total);
        scan_add_kernel<<<(unsigned)nsb, kScanBlock, 0, stream>>>(d_offs, d_sums, nchunks);
        stats().launches += 3;
        WDB_CUDA(cudaGetLastError());
        void *args[] = {ptrs.data(), &d_out, &d_out2, &nn, &d_offs, &nc, &tau, &cap, &zmask, &zshift};
        if (launch(ks, grid, block, 0, stream, args)) return 1;
      }
    }
    return 0;
  }
  Kernel k;
  const bool bulk = p.variant == 1, l2pass = p.variant == 3, sp = p.variant == 4;
  if (get_kernel(d, gen_source(spec), "wdb_compact.cu", bulk ? "wdb_compact_bulk" : (l2pass ? "wdb_compact_l2" : (sp ? "wdb_compact_sp" : "wdb_compact")), &k)) return 1;

  // variant 3 works on slabs of slab_m chunks per warp; the status words are per slab
  const int64_t slab_rows = tile_rows * std::max<int64_t>(1, opt("compact.slab_m", 4));
  const int64_t ntiles = l2pass ? (n + slab_rows - 1) / slab_rows : (n + tile_rows - 1) / tile_rows;
  // scratch: [0,8) survivor count, [8,12) ticket, [64, 64+8*ntiles) tile status words (+ one word per round of
  // the single-pass variant: at least one slab per CTA and round, so ntiles + 2 words always suffice)
  const size_t need = 64 + (size_t)std::max<int64_t>(ntiles, 1) * 8 + (sp ? (size_t)(ntiles + 2) * 8 : 0);
  WDB_CUDA(scratch.alloc(need, stream));
  char *sc = scratch.as<char>();
  WDB_CUDA(cudaMemsetAsync(sc, 0, need, stream));
  long long *d_cnt = (long long *)sc;
  *count_ptr = d_cnt;
  unsigned *d_ticket = (unsigned *)(sc + 8);
  unsigned long long *d_status = (unsigned long long *)(sc + 64);
  if (ntiles > 0) {
    if (p.smem > 48 * 1024) WDB_CUDA(cudaFuncSetAttribute((const void *)k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    int nb = 0;
    WDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void *)k.fn, sp ? block + 32 : block, p.smem));
    if (nb < 1) return fail("compaction kernel does not fit on an SM (%zu bytes of shared memory)", p.smem);
    // the bulk variant assigns tiles statically: its grid must not exceed what is co-resident
    int64_t per_sm = std::min<int64_t>(nb, opt("compact.ctas_per_sm", 8));
    unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)d->num_sms * per_sm);
    long long nn = n, nt = ntiles;
    long long cap = out_cap;
    if (bulk) {
      void *args[] = {ptrs.data(), &d_out, &d_out2, &nn, &d_status, &d_cnt, &nt, &tau, &cap};
      if (launch(k, grid, block, p.smem, stream, args)) return 1;
    } else if (sp) {   // static round-robin slabs: the grid (<= SMs x resident CTAs, above) is co-resident by construction
      const int64_t chunk_rows = tile_rows / (block / 32);
      long long nchunks = (n + chunk_rows - 1) / chunk_rows;
      unsigned long long *d_rbase = d_status + std::max<int64_t>(ntiles, 1);
      void *args[] = {ptrs.data(), &d_out, &d_out2, &nn, &d_status, &d_rbase, &d_ticket, &d_cnt, &nt, &nchunks, &tau, &cap};
      if (launch(k, grid, block + 32, p.smem, stream, args)) return 1;
    } else if (l2pass) {
      const int64_t chunk_rows = tile_rows / (block / 32);
      long long nchunks = (n + chunk_rows - 1) / chunk_rows;
      void *args[] = {ptrs.data(), &d_out, &d_out2, &nn, &d_status, &d_ticket, &d_cnt, &nt, &nchunks, &tau, &cap, &sel};   // sel: auto_sel instantiations only
      if (launch(k, grid, block, 0, stream, args)) return 1;
    } else {
      void *args[] = {ptrs.data(), &d_out, &d_out2, &nn, &d_status, &d_ticket, &d_cnt, &nt, &tau, &cap};
      if (launch(k, grid, block, 0, stream, args)) return 1;
    }
  }
  return 0;
}

__global__ void add_counts_kernel(const long long *a, const long long *b, long long *out) { *out = *a + *b; }

int run_compact_ex(Device *d, cudaStream_t stream, const wdb_col_t *cols, int ncols, const char *expr, const char *expr2,
                   const char *cond, float *d_out, float *d_out2, int64_t n, int64_t *d_count, int64_t *h_count,
                   int thresh, float tau, int64_t out_cap, const unsigned char *zmask, int zshift) {
  const bool two_out = expr2 && *expr2;
  auto check_zones = [&](const CompactPlan &p) {
    if (zmask && ((1ll << zshift) < p.tile_rows / (p.block / 32) || (1ll << zshift) % (p.tile_rows / (p.block / 32)) != 0))
      return fail("zone size %lld is not a multiple of the compaction chunk (%lld rows)", 1ll << zshift, (long long)(p.tile_rows / (p.block / 32)));
    return 0;
  };
  long long *cnt = nullptr;
  Scratch sa, sb, ssel;
  const int64_t auto_mode = (!two_out && !zmask && opt("compact.variant", -1) < 0 && n >= opt("compact.auto_min_rows", 1 << 27)) ? opt("compact.auto", 1) : 0;
  int force = -1;
  unsigned long long *fb_slot = nullptr;
  int cap_hint = 0;
  if (auto_mode == 1) {
    // Optimizer (default): selective filters (selectivity below compact.stage_max_sel_permille) run the staged
    // two-pass kernels, everything else the L2-parked slabs.  The statistic is the query's own result: every call
    // sends its survivor count to a pinned host slot with an asynchronous 8-byte copy, and the NEXT call of the
    // same query shape on the same table size reads it.  No host synchronisation, no sampling pass, no
    // speculative launches; the decision lags one call behind the data, and a wrong one only costs speed.
    CompactPlan p5;
    if (plan_compact(cols, ncols, expr, expr2, cond, true, thresh, &p5, false, 5)) return 1;
    if (p5.variant == 5) {
      fb_slot = sel_feedback_slot(d, std::hash<std::string>{}(gen_source(p5.spec)) ^ (std::hash<long long>{}((long long)n) * 0x9E3779B97F4A7C15ull));
      if (fb_slot) {
        const unsigned long long survivors = *(volatile unsigned long long *)fb_slot;
        if (survivors != ~0ull && survivors * 1000ull <= (unsigned long long)n * (unsigned long long)opt("compact.stage_max_sel_permille", 45)) {
          force = 5;
          const double per_chunk = (double)survivors / (double)n * (double)(p5.tile_rows / (p5.block / 32));
          const double need = per_chunk + 6.0 * std::sqrt(per_chunk) + 2.0;   // mean + 6 sigma of the (binomial) survivors per chunk
          cap_hint = need <= 32.0 ? 32 : (need <= 64.0 ? 64 : 128);
        }
      }
    }
  }
  if (auto_mode == 2) {
    // Device-side selection (compact.auto = 2): the sample stays on the device and BOTH pipelines are launched; the
    // kernels of the one that is not needed return at once.  Always follows the data of this very call, but the
    // idle pipeline still costs ~0.08 ms of launches (measured), more than the staged kernels save above ~3 %.
    CompactPlan p5, p3;
    if (plan_compact(cols, ncols, expr, expr2, cond, true, thresh, &p5, false, 5, true)) return 1;
    if (plan_compact(cols, ncols, expr, expr2, cond, true, thresh, &p3, false, 3, true)) return 1;
    if (p5.variant == 5 && p3.variant == 3) {
      Kernel ksamp;
      if (get_kernel(d, gen_source(p5.spec), "wdb_compact.cu", "wdb_sample_count", &ksamp)) return 1;
      const int nwarps = p5.block / 32;
      const int64_t chunk_rows = p5.tile_rows / nwarps, nchunks = n / chunk_rows;   // whole chunks only
      const int64_t stride = std::max<int64_t>(1, nchunks / 256), nsamp = (nchunks + stride - 1) / stride;
      WDB_CUDA(ssel.alloc(32, stream));
      WDB_CUDA(cudaMemsetAsync(ssel.p, 0, 32, stream));
      unsigned long long *sel = ssel.as<unsigned long long>();
      {
        std::vector<const void *> ptrs;
        for (const auto &u : p5.spec.used) ptrs.push_back(cols[u.table_index].dptr);
        if (ptrs.empty()) ptrs.push_back(nullptr);
        long long nn = n, nc = nchunks, sd = stride;
        void *args[] = {ptrs.data(), &nn, &sel, &nc, &sd, &tau};
        if (launch(ksamp, (unsigned)((nsamp + nwarps - 1) / nwarps), p5.block, 0, stream, args)) return 1;
      }
      long long *c5 = nullptr, *c3 = nullptr;
      if (launch_compact_plan(d, stream, p5, cols, d_out, d_out2, n, tau, out_cap, nullptr, 0, sel, sa, &c5)) return 1;
      if (launch_compact_plan(d, stream, p3, cols, d_out, d_out2, n, tau, out_cap, nullptr, 0, sel, sb, &c3)) return 1;
      cnt = (long long *)(sel + 2);
      add_counts_kernel<<<1, 1, 0, stream>>>(c5, c3, cnt);
      stats().launches++;
      WDB_CUDA(cudaGetLastError());
    }
  }
  if (!cnt) {
    CompactPlan p;
    if (plan_compact(cols, ncols, expr, expr2, cond, true, thresh, &p, zmask != nullptr, force, false, cap_hint)) return 1;
    if (check_zones(p)) return 1;
    if (!two_out) wdb_set_option("compact.last_variant", p.variant);   // introspection (bench.py labels its roofline with it)
    if (launch_compact_plan(d, stream, p, cols, d_out, d_out2, n, tau, out_cap, zmask, zshift, nullptr, sa, &cnt)) return 1;
  }
  if (d_count) WDB_CUDA(cudaMemcpyAsync(d_count, cnt, 8, cudaMemcpyDeviceToDevice, stream));
  if (h_count) WDB_CUDA(cudaMemcpyAsync(h_count, cnt, 8, cudaMemcpyDeviceToHost, stream));
  if (fb_slot) WDB_CUDA(cudaMemcpyAsync(fb_slot, cnt, 8, cudaMemcpyDeviceToHost, stream));   // feedback for the next call of this query shape
  sa.release();
  sb.release();
  ssel.release();
  if (h_count) WDB_CUDA(cudaStreamSynchronize(stream));
  return 0;
}

int run_compact(Device *d, cudaStream_t stream, const wdb_col_t *cols, int ncols, const char *expr, const char *expr2,
                const char *cond, float *d_out, float *d_out2, int64_t n, int64_t *d_count, int64_t *h_count) {
  return run_compact_ex(d, stream, cols, ncols, expr, expr2, cond, d_out, d_out2, n, d_count, h_count, 0, 0.0f, n, nullptr, 0);
}

}  // namespace wdb
