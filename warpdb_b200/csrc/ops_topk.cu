// ops_topk.cu -- ORDER BY key [ASC|DESC] [LIMIT k [OFFSET o]] (kernels/topk.cuh + compact.cuh).
//
// Replaces jit_sort_float + host truncate (src/jit.cpp:283-307, src/warpdb.cpp:453-455,483-495) and
// the keyed host sort of src/warpdb.cpp:470-476.  Results equal a stable full sort of the surviving
// rows followed by OFFSET/LIMIT:
//   k+o <= 16 : one streaming pass, per-thread register top-k + warp-shuffle / shared-memory merge
//   larger    : best key per tile -> threshold tau (the (k+o)-th best tile extreme bounds the
//               (k+o)-th best row) -> fused filter+compaction of (value,key) pairs passing tau in
//               row order -> stable radix sort of the candidates -> slice
//   no LIMIT  : compaction of all surviving (value,key) pairs + stable radix sort
#include <algorithm>
#include <cmath>

#include "core.hpp"

namespace wdb {
int run_compact_ex(Device *d, cudaStream_t stream, const wdb_col_t *cols, int ncols, const char *expr, const char *expr2,
                   const char *cond, float *d_out, float *d_out2, int64_t n, int64_t *d_count, int64_t *h_count,
                   int thresh, float tau, int64_t out_cap, const unsigned char *zmask = nullptr, int zshift = 0);
int sort_f32(Device *d, cudaStream_t s, float *d_keys, float *d_payload, long long n, bool ascending);
std::string order_key(const char *key_expr, bool desc);

struct TopkPlan { GenSpec spec; int block, unroll, vec, K; };

static int plan_topk(const wdb_col_t *cols, int ncols, const char *key, const char *val, const char *cond, bool desc, int K,
                     bool check_alignment, TopkPlan *p) {
  const bool has_cond = cond && *cond;
  GenSpec &spec = p->spec;
  spec.kind = "topk";
  spec.used = find_used_columns(cols, ncols, {key, val, has_cond ? cond : ""});
  for (const auto &u : spec.used)
    if (dtype_size(u.dtype) == 0) return fail("column %s has a non-numeric type and cannot be read on the GPU", u.name.c_str());
  p->block = (int)opt("topk.block", 512);   // profiles/r01_sweep_topk_2e9.jsonl
  p->unroll = (int)opt("topk.unroll", 2);
  p->vec = (int)opt("topk.vec", 8);
  p->K = K;
  if (p->vec != 4 && p->vec != 8) return fail("topk.vec must be 4 or 8");
  const bool aligned = !check_alignment || all_aligned(spec.used, cols, nullptr, (size_t)p->vec * 4);
  spec.defines = {{"WDB_VEC", p->vec}, {"WDB_ALIGNED", aligned ? 1 : 0}, {"WDB_LD_HINT", opt("topk.ld_hint", 0)}, {"WDB_ST_HINT", 0},
                  {"WDB_BLOCK", p->block}, {"WDB_UNROLL", p->unroll}, {"WDB_K", K}, {"WDB_DESC", desc ? 1 : 0},
                  {"WDB_HAS_COND", has_cond ? 1 : 0}, {"WDB_FUSED_TAIL", opt("topk.fused", 1) ? 1 : 0}};
  spec.fns.push_back({"key", "float", key});
  spec.fns.push_back({"val", "float", val});
  if (has_cond) spec.fns.push_back({"cond", "bool", cond});
  spec.bodies = {k_src_topk};
  return 0;
}

int gen_topk_source(const wdb_col_t *cols, int ncols, const char *key, const char *val, const char *cond, int desc, std::string *src) {
  TopkPlan p;
  if (plan_topk(cols, ncols, key, val && *val ? val : key, cond, desc != 0, 5, false, &p)) return 1;
  *src = gen_source(p.spec);
  return 0;
}

static std::vector<const void *> col_ptrs(const GenSpec &spec, const wdb_col_t *cols) {
  std::vector<const void *> ptrs;
  for (const auto &u : spec.used) ptrs.push_back(cols[u.table_index].dptr);
  if (ptrs.empty()) ptrs.push_back(nullptr);
  return ptrs;
}

// Pre-pass of the register top-k: the same kernel over the first 2^20 rows with a handful of CTAs.  The K-th best key
// of that sample is a valid threshold for the whole column (K rows beat it already), so the main pass starts with a
// sharp test instead of warming up ~150 K cold per-thread lists.  `slot` (>= 12 K + 64 bytes of the call's scratch)
// receives the sample's winners; *tau0 points at its K-th key.  Skipped for small inputs.
static int topk_sample_threshold(Device *d, cudaStream_t s, const Kernel &scan, const TopkPlan &p, std::vector<const void *> &ptrs, int64_t n,
                                 int64_t row_base, int K, char *slot, float *cand_k, long long *cand_r, const float **tau0) {
  *tau0 = nullptr;
  const int64_t sample = opt("topk.sample_rows", 1 << 20);
  if (sample <= 0 || n < opt("topk.sample_min_rows", 1 << 26)) return 0;
  slot = (char *)(((uintptr_t)slot + 15) & ~(uintptr_t)15);
  long long *s_r = (long long *)slot;
  long long *s_cnt = s_r + K;
  unsigned *s_done = (unsigned *)(s_cnt + 1);
  float *s_k = (float *)(s_cnt + 2);
  long long nn = std::min<int64_t>(sample, n), rb = row_base;
  const int64_t tile_rows = (int64_t)p.block * p.unroll * p.vec;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((nn + tile_rows - 1) / tile_rows, d->num_sms));
  int zero = 0;
  float *no_f = nullptr;
  const float *no_tau = nullptr;
  WDB_CUDA(cudaMemsetAsync(s_done, 0, 4, s));
  // the candidates of this pre-pass land in the head of the main pass's candidate arrays, which the main pass rewrites
  void *args[] = {ptrs.data(), &nn, &rb, &cand_k, &cand_r, &no_tau, &s_done, &s_k, &s_r, &zero, &no_f, &no_f, &s_cnt};
  if (launch(scan, grid, p.block, 0, s, args)) return 1;
  *tau0 = s_k + (K - 1);
  return 0;
}

static int topk_small(Device *d, cudaStream_t s, const wdb_col_t *cols, int ncols, const char *key, const char *val, const char *cond,
                      bool desc, int K, int offset, int64_t n, float *d_out_vals, float *d_out_keys, int64_t *h_n) {
  TopkPlan p;
  if (plan_topk(cols, ncols, key, val, cond, desc, K, true, &p)) return 1;
  const std::string src = gen_source(p.spec);
  Kernel scan, fin, emit;
  if (get_kernel(d, src, "wdb_topk.cu", "wdb_topk_scan", &scan) || get_kernel(d, src, "wdb_topk.cu", "wdb_topk_final", &fin) ||
      get_kernel(d, src, "wdb_topk.cu", "wdb_topk_emit", &emit))
    return 1;
  int nb = 0;
  WDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void *)scan.fn, p.block, 0));
  nb = std::max(1, std::min<int>(nb, (int)opt("topk.ctas_per_sm", 8)));
  const int64_t tile_rows = (int64_t)p.block * p.unroll * p.vec;
  const int64_t ntiles = std::max<int64_t>(1, (n + tile_rows - 1) / tile_rows);
  // a few waves of CTAs instead of one persistent wave: the hardware scheduler evens out the CTAs' finishing times
  const unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)d->num_sms * nb * std::max<int64_t>(1, opt("topk.waves", 1)));
  // scratch: candidates of every CTA, the K winners, the count, the done-counter of the fused tail, the pre-pass's winners
  const size_t cand = (size_t)grid * K;
  const size_t bytes = cand * 12 + (size_t)K * 12 + 64 + (size_t)K * 12 + 64;
  Scratch scratch;
  WDB_CUDA(scratch.alloc(bytes, s));
  char *buf = scratch.as<char>();
  long long *cand_r = (long long *)buf;
  long long *best_r = cand_r + cand;
  long long *d_cnt = best_r + K;
  unsigned *d_done = (unsigned *)(d_cnt + 1);
  float *cand_k = (float *)(d_cnt + 2);
  float *best_k = cand_k + cand;
  auto ptrs = col_ptrs(p.spec, cols);
  long long nn = n, rb = 0, m = (long long)cand;
  const float *tau0 = nullptr;
  if (opt("topk.fused", 1)) {   // one launch: the last CTA to finish selects among all candidates and evaluates the SELECT expression
    if (topk_sample_threshold(d, s, scan, p, ptrs, n, 0, K, (char *)(best_k + K), cand_k, cand_r, &tau0)) return 1;
    WDB_CUDA(cudaMemsetAsync(d_done, 0, 4, s));
    void *args[] = {ptrs.data(), &nn, &rb, &cand_k, &cand_r, &tau0, &d_done, &best_k, &best_r, &offset, &d_out_vals, &d_out_keys, &d_cnt};
    if (launch(scan, grid, p.block, 0, s, args)) return 1;
  } else {
    {
      void *args[] = {ptrs.data(), &nn, &rb, &cand_k, &cand_r, &tau0};
      if (launch(scan, grid, p.block, 0, s, args)) return 1;
    }
    {
      void *args[] = {&cand_k, &cand_r, &m, &best_k, &best_r};
      if (launch(fin, 1, p.block, 0, s, args)) return 1;
    }
    {
      void *args[] = {ptrs.data(), &rb, &best_k, &best_r, &offset, &d_out_vals, &d_out_keys, &d_cnt};
      if (launch(emit, 1, 32, 0, s, args)) return 1;
    }
  }
  long long cnt = 0;
  if (h_n) {
    WDB_CUDA(cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, s));
    WDB_CUDA(cudaStreamSynchronize(s));
    *h_n = cnt;
  }
  return 0;
}

// NaN keys order after every number, whatever the direction and the path (prelude.cuh: wdb_nanlast_*)
std::string order_key(const char *key_expr, bool desc) { return std::string(desc ? "wdb_nanlast_d(" : "wdb_nanlast_a(") + key_expr + ")"; }

// Local phase of a sharded ORDER BY ... LIMIT (ops_comm.cu): the K = k+offset best (key, row) pairs of
// this shard and the SELECT value at each of them, written as [K keys f32 | K vals f32 | K rows i64]
// at `cand` (this rank's slot of the all-gather buffer).  Rows are global ids (row_base + local row);
// empty entries carry WDB_ROW_NONE.  Asynchronous on `s`.
int topk_candidates(Device *d, cudaStream_t s, const wdb_col_t *cols, int ncols, const char *key, const char *val, const char *cond,
                    bool desc, int K, int64_t n, int64_t row_base, char *cand) {
  TopkPlan p;
  if (plan_topk(cols, ncols, key, val, cond, desc, K, true, &p)) return 1;
  const std::string src = gen_source(p.spec);
  Kernel scan, fin, emit;
  if (get_kernel(d, src, "wdb_topk.cu", "wdb_topk_scan", &scan) || get_kernel(d, src, "wdb_topk.cu", "wdb_topk_final", &fin) ||
      get_kernel(d, src, "wdb_topk.cu", "wdb_topk_emit", &emit))
    return 1;
  int nb = 0;
  WDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void *)scan.fn, p.block, 0));
  nb = std::max(1, std::min<int>(nb, (int)opt("topk.ctas_per_sm", 8)));
  const int64_t tile_rows = (int64_t)p.block * p.unroll * p.vec;
  const int64_t ntiles = std::max<int64_t>(1, (n + tile_rows - 1) / tile_rows);
  // a few waves of CTAs instead of one persistent wave: the hardware scheduler evens out the CTAs' finishing times
  const unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)d->num_sms * nb * std::max<int64_t>(1, opt("topk.waves", 1)));
  const size_t ncand = (size_t)grid * K;
  Scratch scratch;
  WDB_CUDA(scratch.alloc(ncand * 12 + 64 + (size_t)K * 12 + 64, s));
  char *buf = scratch.as<char>();
  long long *cand_r = (long long *)buf;
  long long *d_cnt = cand_r + ncand;
  unsigned *d_done = (unsigned *)(d_cnt + 1);
  float *cand_k = (float *)(d_cnt + 2);
  float *best_k = (float *)cand, *best_v = best_k + K, *no_keys = nullptr;
  long long *best_r = (long long *)(cand + (size_t)K * 8);
  auto ptrs = col_ptrs(p.spec, cols);
  long long nn = n, rb = row_base, m = (long long)ncand;
  int zero = 0;
  const float *tau0 = nullptr;
  if (opt("topk.fused", 1)) {
    if (topk_sample_threshold(d, s, scan, p, ptrs, n, row_base, K, (char *)(cand_k + ncand), cand_k, cand_r, &tau0)) return 1;
    WDB_CUDA(cudaMemsetAsync(d_done, 0, 4, s));
    void *args[] = {ptrs.data(), &nn, &rb, &cand_k, &cand_r, &tau0, &d_done, &best_k, &best_r, &zero, &best_v, &no_keys, &d_cnt};
    if (launch(scan, grid, p.block, 0, s, args)) return 1;
  } else {
    {
      void *args[] = {ptrs.data(), &nn, &rb, &cand_k, &cand_r, &tau0};
      if (launch(scan, grid, p.block, 0, s, args)) return 1;
    }
    {
      void *args[] = {&cand_k, &cand_r, &m, &best_k, &best_r};
      if (launch(fin, 1, p.block, 0, s, args)) return 1;
    }
    {
      void *args[] = {ptrs.data(), &rb, &best_k, &best_r, &zero, &best_v, &no_keys, &d_cnt};
      if (launch(emit, 1, 32, 0, s, args)) return 1;
    }
  }
  return 0;
}

static int topk_large(Device *d, cudaStream_t s, const wdb_col_t *cols, int ncols, const char *key, const char *val, const char *cond,
                      bool desc, int64_t k, int64_t offset, int64_t n, float *d_out_vals, float *d_out_keys, int64_t *h_n) {
  const bool limited = k >= 0;
  const int64_t K = limited ? k + offset : n;
  float tau = desc ? -INFINITY : INFINITY;
  int thresh = 0;
  if (limited && K > 0 && n > 0) {
    TopkPlan p;
    if (plan_topk(cols, ncols, key, val, cond, desc, 1, true, &p)) return 1;
    Kernel tb;
    if (get_kernel(d, gen_source(p.spec), "wdb_topk.cu", "wdb_tile_best", &tb)) return 1;
    const int64_t tile_rows = (int64_t)p.block * p.unroll * p.vec;
    const int64_t ntiles = (n + tile_rows - 1) / tile_rows;
    if (ntiles >= K) {
      Scratch tb_scratch;
      WDB_CUDA(tb_scratch.alloc(sizeof(float) * (size_t)ntiles, s));
      float *tile_best = tb_scratch.as<float>();
      auto ptrs = col_ptrs(p.spec, cols);
      long long nn = n, nt = ntiles;
      void *args[] = {ptrs.data(), &nn, &tile_best, &nt};
      const unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)d->num_sms * 8);
      if (launch(tb, grid, p.block, 0, s, args)) return 1;
      if (sort_f32(d, s, tile_best, nullptr, ntiles, !desc)) return 1;
      WDB_CUDA(cudaMemcpyAsync(&tau, tile_best + (K - 1), 4, cudaMemcpyDeviceToHost, s));
      WDB_CUDA(cudaStreamSynchronize(s));
      // at least K tiles hold a surviving row whose key is at least as good as tau, so the K-th best
      // row passes the test; tau == worst sentinel (fewer than K tiles with survivors) keeps everything
      thresh = desc ? 1 : 2;
    }
  }
  // candidates (value,key) in row order; grow the buffers if the first guess was too small
  int64_t cap = limited ? std::min<int64_t>(n, std::max<int64_t>(1 << 20, 64 * K)) : n;
  for (int attempt = 0; attempt < 2; ++attempt) {
    Scratch cand_scratch;
    WDB_CUDA(cand_scratch.alloc(sizeof(float) * (size_t)std::max<int64_t>(cap, 1) * 2, s));
    float *cv = cand_scratch.as<float>(), *ck = cv + std::max<int64_t>(cap, 1);
    int64_t c = 0;
    const char *cc = (cond && *cond) ? cond : "true";
    if (run_compact_ex(d, s, cols, ncols, val, key, cc, cv, ck, n, nullptr, &c, thresh, tau, cap)) return 1;
    if (c > cap) {  // more candidates than guessed (many ties at the threshold): retry with room for all
      cap = c;
      continue;
    }
    if (sort_f32(d, s, ck, cv, c, !desc)) return 1;
    const int64_t avail = std::max<int64_t>(c - offset, 0);
    const int64_t m = limited ? std::min<int64_t>(k, avail) : avail;
    if (m > 0) {
      if (d_out_vals) WDB_CUDA(cudaMemcpyAsync(d_out_vals, cv + offset, sizeof(float) * (size_t)m, cudaMemcpyDeviceToDevice, s));
      if (d_out_keys) WDB_CUDA(cudaMemcpyAsync(d_out_keys, ck + offset, sizeof(float) * (size_t)m, cudaMemcpyDeviceToDevice, s));
    }
    cand_scratch.release();
    if (h_n) { *h_n = m; WDB_CUDA(cudaStreamSynchronize(s)); }
    return 0;
  }
  return fail("top-k candidate buffer could not be sized");
}
}  // namespace wdb

using namespace wdb;

extern "C" int wdb_topk(int device, void *stream, const wdb_col_t *cols, int ncols, const char *key_expr, const char *val_expr,
                        const char *cond, int descending, int64_t k, int64_t offset, int64_t n, float *d_out_vals,
                        float *d_out_keys, int64_t *h_n) {
  if (!key_expr || !*key_expr) return fail("empty ORDER BY expression");
  if (!val_expr || !*val_expr) val_expr = key_expr;
  if (n < 0 || offset < 0) return fail("negative row count or offset");
  Device *d;
  if (get_device(device, &d)) return 1;
  cudaStream_t s = (cudaStream_t)stream;
  const std::string nan_last = order_key(key_expr, descending != 0);   // NaN keys order last in every path
  key_expr = nan_last.c_str();
  if (k == 0) { if (h_n) *h_n = 0; return 0; }
  const int64_t reg_max = opt("topk.reg_max", 16);
  if (k > 0 && k + offset <= reg_max)
    return topk_small(d, s, cols, ncols, key_expr, val_expr, cond, descending != 0, (int)(k + offset), (int)offset, n, d_out_vals, d_out_keys, h_n);
  return topk_large(d, s, cols, ncols, key_expr, val_expr, cond, descending != 0, k, offset, n, d_out_vals, d_out_keys, h_n);
}
