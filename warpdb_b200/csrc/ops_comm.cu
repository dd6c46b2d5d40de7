// ops_comm.cu -- cross-GPU half of the path: communicators and the sharded operators that merge
// partial aggregates and top-k candidates over NVLink inside the core.
//
// Replaces the multi-GPU driver of the reference (src/multi_gpu_utils.cpp:5-63, src/warpdb.cpp:508-542):
// the same contiguous row-range shards chunk = ceil(N/ndev) (:24-31), but every GPU keeps its shard
// resident, runs the local kernel and the partial results are merged GPU to GPU (the reference
// copies every shard's full output through the host and has no aggregate / ORDER BY path at all).
//
//   GROUP BY   keys with a known range (optimizer statistics): every GPU aggregates into a
//              direct-addressed table over the SAME global key range, so the merge is one
//              ncclAllReduce per accumulator array (sum / sum / min / max) and the ordered export
//              runs on the merged table -- no host synchronisation between the local kernel and the
//              final groups.  Anything else (first-appearance order, huge or unknown ranges):
//              ordered partials, all-gather, merge into a table, export.
//   ORDER BY .. LIMIT k   every GPU selects its k+offset best (key, global row) pairs and evaluates
//              the SELECT expression there; ONE fixed-size all-gather and a one-warp selection over
//              world x (k+offset) candidates; ties keep global row order.
//   filter + compaction   no data-path collective: the survivors stay sharded; an all-gather of one
//              count per GPU yields every shard's global offset.
//
// NCCL is loaded with dlopen (libnccl.so.2: the copy already mapped by PyTorch when there is one),
// so the library still loads on a GPU-less box and links against nothing but the CUDA runtime.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <thread>
#include <tuple>

#include "agg.hpp"

namespace wdb {

// ---- NCCL, resolved at run time ------------------------------------------------------------------
struct Nccl {
  void *h = nullptr;
  int version = 0;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
};
static Nccl g_nccl;
static int load_nccl() {
  static std::mutex mu;
  std::lock_guard<std::mutex> l(mu);
  if (g_nccl.h) return 0;
  std::vector<std::string> cand;
  if (const char *e = getenv("WARPDB_NCCL")) cand.push_back(e);
  cand.push_back("libnccl.so.2");   // resolves to the copy PyTorch already mapped, if any
  cand.push_back("/usr/lib/x86_64-linux-gnu/libnccl.so.2");
  std::string tried;
  for (const auto &c : cand) {
    void *h = dlopen(c.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (!h) { tried += c + " "; continue; }
    Nccl n;
    n.h = h;
#define WDB_SYM(f) *(void **)(&n.f) = dlsym(h, "nccl" #f)
    WDB_SYM(GetVersion); WDB_SYM(GetUniqueId); WDB_SYM(CommInitRank); WDB_SYM(CommInitAll); WDB_SYM(CommDestroy);
    WDB_SYM(GetErrorString); WDB_SYM(AllReduce); WDB_SYM(AllGather); WDB_SYM(GroupStart); WDB_SYM(GroupEnd);
#undef WDB_SYM
    if (!n.GetUniqueId || !n.CommInitRank || !n.CommInitAll || !n.AllReduce || !n.AllGather || !n.GroupStart || !n.GroupEnd) {
      tried += c + "(symbols) ";
      dlclose(h);
      continue;
    }
    if (n.GetVersion) n.GetVersion(&n.version);
    g_nccl = n;
    return 0;
  }
  return fail("NCCL error: libnccl.so.2 not found (tried: %s); set WARPDB_NCCL", tried.c_str());
}
#define WDB_NCCL(call)                                                                                         \
  do {                                                                                                         \
    ncclResult_t r__ = (call);                                                                                 \
    if (r__ != ncclSuccess)                                                                                    \
      return ::wdb::fail("NCCL error: %s (%s)", g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?", #call); \
  } while (0)

}  // namespace wdb

struct wdb_comm {
  wdb::Device *dev = nullptr;
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
  std::map<std::tuple<int, int, int64_t>, wdb_agg *> tables;   // (role, needs, expected groups) -> table reused across calls
  char *scratch = nullptr;                                 // persistent exchange buffer (candidates, ranges, counts)
  size_t scratch_bytes = 0;
  cudaStream_t side = nullptr;                             // the all-reduce of a finished table slice runs here, next to the following slice's kernel
  cudaEvent_t ev_slice = nullptr, ev_done = nullptr;
};

using namespace wdb;

namespace wdb {

static int comm_scratch(wdb_comm *c, size_t bytes, char **out) {
  if (c->scratch_bytes < bytes) {
    if (c->scratch) { WDB_CUDA(cudaDeviceSynchronize()); WDB_CUDA(cudaFree(c->scratch)); c->scratch = nullptr; c->scratch_bytes = 0; }
    const size_t want = std::max<size_t>(bytes, 1 << 16);
    WDB_CUDA(cudaMalloc((void **)&c->scratch, want));
    c->scratch_bytes = want;
  }
  *out = c->scratch;
  return 0;
}

enum { kRolePartial = 0, kRoleMerge = 1 };
static int comm_table(wdb_comm *c, int role, int needs, int64_t expected, cudaStream_t s, wdb_agg **out) {
  auto key = std::make_tuple(role, needs, expected);
  auto it = c->tables.find(key);
  if (it != c->tables.end()) {
    *out = it->second;
    return wdb_agg_reset(it->second, s);
  }
  if (c->tables.size() >= 8) {   // bounded cache: drop everything rather than track recency
    WDB_CUDA(cudaDeviceSynchronize());
    for (auto &kv : c->tables) wdb_agg_destroy(kv.second);
    c->tables.clear();
  }
  wdb_agg *t = nullptr;
  if (agg_create_on(c->dev->id, expected, needs, s, &t)) return 1;
  c->tables[key] = t;
  *out = t;
  return 0;
}

// ---- static kernels ------------------------------------------------------------------------------
#define WDB_ROW_NONE_H 0x7fffffffffffffffll
// Final selection of a sharded ORDER BY ... LIMIT: `gathered` holds, per rank, [K keys f32 | K vals f32 |
// K rows i64] (rows are global row ids, WDB_ROW_NONE = empty).  One warp picks the K best by (key, row):
// the order of a stable sort of all rows, because equal keys are taken in ascending global row order.
__global__ void __launch_bounds__(32) topk_merge_kernel(const char *__restrict__ gathered, int nranks, int K, int desc, int offset,
                                                        float *__restrict__ out_vals, float *__restrict__ out_keys, long long *__restrict__ out_n) {
  __shared__ unsigned char taken[2048];
  const int total = nranks * K;
  const int lane = threadIdx.x;
  for (int i = lane; i < total; i += 32) taken[i] = 0;
  __syncwarp();
  const size_t per_rank = (size_t)K * 16;
  int found = 0;
  for (int round = 0; round < K; ++round) {
    float bk = 0.f;
    long long br = WDB_ROW_NONE_H;
    int bi = -1;
    for (int i = lane; i < total; i += 32) {
      if (taken[i]) continue;
      const char *base = gathered + (size_t)(i / K) * per_rank;
      const int j = i % K;
      const long long r = reinterpret_cast<const long long *>(base + (size_t)K * 8)[j];
      if (r == WDB_ROW_NONE_H) continue;
      const float k = reinterpret_cast<const float *>(base)[j];
      const bool better = bi < 0 || (desc ? (k > bk) : (k < bk)) || (k == bk && r < br);
      if (better) { bk = k; br = r; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ok = __shfl_xor_sync(0xffffffffu, bk, o);
      const long long orow = __shfl_xor_sync(0xffffffffu, br, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const bool better = oi >= 0 && (bi < 0 || (desc ? (ok > bk) : (ok < bk)) || (ok == bk && orow < br));
      if (better) { bk = ok; br = orow; bi = oi; }
    }
    if (bi < 0) break;   // warp-uniform after the butterfly
    if (lane == 0) {
      taken[bi] = 1;
      if (round >= offset) {
        const char *base = gathered + (size_t)(bi / K) * per_rank;
        if (out_vals) out_vals[round - offset] = reinterpret_cast<const float *>(base)[K + bi % K];
        if (out_keys) out_keys[round - offset] = bk;
      }
    }
    ++found;
    __syncwarp();
  }
  if (lane == 0) *out_n = found > offset ? found - offset : 0;
}

int topk_merge_launch(cudaStream_t s, const char *gathered, int nparts, int K, bool desc, int offset, float *out_vals, float *out_keys,
                      long long *out_n) {
  if ((size_t)nparts * (size_t)K > 2048) return fail("too many candidate lists for the register top-k merge");
  topk_merge_kernel<<<1, 32, 0, s>>>(gathered, nparts, K, desc ? 1 : 0, offset, out_vals, out_keys, out_n);
  stats().launches++;
  WDB_CUDA(cudaGetLastError());
  return 0;
}

// counts[r] = survivors of rank r (all-gathered) -> out3 = {count of this rank, its global offset, global total}
__global__ void compact_offsets_kernel(const long long *__restrict__ counts, int nranks, int rank, long long *__restrict__ out3) {
  long long off = 0, tot = 0;
  for (int r = 0; r < nranks; ++r) {
    if (r < rank) off += counts[r];
    tot += counts[r];
  }
  out3[0] = counts[rank];
  out3[1] = off;
  out3[2] = tot;
}

// ---- collectives ---------------------------------------------------------------------------------
static int allreduce(wdb_comm *c, void *buf, size_t count, ncclDataType_t dt, ncclRedOp_t op, cudaStream_t s) {
  if (c->nranks == 1 || count == 0) return 0;
  WDB_NCCL(g_nccl.AllReduce(buf, buf, count, dt, op, c->comm, s));
  return 0;
}
// in place: rank r's contribution sits at recv + r * bytes
static int allgather_inplace(wdb_comm *c, char *recv, size_t bytes, cudaStream_t s) {
  if (c->nranks == 1 || bytes == 0) return 0;
  WDB_NCCL(g_nccl.AllGather(recv + (size_t)c->rank * bytes, recv, bytes, ncclUint8, c->comm, s));
  return 0;
}

int topk_candidates(Device *d, cudaStream_t s, const wdb_col_t *cols, int ncols, const char *key, const char *val, const char *cond,
                    bool desc, int K, int64_t n, int64_t row_base, char *cand);
int sort_f32(Device *d, cudaStream_t s, float *d_keys, float *d_payload, long long n, bool ascending);
std::string order_key(const char *key_expr, bool desc);

}  // namespace wdb

extern "C" {

int wdb_comm_unique_id(void *id128) {
  if (!id128) return fail("null output");
  if (load_nccl()) return 1;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  WDB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return 0;
}

int wdb_comm_init_rank(int device, int nranks, int rank, const void *id128, wdb_comm_t **out) {
  if (!out) return fail("null output");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail("invalid rank %d of %d", rank, nranks);
  Device *d;
  if (get_device(device, &d)) return 1;
  wdb_comm *c = new wdb_comm();
  c->dev = d;
  c->rank = rank;
  c->nranks = nranks;
  if (nranks > 1) {
    if (!id128) { delete c; return fail("a communicator of %d ranks needs the unique id of wdb_comm_unique_id", nranks); }
    if (load_nccl()) { delete c; return 1; }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) { delete c; return fail("NCCL error: %s (ncclCommInitRank)", g_nccl.GetErrorString(r)); }
  }
  *out = c;
  return 0;
}

int wdb_comm_init_all(int ndev, const int *devices, wdb_comm_t **out) {
  if (!out) return fail("null output");
  int avail = 0;
  wdb_device_count(&avail);
  if (avail == 0) return fail("CUDA error: no CUDA device available; warpcore has no CPU fallback");
  if (ndev <= 0) ndev = avail;
  std::vector<int> devs(ndev);
  for (int i = 0; i < ndev; ++i) {
    devs[i] = devices ? devices[i] : i;
    if (devs[i] < 0 || devs[i] >= avail) return fail("invalid device id %d", devs[i]);
  }
  std::vector<ncclComm_t> comms(ndev, nullptr);
  if (ndev > 1) {
    if (load_nccl()) return 1;
    WDB_NCCL(g_nccl.CommInitAll(comms.data(), ndev, devs.data()));
  }
  for (int i = 0; i < ndev; ++i) {
    Device *d;
    if (get_device(devs[i], &d)) return 1;
    wdb_comm *c = new wdb_comm();
    c->dev = d;
    c->rank = i;
    c->nranks = ndev;
    c->comm = comms[i];
    out[i] = c;
  }
  return 0;
}

int wdb_comm_destroy(wdb_comm_t *c) {
  if (!c) return 0;
  cudaSetDevice(c->dev->id);
  cudaDeviceSynchronize();
  for (auto &kv : c->tables) wdb_agg_destroy(kv.second);
  if (c->scratch) cudaFree(c->scratch);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->ev_slice) cudaEventDestroy(c->ev_slice);
  if (c->ev_done) cudaEventDestroy(c->ev_done);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  delete c;
  return 0;
}

int wdb_comm_info(const wdb_comm_t *c, int *rank, int *nranks, int *device) {
  if (!c) return fail("null communicator");
  if (rank) *rank = c->rank;
  if (nranks) *nranks = c->nranks;
  if (device) *device = c->dev->id;
  return 0;
}

// ---- filter + project over resident shards ---------------------------------------------------------
int wdb_multi_project_filter(wdb_comm_t *c, void *stream, const wdb_col_t *cols, int ncols, const char *expr, const char *cond,
                             float *d_out, int64_t n_local, int mode, int64_t *d_count3, int64_t *h_count3) {
  if (!c) return fail("null communicator");
  WDB_CUDA(cudaSetDevice(c->dev->id));
  cudaStream_t s = (cudaStream_t)stream;
  char *sc;
  if (comm_scratch(c, 8 * (size_t)c->nranks + 32, &sc)) return 1;
  long long *counts = (long long *)sc, *out3 = counts + c->nranks;
  if (wdb_project_filter(c->dev->id, stream, cols, ncols, expr, cond, d_out, n_local, mode, (int64_t *)(counts + c->rank), nullptr)) return 1;
  if (allgather_inplace(c, sc, 8, s)) return 1;
  compact_offsets_kernel<<<1, 1, 0, s>>>(counts, c->nranks, c->rank, out3);
  stats().launches++;
  WDB_CUDA(cudaGetLastError());
  if (d_count3) WDB_CUDA(cudaMemcpyAsync(d_count3, out3, 24, cudaMemcpyDeviceToDevice, s));
  if (h_count3) {
    WDB_CUDA(cudaMemcpyAsync(h_count3, out3, 24, cudaMemcpyDeviceToHost, s));
    WDB_CUDA(cudaStreamSynchronize(s));
  }
  return 0;
}

// ---- GROUP BY over resident shards -------------------------------------------------------------------
int wdb_multi_group_agg(wdb_comm_t *c, void *stream, const wdb_col_t *cols, int ncols, const char *val_expr, const char *key_expr,
                        const char *cond, int agg, int order, int64_t n_local, int64_t row_base, int64_t expected_groups,
                        int range_known, int64_t key_lo, int64_t key_hi, int32_t *d_keys, float *d_vals, int64_t cap,
                        int64_t *d_groups, int64_t *h_groups) {
  if (!c) return fail("null communicator");
  if (agg < WDB_SUM || agg > WDB_MAX) return fail("invalid aggregation %d", agg);
  if (order < 0 || order > 2) return fail("invalid order %d", order);
  if (!key_expr || !*key_expr) return fail("empty GROUP BY key expression");
  if (n_local < 0) return fail("negative row count");
  if (range_known && (key_lo > key_hi || key_lo < INT32_MIN || key_hi > INT32_MAX)) return fail("invalid key range [%lld, %lld]", (long long)key_lo, (long long)key_hi);
  Device *d = c->dev;
  WDB_CUDA(cudaSetDevice(d->id));
  cudaStream_t s = (cudaStream_t)stream;
  const int needs = needs_for_agg(agg) | (order == WDB_ORDER_FIRST ? WDB_NEED_FIRST_BIT : 0);
  if (expected_groups <= 0) expected_groups = 1 << 16;
  char *sc;
  if (comm_scratch(c, 256 + 8 * (size_t)c->nranks, &sc)) return 1;
  long long *d_stat = (long long *)sc;            // [lo, -hi, -max rows, min rows] (MIN all-reduce) | groups | spare
  long long *d_total = d_stat + 4;

  // 1. the GLOBAL key range (every rank must take the same path and index the same table layout)
  long long st[4] = {INT64_MAX, INT64_MAX, -(long long)n_local, (long long)n_local};   // MIN-reduced: lo, -hi, -max rows, min rows
  if (range_known) { st[0] = key_lo; st[1] = -key_hi; }
  else if (n_local > 0) {
    KeyRange r{false, 0, -1};
    if (auto_key_range(d, s, cols, ncols, key_expr, n_local, &r)) return 1;
    if (r.known) { st[0] = r.lo; st[1] = -r.hi; }
  }
  if (c->nranks > 1) {
    WDB_CUDA(cudaMemcpyAsync(d_stat, st, 32, cudaMemcpyHostToDevice, s));
    if (allreduce(c, d_stat, 4, ncclInt64, ncclMin, s)) return 1;
    WDB_CUDA(cudaMemcpyAsync(st, d_stat, 32, cudaMemcpyDeviceToHost, s));
    WDB_CUDA(cudaStreamSynchronize(s));
  }
  const bool known = st[0] != INT64_MAX && st[1] != INT64_MAX;
  const int64_t lo = known ? st[0] : 0, hi = known ? -st[1] : -1, span = hi - lo + 1, rows_max = -st[2], rows_min = st[3];
  const bool fast = known && !(needs & WDB_NEED_FIRST_BIT) && span <= opt("group.dense_max_span", 1 << 26) &&
                    (span <= (1 << 20) || span <= 4 * rows_max * c->nranks);

  if (fast) {
    // 2a. local aggregation; whatever kernel the optimizer picks, the partial ends in the side table
    wdb_agg *t = nullptr;
    if (comm_table(c, kRolePartial, needs, span < 32768 ? std::max<int64_t>(span, 1024) : 1024, s, &t)) return 1;
    if (wdb_agg_set_key_range(t, 1, lo, hi)) return 1;
    // A table beyond the L2 is filled slice by slice (one launch per index slice, ops_group.cu): the all-reduce of a
    // finished slice runs on a side stream while the next slice's kernel streams the columns again, so only the
    // last slice's exchange is exposed (10 M keys, 8 GPUs: 0.34 ms of NCCL for 80 MB, half of it hidden).
    int64_t reduced_to = 0;   // entries [0, reduced_to) have been all-reduced slice by slice
    bool overlap_failed = false;
    // Every rank must issue the same collectives: the slice-by-slice exchange is armed only where EVERY rank is sure to
    // take the sliced direct-addressed path (the decision of wdb_agg_consume, evaluated on the smallest shard)
    const bool all_sliced = rows_min > 0 && span >= opt("group.dense_min_span", 32768) && (span <= (1 << 20) || span <= 4 * rows_min) &&
                            span > opt("group.wp_max_span", 4096);
    if (c->nranks > 1 && all_sliced && opt("multi.overlap_slices", 1)) {
      if (!c->side) {
        int prio_lo = 0, prio_hi = 0;
        WDB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        // highest priority: the exchange's CTAs take the first CTA slots the aggregation kernel frees (its CTAs are
        // kept short while an exchange may be pending: group.dense_waves below)
        WDB_CUDA(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, prio_hi));
        WDB_CUDA(cudaEventCreateWithFlags(&c->ev_slice, cudaEventDisableTiming));
        WDB_CUDA(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
      }
      t->after_slice = [&, t](unsigned slo, unsigned shi) -> int {
        if ((int64_t)slo != reduced_to) { overlap_failed = true; return 0; }   // slices come in index order; anything else: fall back to one exchange at the end
        WDB_CUDA(cudaEventRecord(c->ev_slice, s));
        WDB_CUDA(cudaStreamWaitEvent(c->side, c->ev_slice, 0));
        const size_t cnt = (size_t)shi - slo;
        WDB_NCCL(g_nccl.GroupStart());
        int rc = 0;
        if (needs & WDB_NEED_SUM_BIT) rc |= allreduce(c, t->T.dsums + slo, cnt, ncclFloat64, ncclSum, c->side);
        if (needs & WDB_NEED_CNT_BIT) rc |= allreduce(c, t->T.dcnts + slo, cnt, ncclUint64, ncclSum, c->side);
        if (needs & WDB_NEED_MINMAX_BIT) {
          rc |= allreduce(c, t->T.dmins + slo, cnt, ncclInt64, ncclMin, c->side);
          rc |= allreduce(c, t->T.dmaxs + slo, cnt, ncclInt64, ncclMax, c->side);
        }
        WDB_NCCL(g_nccl.GroupEnd());
        if (rc) return 1;
        reduced_to = shi;
        return 0;
      };
    }
    const int crc = wdb_agg_consume(t, stream, cols, ncols, val_expr, key_expr, cond, n_local, row_base);
    t->after_slice = nullptr;
    if (c->side && reduced_to > 0) {   // the main stream continues after the side stream's exchanges
      WDB_CUDA(cudaEventRecord(c->ev_done, c->side));
      WDB_CUDA(cudaStreamWaitEvent(s, c->ev_done, 0));
    }
    if (crc) return 1;
    if (overlap_failed || (reduced_to != 0 && reduced_to != span)) return fail("internal: table slices were not exchanged in order");
    if (dense_prepare(t, s, lo, span)) return 1;
    if (agg_hash_to_dense(t, s)) return 1;
    // 3a. merge: one all-reduce per accumulator array over NVLink (unless the slices have been exchanged already)
    if (c->nranks > 1 && reduced_to == 0) {
      WDB_NCCL(g_nccl.GroupStart());
      int rc = 0;
      if (needs & WDB_NEED_SUM_BIT) rc |= allreduce(c, t->T.dsums, (size_t)span, ncclFloat64, ncclSum, s);
      if (needs & WDB_NEED_CNT_BIT) rc |= allreduce(c, t->T.dcnts, (size_t)span, ncclUint64, ncclSum, s);
      if (needs & WDB_NEED_MINMAX_BIT) {
        rc |= allreduce(c, t->T.dmins, (size_t)span, ncclInt64, ncclMin, s);
        rc |= allreduce(c, t->T.dmaxs, (size_t)span, ncclInt64, ncclMax, s);
      }
      WDB_NCCL(g_nccl.GroupEnd());
      if (rc) return 1;
    }
    // 4a. ordered export of the merged table (every rank ends with the final groups)
    if (agg_export_dense_async(t, s, agg, order, d_keys, d_vals, nullptr, nullptr, nullptr, nullptr, cap, d_total)) return 1;
    if (d_groups) WDB_CUDA(cudaMemcpyAsync(d_groups, d_total, 8, cudaMemcpyDeviceToDevice, s));
    if (h_groups) {
      long long g = 0;
      unsigned meta[8];
      WDB_CUDA(cudaMemcpyAsync(&g, d_total, 8, cudaMemcpyDeviceToHost, s));
      WDB_CUDA(cudaMemcpyAsync(meta, t->T.meta, 32, cudaMemcpyDeviceToHost, s));
      WDB_CUDA(cudaStreamSynchronize(s));
      if (meta[1]) return fail("aggregation table overflow");
      if (meta[4]) return fail("key statistics are stale: %u groups lie outside the promised key range [%lld, %lld]", meta[4], (long long)lo, (long long)hi);
      if (g > cap) return fail("%lld groups exceed the output capacity %lld", g, (long long)cap);
      *h_groups = g;
    }
    return 0;
  }

  // 2b. general path: ordered partials -> all-gather -> merge -> export (host-synchronised sizes)
  wdb_agg *t = nullptr;
  if (comm_table(c, kRolePartial, needs, expected_groups, s, &t)) return 1;
  if (wdb_agg_set_key_range(t, known ? 1 : 0, lo, hi)) return 1;
  if (n_local > 0 && wdb_agg_consume(t, stream, cols, ncols, val_expr, key_expr, cond, n_local, row_base)) return 1;
  if (c->nranks == 1) {
    int64_t g = 0;
    if (wdb_agg_export(t, stream, agg, order, d_keys, d_vals, nullptr, nullptr, nullptr, nullptr, nullptr, cap, &g)) return 1;
    long long gg = g;
    if (d_groups) WDB_CUDA(cudaMemcpyAsync(d_groups, &gg, 8, cudaMemcpyHostToDevice, s));
    if (h_groups) *h_groups = g;
    WDB_CUDA(cudaStreamSynchronize(s));
    return 0;
  }
  int64_t g_local = 0;
  if (wdb_agg_size(t, stream, &g_local)) return 1;
  long long *d_counts = d_stat + 8;
  long long gl = g_local;
  WDB_CUDA(cudaMemcpyAsync(d_counts + c->rank, &gl, 8, cudaMemcpyHostToDevice, s));
  if (allgather_inplace(c, (char *)d_counts, 8, s)) return 1;
  std::vector<long long> counts(c->nranks);
  WDB_CUDA(cudaMemcpyAsync(counts.data(), d_counts, 8 * (size_t)c->nranks, cudaMemcpyDeviceToHost, s));
  WDB_CUDA(cudaStreamSynchronize(s));
  long long gmax = 0, gsum = 0;
  for (auto x : counts) { gmax = std::max(gmax, x); gsum += x; }
  if (gsum == 0) {
    long long z = 0;
    if (d_groups) WDB_CUDA(cudaMemcpyAsync(d_groups, &z, 8, cudaMemcpyHostToDevice, s));
    if (h_groups) *h_groups = 0;
    WDB_CUDA(cudaStreamSynchronize(s));
    return 0;
  }
  // gathered arrays: keys i32 | sums f64 | counts i64 | mins f64 | maxs f64 | first i64, each [nranks][gmax]
  const size_t G = (size_t)((gmax + 1) / 2 * 2), W = (size_t)c->nranks;
  Scratch gathered;
  WDB_CUDA(gathered.alloc(W * G * (4 + 8 * 5) + 64, s));
  char *p = gathered.as<char>();
  double *a_sums = (double *)p; p += W * G * 8;
  long long *a_cnts = (long long *)p; p += W * G * 8;
  double *a_mins = (double *)p; p += W * G * 8;
  double *a_maxs = (double *)p; p += W * G * 8;
  long long *a_first = (long long *)p; p += W * G * 8;
  int *a_keys = (int *)p;
  const size_t me = (size_t)c->rank * G;
  int64_t gchk = 0;
  if (g_local > 0 &&
      wdb_agg_export(t, stream, needs & WDB_NEED_SUM_BIT ? WDB_SUM : (needs & WDB_NEED_CNT_BIT ? WDB_COUNT : WDB_MIN), WDB_ORDER_KEY_ASC, a_keys + me, nullptr,
                     (needs & WDB_NEED_SUM_BIT) ? a_sums + me : nullptr, (needs & WDB_NEED_CNT_BIT) ? (int64_t *)a_cnts + me : nullptr,
                     (needs & WDB_NEED_MINMAX_BIT) ? a_mins + me : nullptr, (needs & WDB_NEED_MINMAX_BIT) ? a_maxs + me : nullptr,
                     (needs & WDB_NEED_FIRST_BIT) ? (int64_t *)a_first + me : nullptr, (int64_t)G, &gchk))
    return 1;
  {
    WDB_NCCL(g_nccl.GroupStart());
    int rc = allgather_inplace(c, (char *)a_keys, G * 4, s);
    if (needs & WDB_NEED_SUM_BIT) rc |= allgather_inplace(c, (char *)a_sums, G * 8, s);
    if (needs & WDB_NEED_CNT_BIT) rc |= allgather_inplace(c, (char *)a_cnts, G * 8, s);
    if (needs & WDB_NEED_MINMAX_BIT) { rc |= allgather_inplace(c, (char *)a_mins, G * 8, s); rc |= allgather_inplace(c, (char *)a_maxs, G * 8, s); }
    if (needs & WDB_NEED_FIRST_BIT) rc |= allgather_inplace(c, (char *)a_first, G * 8, s);
    WDB_NCCL(g_nccl.GroupEnd());
    if (rc) return 1;
  }
  wdb_agg *m = nullptr;
  {  // capacity: the distinct keys of all partials, rounded up to a power of two so that repeated queries reuse the table
    int64_t want = 1024;
    while (want < gsum) want <<= 1;
    if (comm_table(c, kRoleMerge, needs, want, s, &m)) return 1;
  }
  if (wdb_agg_set_key_range(m, known ? 1 : 0, lo, hi)) return 1;
  for (int r = 0; r < c->nranks; ++r) {
    const size_t o = (size_t)r * G;
    if (counts[r] > 0 && wdb_agg_merge(m, stream, a_keys + o, a_sums + o, (const int64_t *)a_cnts + o, a_mins + o, a_maxs + o, (const int64_t *)a_first + o, counts[r]))
      return 1;
  }
  int64_t g = 0;
  if (wdb_agg_export(m, stream, agg, order, d_keys, d_vals, nullptr, nullptr, nullptr, nullptr, nullptr, cap, &g)) return 1;
  gathered.release();
  long long gg = g;
  if (d_groups) WDB_CUDA(cudaMemcpyAsync(d_groups, &gg, 8, cudaMemcpyHostToDevice, s));
  if (h_groups) *h_groups = g;
  WDB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

// ---- ORDER BY ... LIMIT over resident shards ---------------------------------------------------------
int wdb_multi_topk(wdb_comm_t *c, void *stream, const wdb_col_t *cols, int ncols, const char *key_expr, const char *val_expr,
                   const char *cond, int descending, int64_t k, int64_t offset, int64_t n_local, int64_t row_base,
                   float *d_out_vals, float *d_out_keys, int64_t *d_n, int64_t *h_n) {
  if (!c) return fail("null communicator");
  if (!key_expr || !*key_expr) return fail("empty ORDER BY expression");
  if (!val_expr || !*val_expr) val_expr = key_expr;
  if (n_local < 0 || offset < 0) return fail("negative row count or offset");
  if (k < 0) return fail("a sharded ORDER BY needs a LIMIT");
  Device *d = c->dev;
  WDB_CUDA(cudaSetDevice(d->id));
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t K = k + offset;
  if (k == 0) {
    long long z = 0;
    if (d_n) WDB_CUDA(cudaMemcpyAsync(d_n, &z, 8, cudaMemcpyHostToDevice, s));
    if (h_n) *h_n = 0;
    return 0;
  }
  if (K <= opt("topk.reg_max", 16)) {
    // local winners straight into this rank's slot of the gather buffer, one all-gather, one-warp selection
    const size_t per_rank = (size_t)K * 16;
    char *sc;
    if (comm_scratch(c, per_rank * c->nranks + 64, &sc)) return 1;
    long long *d_cnt = (long long *)(sc + per_rank * c->nranks);
    if (topk_candidates(d, s, cols, ncols, order_key(key_expr, descending != 0).c_str(), val_expr, cond, descending != 0, (int)K, n_local, row_base,
                        sc + per_rank * c->rank))
      return 1;
    if (allgather_inplace(c, sc, per_rank, s)) return 1;
    if (topk_merge_launch(s, sc, c->nranks, (int)K, descending != 0, (int)offset, d_out_vals, d_out_keys, d_cnt)) return 1;
    if (d_n) WDB_CUDA(cudaMemcpyAsync(d_n, d_cnt, 8, cudaMemcpyDeviceToDevice, s));
    if (h_n) {
      long long cnt = 0;
      WDB_CUDA(cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, s));
      WDB_CUDA(cudaStreamSynchronize(s));
      *h_n = cnt;
    }
    return 0;
  }
  // large limits: K best (value, key) pairs per rank, gathered in rank order (= global row order among
  // equal keys), one stable sort of the candidates, slice
  const size_t W = (size_t)c->nranks;
  Scratch scratch;
  WDB_CUDA(scratch.alloc((W * (size_t)K * 4 + 2 * W * (size_t)K) * sizeof(float) + 64, s));
  float *buf = scratch.as<float>();
  float *g_vals = buf, *g_keys = buf + W * K, *s_vals = g_keys + W * K, *s_keys = s_vals + W * K;
  int64_t c_local = 0;
  if (wdb_topk(d->id, stream, cols, ncols, key_expr, val_expr, cond, descending, K, 0, n_local, g_vals + (size_t)c->rank * K, g_keys + (size_t)c->rank * K, &c_local))
    return 1;
  char *sc;
  if (comm_scratch(c, 8 * W + 64, &sc)) return 1;
  long long *d_counts = (long long *)sc, cl = c_local;
  WDB_CUDA(cudaMemcpyAsync(d_counts + c->rank, &cl, 8, cudaMemcpyHostToDevice, s));
  std::vector<long long> counts(W, cl);
  if (c->nranks > 1) {
    WDB_NCCL(g_nccl.GroupStart());
    int rc = allgather_inplace(c, (char *)d_counts, 8, s);
    rc |= allgather_inplace(c, (char *)g_vals, (size_t)K * 4, s);
    rc |= allgather_inplace(c, (char *)g_keys, (size_t)K * 4, s);
    WDB_NCCL(g_nccl.GroupEnd());
    if (rc) return 1;
    WDB_CUDA(cudaMemcpyAsync(counts.data(), d_counts, 8 * W, cudaMemcpyDeviceToHost, s));
    WDB_CUDA(cudaStreamSynchronize(s));
  }
  int64_t total = 0;
  for (size_t r = 0; r < W; ++r) {
    if (counts[r] > 0) {
      WDB_CUDA(cudaMemcpyAsync(s_vals + total, g_vals + r * K, (size_t)counts[r] * 4, cudaMemcpyDeviceToDevice, s));
      WDB_CUDA(cudaMemcpyAsync(s_keys + total, g_keys + r * K, (size_t)counts[r] * 4, cudaMemcpyDeviceToDevice, s));
    }
    total += counts[r];
  }
  if (total > 0 && sort_f32(d, s, s_keys, s_vals, total, descending == 0)) return 1;
  const long long m = std::max<int64_t>(0, std::min<int64_t>(k, total - offset));
  if (m > 0) {
    if (d_out_vals) WDB_CUDA(cudaMemcpyAsync(d_out_vals, s_vals + offset, (size_t)m * 4, cudaMemcpyDeviceToDevice, s));
    if (d_out_keys) WDB_CUDA(cudaMemcpyAsync(d_out_keys, s_keys + offset, (size_t)m * 4, cudaMemcpyDeviceToDevice, s));
  }
  if (d_n) WDB_CUDA(cudaMemcpyAsync(d_n, &m, 8, cudaMemcpyHostToDevice, s));
  scratch.release();
  WDB_CUDA(cudaStreamSynchronize(s));
  if (h_n) *h_n = m;
  return 0;
}

}  // extern "C"

// ---- host-buffer variants: one process drives every GPU (the shape of run_multi_gpu_jit_host) ---------
namespace wdb {

struct HostComms {
  std::vector<int> devs;
  std::vector<wdb_comm *> comms;
};
static std::mutex g_host_mu;
static HostComms g_host;

static int host_comms(int ndev, const int *devices, HostComms **out) {
  int avail = 0;
  wdb_device_count(&avail);
  if (avail == 0) return fail("CUDA error: no CUDA device available; warpcore has no CPU fallback");
  if (ndev <= 0) ndev = avail;
  std::vector<int> devs(ndev);
  for (int i = 0; i < ndev; ++i) devs[i] = devices ? devices[i] : i;
  if (g_host.devs != devs) {
    for (auto *c : g_host.comms) wdb_comm_destroy(c);
    g_host.comms.assign(ndev, nullptr);
    g_host.devs.clear();
    if (wdb_comm_init_all(ndev, devs.data(), g_host.comms.data())) { g_host.comms.clear(); return 1; }
    g_host.devs = devs;
  }
  *out = &g_host;
  return 0;
}

// upload the columns a query reads of shard [start, end) to `dev`; returns device column descriptors
struct ShardUpload {
  std::vector<wdb_col_t> cols;
  std::vector<void *> bufs;
  cudaStream_t stream = nullptr;
};
static int upload_shard(int dev, const wdb_col_t *h_cols, int ncols, const std::vector<std::string> &texts, int64_t start, int64_t end,
                        ShardUpload *up) {
  Device *d;
  if (get_device(dev, &d)) return 1;
  WDB_CUDA(cudaStreamCreateWithFlags(&up->stream, cudaStreamNonBlocking));
  const int64_t rows = end - start;
  up->cols.assign(h_cols, h_cols + ncols);
  for (auto &c : up->cols) { c.dptr = nullptr; c.len = rows; }
  for (const auto &u : find_used_columns(h_cols, ncols, texts)) {
    const size_t sz = dtype_size(u.dtype);
    if (sz == 0) return fail("column %s has a non-numeric type and cannot be read on the GPU", u.name.c_str());
    void *p = nullptr;
    WDB_CUDA(cudaMallocAsync(&p, std::max<size_t>((size_t)rows * sz, 256), up->stream));
    up->bufs.push_back(p);
    if (rows > 0)
      WDB_CUDA(cudaMemcpyAsync(p, (const char *)h_cols[u.table_index].dptr + (size_t)start * sz, (size_t)rows * sz, cudaMemcpyHostToDevice, up->stream));
    up->cols[u.table_index].dptr = p;
  }
  return 0;
}
// one host thread per device runs fn(i); the first failure's message is kept
template <class F> static int per_device(int W, F fn) {
  std::vector<int> rcs(W, 0);
  std::vector<std::string> errs(W);
  std::vector<std::thread> th;
  for (int i = 0; i < W; ++i)
    th.emplace_back([&, i]() {
      rcs[i] = fn(i);
      if (rcs[i]) errs[i] = wdb_last_error();
    });
  for (auto &t : th) t.join();
  for (int i = 0; i < W; ++i)
    if (rcs[i]) { set_error(errs[i]); return 1; }
  return 0;
}

static void release_shard(ShardUpload *up) {
  if (!up->stream) return;
  for (void *p : up->bufs) cudaFreeAsync(p, up->stream);
  cudaStreamSynchronize(up->stream);
  cudaStreamDestroy(up->stream);
  up->stream = nullptr;
}

}  // namespace wdb

extern "C" {

int wdb_multi_group_agg_host(int ndev, const int *devices, const wdb_col_t *h_cols, int ncols, const char *val_expr, const char *key_expr,
                             const char *cond, int agg, int order, int64_t n, int64_t expected_groups, int32_t *h_keys, float *h_vals,
                             int64_t cap, int64_t *h_groups) {
  if (n < 0 || cap < 0) return fail("negative row count or capacity");
  if (!key_expr || !*key_expr) return fail("empty GROUP BY key expression");
  std::lock_guard<std::mutex> lock(g_host_mu);
  HostComms *hc;
  if (host_comms(ndev, devices, &hc)) return 1;
  const int W = (int)hc->comms.size();
  std::vector<ShardUpload> up(W);
  std::vector<int32_t *> d_keys(W, nullptr);
  std::vector<int64_t> groups(W, 0);
  // phase 1 (no collective yet, so a local failure cannot strand the other ranks): upload the shards
  int rc = per_device(W, [&](int i) {
    int64_t start = 0, end = 0;
    wdb_shard_range(n, W, i, &start, &end);
    if (upload_shard(hc->devs[i], h_cols, ncols, {val_expr ? val_expr : "", key_expr, cond ? cond : ""}, start, end, &up[i])) return 1;
    WDB_CUDA(cudaMallocAsync((void **)&d_keys[i], std::max<size_t>((size_t)cap * 8, 256), up[i].stream));
    return 0;
  });
  // phase 2: local aggregation + NCCL merge; every rank ends with the final groups, rank 0 returns them
  if (!rc) rc = per_device(W, [&](int i) {
    int64_t start = 0, end = 0;
    wdb_shard_range(n, W, i, &start, &end);
    float *d_vals = (float *)(d_keys[i] + cap);
    if (wdb_multi_group_agg(hc->comms[i], up[i].stream, up[i].cols.data(), ncols, val_expr, key_expr, cond, agg, order, end - start, start,
                            expected_groups, 0, 0, -1, d_keys[i], d_vals, cap, nullptr, &groups[i]))
      return 1;
    if (i == 0 && groups[i] > 0) {
      if (h_keys) WDB_CUDA(cudaMemcpyAsync(h_keys, d_keys[i], (size_t)groups[i] * 4, cudaMemcpyDeviceToHost, up[i].stream));
      if (h_vals) WDB_CUDA(cudaMemcpyAsync(h_vals, d_vals, (size_t)groups[i] * 4, cudaMemcpyDeviceToHost, up[i].stream));
      WDB_CUDA(cudaStreamSynchronize(up[i].stream));
    }
    return 0;
  });
  const std::string err = rc ? wdb_last_error() : "";
  for (int i = 0; i < W; ++i) {
    if (up[i].stream) {
      cudaSetDevice(hc->devs[i]);
      if (d_keys[i]) cudaFreeAsync(d_keys[i], up[i].stream);
      release_shard(&up[i]);
    }
  }
  if (rc) { set_error(err); return 1; }
  if (h_groups) *h_groups = groups[0];
  return 0;
}

int wdb_multi_topk_host(int ndev, const int *devices, const wdb_col_t *h_cols, int ncols, const char *key_expr, const char *val_expr,
                        const char *cond, int descending, int64_t k, int64_t offset, int64_t n, float *h_out_vals, int64_t *h_n) {
  if (n < 0) return fail("negative row count");
  if (!key_expr || !*key_expr) return fail("empty ORDER BY expression");
  if (k < 0) return fail("a sharded ORDER BY needs a LIMIT");
  std::lock_guard<std::mutex> lock(g_host_mu);
  HostComms *hc;
  if (host_comms(ndev, devices, &hc)) return 1;
  const int W = (int)hc->comms.size();
  std::vector<ShardUpload> up(W);
  std::vector<float *> d_vals(W, nullptr);
  std::vector<int64_t> counts(W, 0);
  int rc = per_device(W, [&](int i) {
    int64_t start = 0, end = 0;
    wdb_shard_range(n, W, i, &start, &end);
    if (upload_shard(hc->devs[i], h_cols, ncols, {key_expr, val_expr ? val_expr : "", cond ? cond : ""}, start, end, &up[i])) return 1;
    WDB_CUDA(cudaMallocAsync((void **)&d_vals[i], std::max<size_t>((size_t)std::max<int64_t>(k, 1) * 4, 256), up[i].stream));
    return 0;
  });
  if (!rc) rc = per_device(W, [&](int i) {
    int64_t start = 0, end = 0;
    wdb_shard_range(n, W, i, &start, &end);
    if (wdb_multi_topk(hc->comms[i], up[i].stream, up[i].cols.data(), ncols, key_expr, val_expr, cond, descending, k, offset, end - start, start,
                       d_vals[i], nullptr, nullptr, &counts[i]))
      return 1;
    if (i == 0 && counts[i] > 0 && h_out_vals) {
      WDB_CUDA(cudaMemcpyAsync(h_out_vals, d_vals[i], (size_t)counts[i] * 4, cudaMemcpyDeviceToHost, up[i].stream));
      WDB_CUDA(cudaStreamSynchronize(up[i].stream));
    }
    return 0;
  });
  const std::string err = rc ? wdb_last_error() : "";
  for (int i = 0; i < W; ++i) {
    if (up[i].stream) {
      cudaSetDevice(hc->devs[i]);
      if (d_vals[i]) cudaFreeAsync(d_vals[i], up[i].stream);
      release_shard(&up[i]);
    }
  }
  if (rc) { set_error(err); return 1; }
  if (h_n) *h_n = counts[0];
  return 0;
}

}  // extern "C"
