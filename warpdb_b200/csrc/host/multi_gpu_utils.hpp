// multi_gpu_utils.hpp -- reference: include/multi_gpu_utils.hpp:10-12
#pragma once
#include <string>
#include <vector>

#include "csv_loader.hpp"
#include "expression.hpp"
#include "jit.hpp"

// Row-range shards over all visible GPUs (chunk = ceil(N/ndev)), results concatenated in row order.
std::vector<float> run_multi_gpu_jit_host(const HostTable &host, const std::string &expr_cuda, const std::string &cond_cuda);

// Aggregates and ORDER BY ... LIMIT over the same row-range shards (not in the reference, whose
// multi-GPU path can only project / filter): every GPU aggregates / selects on its shard and the
// partial aggregates / top-k candidates are merged GPU to GPU with NCCL inside the core
// (wdb_multi_group_agg_host / wdb_multi_topk_host).  Groups come back in key order.
struct MultiGpuGroups {
  std::vector<int> keys;
  std::vector<float> vals;
};
MultiGpuGroups run_multi_gpu_group_host(const HostTable &host, const std::string &val_cuda, const std::string &key_cuda,
                                        const std::string &cond_cuda, AggregationType agg, bool descending = false);
std::vector<float> run_multi_gpu_topk_host(const HostTable &host, const std::string &key_cuda, const std::string &val_cuda,
                                           const std::string &cond_cuda, bool descending, int limit, int offset = 0);
// stable compaction over all GPUs: the surviving rows' values in row order
std::vector<float> run_multi_gpu_compact_host(const HostTable &host, const std::string &expr_cuda, const std::string &cond_cuda);
