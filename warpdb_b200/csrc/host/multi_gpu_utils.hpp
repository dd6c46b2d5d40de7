// multi_gpu_utils.hpp -- reference: include/multi_gpu_utils.hpp:10-12
#pragma once
#include <string>
#include <vector>

#include "csv_loader.hpp"
#include "jit.hpp"

// Row-range shards over all visible GPUs (chunk = ceil(N/ndev)), results concatenated in row order.
std::vector<float> run_multi_gpu_jit_host(const HostTable &host, const std::string &expr_cuda, const std::string &cond_cuda);
