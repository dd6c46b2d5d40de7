// jit.hpp -- the reference's execution seam (include/jit.hpp:7-27) with identical signatures,
// implemented on the C ABI of libwarpcore.so (include/warpcore.h).  Errors surface as
// std::runtime_error like the reference's NVRTC_CHECK / CU_CHECK (src/jit.cpp:11-28) and keep the
// text "Kernel compilation failed." (src/jit.cpp:128).
#pragma once
#include <string>

#include "csv_loader.hpp"

// if (cond) output[i] = expr, for every row of `table`; rows failing cond keep their slot untouched
void jit_compile_and_launch(const std::string &expr_code, const std::string &condition_code, const Table &table,
                            float *d_output, int device_id = 0);
// SUM(val) GROUP BY (int)key over N rows of (price, quantity); groups in first-appearance order,
// keys/values written compactly, *d_count = number of groups
void jit_group_sum(const std::string &val_expr_code, const std::string &key_expr_code, float *d_price, int *d_quantity,
                   float *d_out_vals, int *d_out_keys, int *d_count, int N, int device_id = 0);
// stable in-place sorts (ascending or descending)
void jit_sort_pairs(int *d_keys, float *d_vals, int count, bool ascending, int device_id = 0);
void jit_sort_float(float *d_vals, int count, bool ascending, int device_id = 0);
