// jit_shim_test.cpp -- exercises the reference-signature C++ seam (jit.hpp, multi_gpu_utils.hpp,
// optimizer.hpp, warpdb.hpp of warpdb_b200/csrc/host) the way the reference's own tests do:
// tests/jit_arch_test.cpp, jit_error_test.cpp, sql_features_test.cpp, having_distinct_test.cpp,
// extended_types_test.cpp.  Run from tests/data (needs test.csv, extended.csv, custom.cu).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <functional>
#include <iostream>
#include <map>
#include <stdexcept>
#include <vector>

#include "jit.hpp"
#include "multi_gpu_utils.hpp"
#include "optimizer.hpp"
#include "warpdb.hpp"

#define CHECK(cond)                                                               \
  do {                                                                            \
    if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
  } while (0)

template <class T> T *to_device(const std::vector<T> &h) {
  T *d = nullptr;
  cudaMalloc(&d, sizeof(T) * std::max<size_t>(h.size(), 1));
  cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice);
  return d;
}
template <class T> std::vector<T> to_host(const T *d, size_t n) {
  std::vector<T> h(n);
  cudaMemcpy(h.data(), d, sizeof(T) * n, cudaMemcpyDeviceToHost);
  return h;
}

int main() {
  // ---- tests/jit_arch_test.cpp: expr "price" over a one-row table returns the price bit-exactly
  {
    float *d_price = to_device(std::vector<float>{2.0f});
    int *d_quantity = to_device(std::vector<int>{0});
    float *d_output = to_device(std::vector<float>{0.0f});
    Table table;
    table.num_rows = 1;
    table.columns.push_back({"price", DataType::Float32, d_price, 1});
    table.columns.push_back({"quantity", DataType::Int32, d_quantity, 1});
    jit_compile_and_launch("price[idx]", "", table, d_output);
    CHECK(to_host(d_output, 1)[0] == 2.0f);
    // ---- tests/jit_error_test.cpp: a failed compilation throws and leaves the library usable
    bool threw = false;
    try {
      jit_compile_and_launch("invalid@", "", table, d_output);
    } catch (const std::exception &e) {
      threw = std::string(e.what()) == "Kernel compilation failed.";
    }
    CHECK(threw);
    jit_compile_and_launch("(price[idx] + 1.0f)", "", table, d_output);
    CHECK(to_host(d_output, 1)[0] == 3.0f);
    // reference semantics of the dense filter: slots failing the condition stay untouched
    cudaMemcpy(d_output, std::vector<float>{-7.0f}.data(), 4, cudaMemcpyHostToDevice);
    jit_compile_and_launch("price[idx]", "(price[idx] > 5.0f)", table, d_output);
    CHECK(to_host(d_output, 1)[0] == -7.0f);
    cudaFree(d_price); cudaFree(d_quantity); cudaFree(d_output);
  }
  // ---- jit_group_sum / jit_sort_pairs / jit_sort_float (include/jit.hpp:15-27)
  {
    const std::vector<float> price = {10.5f, 20.0f, 15.25f, 30.0f, 1.0f, 2.0f};
    const std::vector<int> qty = {3, 4, 2, 5, 4, 3};
    float *d_price = to_device(price);
    int *d_qty = to_device(qty);
    float *d_vals; int *d_keys; int *d_count;
    cudaMalloc(&d_vals, 4 * 6); cudaMalloc(&d_keys, 4 * 6); cudaMalloc(&d_count, 4);
    jit_group_sum("price[idx]", "quantity[idx]", d_price, d_qty, d_vals, d_keys, d_count, 6);
    const int g = to_host(d_count, 1)[0];
    CHECK(g == 4);
    // first-appearance order (src/jit.cpp:196-213): keys 3,4,2,5
    CHECK((to_host(d_keys, 4) == std::vector<int>{3, 4, 2, 5}));
    CHECK((to_host(d_vals, 4) == std::vector<float>{12.5f, 21.0f, 15.25f, 30.0f}));
    jit_sort_pairs(d_keys, d_vals, g, true);
    CHECK((to_host(d_keys, 4) == std::vector<int>{2, 3, 4, 5}));
    CHECK((to_host(d_vals, 4) == std::vector<float>{15.25f, 12.5f, 21.0f, 30.0f}));
    jit_sort_pairs(d_keys, d_vals, g, false);
    CHECK((to_host(d_keys, 4) == std::vector<int>{5, 4, 3, 2}));
    jit_sort_float(d_price, 6, false);
    CHECK((to_host(d_price, 6) == std::vector<float>{30.0f, 20.0f, 15.25f, 10.5f, 2.0f, 1.0f}));
    jit_sort_float(d_price, 6, true);
    CHECK((to_host(d_price, 6) == std::vector<float>{1.0f, 2.0f, 10.5f, 15.25f, 20.0f, 30.0f}));
    cudaFree(d_price); cudaFree(d_qty); cudaFree(d_vals); cudaFree(d_keys); cudaFree(d_count);
  }
  // ---- WarpDB on the bundled fixtures
  const std::vector<DataType> schema = {DataType::Float32, DataType::Int32};
  {
    WarpDB db("test.csv", schema);
    const auto r = db.query("price * quantity WHERE price > 10");           // BASELINE config 1
    CHECK((r == std::vector<float>{31.5f, 80.0f, 30.5f, 150.0f}));
    const auto f = db.query("price * 0.9 WHERE price > 20");
    CHECK(f.size() == 4 && f[0] == 0.0f && f[1] == 0.0f && f[2] == 0.0f && f[3] == 27.0f);
    CHECK((db.query("discount(price, 0.9)") == db.query("price * 0.9")));   // custom.cu UDF
    // tests/sql_features_test.cpp
    HostTable h = load_csv_to_host("test.csv", schema);
    const auto &hp = std::get<std::vector<float>>(h.columns[0].data);
    const auto &hq = std::get<std::vector<int32_t>>(h.columns[1].data);
    std::map<int, double> groups;
    for (size_t i = 0; i < hp.size(); ++i) groups[hq[i]] += hp[i];
    std::vector<float> expected;
    for (auto &kv : groups) expected.push_back(static_cast<float>(kv.second));
    const auto res = db.query_sql("SELECT SUM(price) FROM test GROUP BY quantity ORDER BY quantity ASC");
    CHECK(res.size() == expected.size());
    for (size_t i = 0; i < res.size(); ++i) CHECK(std::abs(res[i] - expected[i]) < 1e-5);
    const auto limited = db.query_sql("SELECT price FROM test ORDER BY price DESC LIMIT 2");
    std::vector<float> prices = hp;
    std::sort(prices.begin(), prices.end(), std::greater<float>());
    CHECK(limited.size() == 2 && limited[0] == prices[0] && limited[1] == prices[1]);
    const auto offset = db.query_sql("SELECT price FROM test ORDER BY price DESC OFFSET 1 LIMIT 2");
    CHECK(offset.size() == 2 && offset[0] == prices[1] && offset[1] == prices[2]);
    const auto having = db.query_sql("SELECT SUM(price) FROM test GROUP BY quantity HAVING SUM(price) > 15 ORDER BY quantity ASC");
    CHECK(having.size() == 3);
    // tests/having_distinct_test.cpp
    CHECK(db.query_sql("SELECT SUM(price) FROM test GROUP BY quantity HAVING COUNT(price) > 1").empty());
    const auto res2 = db.query_sql("SELECT DISTINCT quantity FROM test ORDER BY quantity DESC");
    CHECK(res2.size() == 4 && res2.front() > res2.back());
    // BASELINE configs 4 and 5 on the fixture
    CHECK((db.query_sql("SELECT SUM(price) FROM t GROUP BY quantity") == std::vector<float>{15.25f, 10.5f, 20.0f, 30.0f}));
    CHECK((db.query_sql("SELECT price FROM t ORDER BY price DESC LIMIT 5") == std::vector<float>{30.0f, 20.0f, 15.25f, 10.5f}));
    CHECK((db.query_sql("SELECT discount(price, 0.9) FROM t ORDER BY discount(price, 0.9) DESC LIMIT 5") ==
           std::vector<float>{30.0f * 0.9f, 20.0f * 0.9f, 15.25f * 0.9f, 10.5f * 0.9f}));
    CHECK((db.query_sql("SELECT price * 0.9 FROM t WHERE price > 20") == std::vector<float>{27.0f}));
    // multi-GPU entry points (run on however many devices are visible)
    CHECK((db.query_multi_gpu("price * quantity WHERE price > 10") == r));
    CHECK((WarpDB::query_multi_gpu_csv("test.csv", "price + 1", 3) == std::vector<float>{11.5f, 21.0f, 16.25f, 31.0f}));
    // error contract
    bool threw = false;
    try { db.query("foo + 1"); } catch (const std::exception &e) { threw = std::string(e.what()) == "Unknown column: foo"; }
    CHECK(threw);
    threw = false;
    try { db.query("1 2"); } catch (const std::exception &e) { threw = std::string(e.what()) == "Failed to parse expression: Unexpected tokens remaining: 2"; }
    CHECK(threw);
    threw = false;
    try { db.query(""); } catch (const std::exception &e) { threw = std::string(e.what()) == "Empty query expression"; }
    CHECK(threw);
    // Arrow export
    ArrowArray arr; ArrowSchema sch;
    db.query_arrow("price + 1", &arr, &sch);
    CHECK(arr.length == 4 && arr.n_buffers == 2 && std::string(sch.format) == "f" && std::string(sch.name) == "result");
    CHECK(static_cast<const float *>(arr.buffers[1])[3] == 31.0f);
    arr.release(&arr); sch.release(&sch);
    // device-side Arrow export: the result stays in HBM
    ArrowDeviceArray darr; ArrowSchema dsch;
    db.query_arrow_device("price * quantity WHERE price > 10", &darr, &dsch);
    CHECK(darr.device_type == ARROW_DEVICE_CUDA && darr.array.length == 4 && std::string(dsch.format) == "f");
    CHECK((to_host(static_cast<const float *>(darr.array.buffers[1]), 4) == r));
    darr.array.release(&darr.array);
    // optimizer: a condition no row can satisfy is pruned from the table statistics
    Table t = db.table();
    execute_query_optimized("price", "price > 1000", t);   // prints "[Optimizer] Filter eliminates all rows."
    execute_query_optimized("price", "price > 5", t);      // always true: runs without the filter
  }
  // ---- tests/extended_types_test.cpp
  {
    WarpDB db("extended.csv", {DataType::Float32, DataType::Int32, DataType::Float32});
    const auto res = db.query("price * discount");
    CHECK(res.size() == 4 && static_cast<int>(res[0]) == 1);
  }
  {
    WarpDB db("test.json");
    CHECK((db.query("price * quantity") == std::vector<float>{31.5f, 80.0f, 30.5f, 150.0f}));
  }
  std::printf("ALL HOST-MIRROR TESTS PASSED\n");
  return 0;
}
