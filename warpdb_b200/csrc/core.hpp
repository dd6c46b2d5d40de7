// core.hpp -- internal declarations shared by the translation units of libwarpcore.so.
// Not part of the C ABI (that is include/warpcore.h).
#pragma once
#include <cuda_runtime.h>
#include <nvrtc.h>

#include <cstdarg>
#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "warpcore.h"

namespace wdb {

// ---- errors ------------------------------------------------------------------------------------
int fail(const char *fmt, ...);            // records the thread-local message, returns 1
void set_error(const std::string &msg);
#define WDB_CUDA(call)                                                                              \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess) return ::wdb::fail("CUDA error: %s (%s)", cudaGetErrorString(e__), #call); \
  } while (0)

// stream-ordered scratch that is released on every exit path (early error returns included)
struct Scratch {
  void *p = nullptr;
  cudaStream_t s = nullptr;
  Scratch() = default;
  Scratch(const Scratch &) = delete;
  Scratch &operator=(const Scratch &) = delete;
  ~Scratch() { release(); }
  cudaError_t alloc(size_t bytes, cudaStream_t stream) { release(); s = stream; return cudaMallocAsync(&p, bytes, stream); }
  void release() { if (p) { cudaFreeAsync(p, s); p = nullptr; } }
  template <class T> T *as() const { return static_cast<T *>(p); }
};

// ---- options / stats ---------------------------------------------------------------------------
int64_t opt(const char *key, int64_t dflt);
struct Stats { int64_t compiled = 0, hits = 0, launches = 0; double last_compile_ms = 0; };
Stats &stats();

// ---- per-device state --------------------------------------------------------------------------
struct Kernel {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t fn = nullptr;
  int regs = 0;
  int max_ctas_per_sm = 0;   // occupancy at the launch configuration it was built for
};
struct Device {
  int id = -1;
  bool ready = false;
  int cc_major = 0, cc_minor = 0, num_sms = 0;
  size_t smem_optin = 0;
  std::string arch;          // "sm_100a"
  std::mutex mu;
  std::unordered_map<std::string, Kernel> cache;
};
int get_device(int id, Device **out);                 // initialises on first use; makes it current

// ---- code generation ---------------------------------------------------------------------------
struct UsedCol { int table_index; std::string name; int dtype; };
struct GenFn { std::string name, ret, text; };        // wdb_fn_<name>, macro WDB_<NAME>(R, j)
struct GenSpec {
  std::string kind;                                   // "project", "compact", "group", "topk"
  std::vector<UsedCol> used;
  std::vector<GenFn> fns;
  std::vector<std::pair<std::string, int64_t>> defines;
  std::vector<const char *> bodies;                   // kernel template texts, in order
};
std::vector<UsedCol> find_used_columns(const wdb_col_t *cols, int ncols, const std::vector<std::string> &texts);
int dtype_size(int dtype);
std::string gen_source(const GenSpec &spec);
// NVRTC -> CUBIN for `arch`; on failure prints the log to stderr and sets "Kernel compilation failed."
int compile_to_cubin(const std::string &source, const std::string &name, const std::string &arch, std::string *cubin);
// cached: compile (if needed), load and resolve `entry` on device d
int get_kernel(Device *d, const std::string &source, const std::string &name, const char *entry, Kernel *out);
int launch(const Kernel &k, unsigned grid, unsigned block, size_t smem, cudaStream_t stream, void **args);
bool all_aligned(const std::vector<UsedCol> &used, const wdb_col_t *cols, const void *out, size_t align);
std::string udf_source();

// embedded kernel template texts (generated from kernels/*.cuh by build.py)
extern const char *const k_src_prelude;
extern const char *const k_src_project;
extern const char *const k_src_compact;
extern const char *const k_src_group_table;
extern const char *const k_src_group;
extern const char *const k_src_keyrange;
extern const char *const k_src_topk;

}  // namespace wdb
