// warpcore.cu -- library state, NVRTC kernel cache, code generator and the fused
// filter/project entry point of libwarpcore.so (C ABI: include/warpcore.h).
//
// Replaces src/jit.cpp:48-174 of the reference (jit_compile_and_launch): instead of compiling a
// one-thread-per-row kernel to PTX on every call inside a throw-away context, the expression
// strings are pasted into hand-written sm_100a kernel templates (kernels/*.cuh), compiled once
// per (template, expression, column types, tuning) with NVRTC straight to an sm_100a CUBIN, and
// cached per device in the primary context.
#include "core.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <iostream>
#include <sstream>

namespace wdb {

// ------------------------------------------------------------------------------------------------
// errors, options, stats
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
int fail(const char *fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_error = buf;
  return 1;
}

static std::mutex g_mu;
static std::map<std::string, int64_t> g_opts;
static std::string g_udf;
static Stats g_stats;
Stats &stats() { return g_stats; }
int64_t opt(const char *key, int64_t dflt) {
  std::lock_guard<std::mutex> l(g_mu);
  auto it = g_opts.find(key);
  return it == g_opts.end() ? dflt : it->second;
}
std::string udf_source() {
  std::lock_guard<std::mutex> l(g_mu);
  return g_udf;
}

// ------------------------------------------------------------------------------------------------
// devices
// ------------------------------------------------------------------------------------------------
static const int kMaxDevices = 64;
static Device g_dev[kMaxDevices];

int get_device(int id, Device **out) {
  if (id < 0 || id >= kMaxDevices) return fail("invalid device id %d", id);
  Device *d = &g_dev[id];
  if (!d->ready) {
    std::lock_guard<std::mutex> l(d->mu);
    if (!d->ready) {
      int count = 0;
      cudaError_t e = cudaGetDeviceCount(&count);
      if (e != cudaSuccess || count == 0)
        return fail("CUDA error: no CUDA device available (%s); warpcore has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
      if (id >= count) return fail("invalid device id %d (%d devices)", id, count);
      WDB_CUDA(cudaSetDevice(id));
      WDB_CUDA(cudaFree(0));  // retains the primary context (vs cuCtxCreate per call: src/jit.cpp:155)
      cudaDeviceProp p;
      WDB_CUDA(cudaGetDeviceProperties(&p, id));
      d->id = id;
      d->cc_major = p.major;
      d->cc_minor = p.minor;
      d->num_sms = p.multiProcessorCount;
      d->smem_optin = p.sharedMemPerBlockOptin;
      if (p.major != 10)
        return fail("warpcore targets Blackwell sm_100a (B200); device %d is sm_%d%d", id, p.major, p.minor);
      d->arch = "sm_" + std::to_string(p.major) + std::to_string(p.minor) + "a";
      // scratch buffers come from the stream-ordered allocator; keep freed memory cached instead of
      // returning it to the driver at every synchronisation (the default release threshold is 0)
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, id) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
      d->ready = true;
    }
  }
  WDB_CUDA(cudaSetDevice(id));
  *out = d;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// code generation
// ------------------------------------------------------------------------------------------------
int dtype_size(int dtype) {
  switch (dtype) {
  case WDB_INT32: case WDB_FLOAT32: return 4;
  case WDB_INT64: case WDB_FLOAT64: return 8;
  }
  return 0;
}
static const char *dtype_cuda(int dtype) {  // src/jit.cpp:31-45
  switch (dtype) {
  case WDB_INT32: return "int";
  case WDB_INT64: return "long long";
  case WDB_FLOAT32: return "float";
  case WDB_FLOAT64: return "double";
  }
  return "void*";
}
// A column is read only if one of the expression strings names it (the reference passes every
// column of the table as a kernel parameter: src/jit.cpp:75-79).
std::vector<UsedCol> find_used_columns(const wdb_col_t *cols, int ncols, const std::vector<std::string> &texts) {
  std::vector<UsedCol> used;
  std::vector<char> seen(ncols, 0);
  for (const auto &t : texts) {
    size_t i = 0;
    while (i < t.size()) {
      if (isalpha((unsigned char)t[i]) || t[i] == '_') {
        size_t s = i;
        while (i < t.size() && (isalnum((unsigned char)t[i]) || t[i] == '_')) ++i;
        std::string id = t.substr(s, i - s);
        for (int c = 0; c < ncols; ++c)
          if (!seen[c] && cols[c].name && id == cols[c].name) {
            seen[c] = 1;
          }
      } else if (isdigit((unsigned char)t[i]) || t[i] == '.') {
        // skip numeric literals including their suffix (10.0f) so 'f' is not taken for a name
        while (i < t.size() && (isalnum((unsigned char)t[i]) || t[i] == '.')) ++i;
      } else
        ++i;
    }
  }
  for (int c = 0; c < ncols; ++c)
    if (seen[c]) used.push_back({c, cols[c].name, cols[c].dtype});
  return used;
}

bool all_aligned(const std::vector<UsedCol> &used, const wdb_col_t *cols, const void *out, size_t align) {
  if (out && ((uintptr_t)out & (align - 1))) return false;
  for (const auto &u : used)
    if ((uintptr_t)cols[u.table_index].dptr & (align - 1)) return false;
  return true;
}

static std::string upper(std::string s) {
  for (auto &c : s) c = (char)toupper((unsigned char)c);
  return s;
}

std::string gen_source(const GenSpec &spec) {
  std::ostringstream o;
  o << "// generated by warpcore: kind=" << spec.kind << "\n";
  for (const auto &d : spec.defines) o << "#define " << d.first << " " << d.second << "\n";
  o << k_src_prelude << "\n";
  o << "// ---- UDF source (custom.cu; src/jit.cpp:65-73) ----\n" << udf_source() << "\n// ---- end UDF ----\n";
  const size_t nu = spec.used.size();
  for (size_t k = 0; k < nu; ++k) o << "typedef " << dtype_cuda(spec.used[k].dtype) << " wdb_t" << k << ";\n";
  o << "#define WDB_NUSED " << nu << "\n";
  // parameter block (an array of pointers as far as the host is concerned)
  o << "struct wdb_cols {";
  for (size_t k = 0; k < nu; ++k) o << " const wdb_t" << k << " *__restrict__ c" << k << ";";
  if (!nu) o << " const void *none;";
  o << " };\n";
  for (const char *sname : {"wdb_rows", "wdb_rows4s"}) {
    o << "struct " << sname << " {";
    for (size_t k = 0; k < nu; ++k) o << " wdb_t" << k << " c" << k << (std::string(sname) == "wdb_rows" ? "[WDB_VEC];" : "[4];");
    if (!nu) o << " int none;";
    o << " };\n";
  }
  o << "__device__ __forceinline__ void wdb_load_rows(const wdb_cols &C, i64 row, wdb_rows &R) {";
  for (size_t k = 0; k < nu; ++k) o << " wdb_load_vec(C.c" << k << ", row, R.c" << k << ");";
  o << " }\n";
  o << "__device__ __forceinline__ void wdb_prefetch_rows(const wdb_cols &C, i64 row) {";
  for (size_t k = 0; k < nu; ++k) o << " wdb_prefetch_l2(C.c" << k << " + row);";
  o << " }\n";
  o << "template <int H> __device__ __forceinline__ void wdb_load_rows_h(const wdb_cols &C, i64 row, wdb_rows &R) {";
  for (size_t k = 0; k < nu; ++k) o << " wdb_load_vec_h<H>(C.c" << k << ", row, R.c" << k << ");";
  o << " }\n";
  o << "template <class RowsT> __device__ __forceinline__ void wdb_load_row1(const wdb_cols &C, i64 row, RowsT &R, int j) {";
  for (size_t k = 0; k < nu; ++k) o << " R.c" << k << "[j] = __ldg(C.c" << k << " + row);";
  o << " }\n";
  // shared-memory tile layout of the bulk variant: column k at byte offset WDB_TILE * (sum of sizes before it)
  size_t off = 0;
  for (size_t k = 0; k < nu; ++k) {
    o << "#define WDB_COL_OFF" << k << " ((size_t)WDB_TILE * " << off << ")\n";
    off += dtype_size(spec.used[k].dtype);
  }
  o << "#define WDB_ROW_BYTES " << off << "\n";
  o << "#ifdef WDB_TILE\n";
  o << "#define WDB_IN_BYTES ((u32)(WDB_TILE * WDB_ROW_BYTES))\n";
  o << "#define WDB_STAGE_BYTES ((size_t)WDB_TILE * (WDB_ROW_BYTES + 4))\n";
  o << "__device__ __forceinline__ void wdb_bulk_load_tile(const wdb_cols &C, i64 row0, unsigned char *sb, u64 *bar) {";
  for (size_t k = 0; k < nu; ++k)
    o << " wdb_bulk_g2s(sb + WDB_COL_OFF" << k << ", C.c" << k << " + row0, (u32)(WDB_TILE * sizeof(wdb_t" << k << ")), bar);";
  o << " }\n";
  o << "template <class T> __device__ __forceinline__ void wdb_lds4(const unsigned char *base, int r, T (&o)[4]) {\n"
       "  if constexpr (sizeof(T) == 4) { const uint4 v = *reinterpret_cast<const uint4 *>(base + (size_t)r * 4);\n"
       "    o[0] = wdb_from_bits32<T>(v.x); o[1] = wdb_from_bits32<T>(v.y); o[2] = wdb_from_bits32<T>(v.z); o[3] = wdb_from_bits32<T>(v.w); }\n"
       "  else { const uint4 a = *reinterpret_cast<const uint4 *>(base + (size_t)r * 8), b = *reinterpret_cast<const uint4 *>(base + (size_t)r * 8 + 16);\n"
       "    o[0] = wdb_from_bits64<T>(a.x, a.y); o[1] = wdb_from_bits64<T>(a.z, a.w); o[2] = wdb_from_bits64<T>(b.x, b.y); o[3] = wdb_from_bits64<T>(b.z, b.w); }\n"
       "}\n";
  o << "__device__ __forceinline__ void wdb_lds_rows(const unsigned char *sb, int r, wdb_rows4s &R) {";
  for (size_t k = 0; k < nu; ++k) o << " wdb_lds4<wdb_t" << k << ">(sb + WDB_COL_OFF" << k << ", r, R.c" << k << ");";
  o << " }\n";
  o << "#endif\n";
  // row functions
  std::string params, args;
  for (size_t k = 0; k < nu; ++k) {
    params += (k ? ", " : "") + std::string("wdb_cell<wdb_t") + std::to_string(k) + "> " + spec.used[k].name;
    args += (k ? ", " : "") + std::string("wdb_cell<wdb_t") + std::to_string(k) + ">{(R).c" + std::to_string(k) + "[j]}";
  }
  if (!nu) { params = "int wdb_none"; args = "0"; }
  o << "#define WDB_ROW_ARGS(R, j) " << args << "\n";
  for (const auto &f : spec.fns) {
    o << "__device__ __forceinline__ " << f.ret << " wdb_fn_" << f.name << "(" << params
      << ") { const wdb_idx_t idx = {}; (void)idx; return (" << f.ret << ")(" << f.text << "); }\n";
    o << "#define WDB_" << upper(f.name) << "(R, j) wdb_fn_" << f.name << "(WDB_ROW_ARGS(R, j))\n";
    o << "#define WDB_" << upper(f.name) << "4(R, j) wdb_fn_" << f.name << "(WDB_ROW_ARGS(R, j))\n";
  }
  for (const char *b : spec.bodies) o << b << "\n";
  return o.str();
}

// ------------------------------------------------------------------------------------------------
// NVRTC
// ------------------------------------------------------------------------------------------------
// NVRTC is loaded with dlopen from the CUDA 12.9 toolkit by absolute path: a process that imported
// PyTorch already has torch's own libnvrtc.so.12 (12.8) mapped under the same soname, and that one
// predates the 256-bit ld/st PTX (ptxas: "Illegal vector size: 8").
struct Nvrtc {
  void *h = nullptr;
  int major = 0, minor = 0;
  nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram *) = nullptr;
  const char *(*GetErrorString)(nvrtcResult) = nullptr;
  nvrtcResult (*Version)(int *, int *) = nullptr;
};
static Nvrtc g_nvrtc;
static int load_nvrtc() {
  static std::mutex mu;
  std::lock_guard<std::mutex> l(mu);
  if (g_nvrtc.h) return 0;
  std::vector<std::string> cand;
  if (const char *e = getenv("WARPDB_NVRTC")) cand.push_back(e);
  if (const char *e = getenv("CUDA_HOME")) cand.push_back(std::string(e) + "/lib64/libnvrtc.so.12");
  cand.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
  cand.push_back("/usr/local/cuda-12.9/lib64/libnvrtc.so.12");
  cand.push_back("libnvrtc.so.12");
  std::string tried;
  for (const auto &c : cand) {
    void *h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) { tried += c + " "; continue; }
    Nvrtc n;
    n.h = h;
#define WDB_SYM(f) *(void **)(&n.f) = dlsym(h, "nvrtc" #f)
    WDB_SYM(CreateProgram); WDB_SYM(CompileProgram); WDB_SYM(GetProgramLogSize); WDB_SYM(GetProgramLog);
    WDB_SYM(GetCUBINSize); WDB_SYM(GetCUBIN); WDB_SYM(DestroyProgram); WDB_SYM(GetErrorString); WDB_SYM(Version);
#undef WDB_SYM
    if (!n.CreateProgram || !n.CompileProgram || !n.GetCUBIN || !n.Version) { tried += c + "(symbols) "; dlclose(h); continue; }
    n.Version(&n.major, &n.minor);
    if (n.major < 12 || (n.major == 12 && n.minor < 9)) {  // sm_100a + 256-bit ld/st need the 12.9 compiler
      tried += c + "(" + std::to_string(n.major) + "." + std::to_string(n.minor) + " < 12.9) ";
      dlclose(h);
      continue;
    }
    g_nvrtc = n;
    return 0;
  }
  return fail("NVRTC error: no NVRTC >= 12.9 found (tried: %s)", tried.c_str());
}

int compile_to_cubin(const std::string &source, const std::string &name, const std::string &arch, std::string *cubin) {
  if (load_nvrtc()) return 1;
  const Nvrtc &N = g_nvrtc;
  nvrtcProgram prog = nullptr;
  nvrtcResult r = N.CreateProgram(&prog, source.c_str(), name.c_str(), 0, nullptr, nullptr);
  if (r != NVRTC_SUCCESS) return fail("NVRTC error: %s", N.GetErrorString(r));
  std::string arch_flag = "--gpu-architecture=" + arch;
  // The reference passes only the architecture (src/jit.cpp:114-117), i.e. NVRTC defaults:
  // --fmad=true, IEEE division and sqrt, no flush-to-zero.  We add the language level, line info
  // for ncu's source page and silence the unused-variable remarks of generated code.
  const char *opts[] = {arch_flag.c_str(), "--std=c++17", "-lineinfo", "-w"};
  r = N.CompileProgram(prog, 4, opts);
  if (r != NVRTC_SUCCESS) {
    size_t n = 0;
    N.GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    N.GetProgramLog(prog, &log[0]);
    std::cerr << "NVRTC Compile Log:\n" << log << "\n";  // src/jit.cpp:123-125
    N.DestroyProgram(&prog);
    if (getenv("WARPDB_DUMP_SOURCE")) std::cerr << source << "\n";
    return fail("Kernel compilation failed.");            // src/jit.cpp:128
  }
  size_t n = 0;
  r = N.GetCUBINSize(prog, &n);
  if (r != NVRTC_SUCCESS || n == 0) {
    N.DestroyProgram(&prog);
    return fail("NVRTC error: no CUBIN produced for %s (%s)", arch.c_str(), N.GetErrorString(r));
  }
  cubin->resize(n);
  N.GetCUBIN(prog, &(*cubin)[0]);
  N.DestroyProgram(&prog);
  return 0;
}

int get_kernel(Device *d, const std::string &source, const std::string &name, const char *entry, Kernel *out) {
  std::string key = std::string(entry) + "\n" + source;
  {
    std::lock_guard<std::mutex> l(d->mu);
    auto it = d->cache.find(key);
    if (it != d->cache.end()) {
      *out = it->second;
      g_stats.hits++;
      return 0;
    }
  }
  auto t0 = std::chrono::steady_clock::now();
  std::string cubin;
  if (compile_to_cubin(source, name, d->arch, &cubin)) return 1;
  Kernel k;
  WDB_CUDA(cudaLibraryLoadData(&k.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  WDB_CUDA(cudaLibraryGetKernel(&k.fn, k.lib, entry));
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, (const void *)k.fn) == cudaSuccess) k.regs = fa.numRegs;
  else cudaGetLastError();
  auto t1 = std::chrono::steady_clock::now();
  g_stats.compiled++;
  g_stats.last_compile_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  {
    std::lock_guard<std::mutex> l(d->mu);
    d->cache[key] = k;
  }
  *out = k;
  return 0;
}

int launch(const Kernel &k, unsigned grid, unsigned block, size_t smem, cudaStream_t stream, void **args) {
  if (grid == 0) return 0;
  WDB_CUDA(cudaLaunchKernel((const void *)k.fn, dim3(grid), dim3(block), args, smem, stream));
  g_stats.launches++;
  return 0;
}

}  // namespace wdb

using namespace wdb;

// ------------------------------------------------------------------------------------------------
// synthetic column generators (bit-identical to oracle/wdb_oracle.c: orc_synth_*)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long wdb_mix64(unsigned long long seed, unsigned long long row) {
  unsigned long long z = seed + (row + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) wdb_synth_f32_kernel(float *__restrict__ out, long long n, unsigned long long seed,
                                                            float lo, float span, long long row0) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float u = (float)(wdb_mix64(seed, (unsigned long long)(row0 + i)) >> 40) * 0x1p-24f;
    out[i] = __fmaf_rn(u, span, lo);
  }
}
__global__ void __launch_bounds__(256) wdb_synth_i32_kernel(int *__restrict__ out, long long n, unsigned long long seed,
                                                            int lo, unsigned long long range, long long row0) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long h = wdb_mix64(seed, (unsigned long long)(row0 + i)) >> 32;
    out[i] = (int)((long long)lo + (long long)((h * range) >> 32));
  }
}

// ------------------------------------------------------------------------------------------------
// filter / project
// ------------------------------------------------------------------------------------------------
namespace wdb {
int run_compact(Device *d, cudaStream_t stream, const wdb_col_t *cols, int ncols, const char *expr, const char *key_expr,
                const char *cond, float *d_out, float *d_out2, int64_t n, int64_t *d_count, int64_t *h_count);
int run_project(Device *d, cudaStream_t stream, const wdb_col_t *cols, int ncols, const char *expr, const char *cond,
                float *d_out, int64_t n, int mode, const unsigned char *zmask = nullptr, int zshift = 0);

struct ProjectPlan {
  GenSpec spec;
  int variant, block, unroll, vec, tile, stages;
  bool aligned;
  const char *entry;
};

static int plan_project(const wdb_col_t *cols, int ncols, const char *expr, const char *cond, const float *d_out,
                        int mode, ProjectPlan *p, bool check_alignment, bool prune = false) {
  const bool has_cond = cond && *cond;
  p->spec.kind = "project";
  p->spec.used = find_used_columns(cols, ncols, {expr, has_cond ? cond : ""});
  for (const auto &u : p->spec.used)
    if (dtype_size(u.dtype) == 0) return fail("column %s has a non-numeric type and cannot be read on the GPU", u.name.c_str());
  p->variant = (int)opt("project.variant", 0);
  p->block = (int)opt("project.block", 512);   // defaults from profiles/r01_sweep_project_1e9.jsonl
  p->unroll = (int)opt("project.unroll", 2);
  p->vec = (int)opt("project.vec", 8);
  p->tile = (int)opt("project.tile", 4096);
  p->stages = (int)opt("project.stages", 3);
  if (p->vec != 4 && p->vec != 8) return fail("project.vec must be 4 or 8");
  if (p->block < 32 || p->block > 1024 || (p->block & 31)) return fail("project.block must be a multiple of 32 in [32,1024]");
  if (p->unroll < 1 || p->unroll > 16) return fail("project.unroll must be in [1,16]");
  // dense-untouched output cannot leave through a bulk store
  if (p->variant == 2 && has_cond && mode != WDB_DENSE_ZERO) p->variant = 0;
  if (prune) p->variant = 0;
  if (p->variant == 2) {
    if (p->stages < 2 || p->stages > 8) return fail("project.stages must be in [2,8]");
    if (p->tile % (4 * p->block)) return fail("project.tile must be a multiple of 4*project.block");
  }
  p->aligned = !check_alignment || all_aligned(p->spec.used, cols, d_out, (size_t)p->vec * 4);
  if (!p->aligned && p->variant == 2) p->variant = 0;
  auto &D = p->spec.defines;
  D.push_back({"WDB_VEC", p->vec});
  D.push_back({"WDB_ALIGNED", p->aligned ? 1 : 0});
  D.push_back({"WDB_LD_HINT", opt("project.ld_hint", 0)});
  D.push_back({"WDB_ST_HINT", opt("project.st_hint", p->vec == 8 ? 3 : 0)});
  D.push_back({"WDB_BLOCK", p->block});
  D.push_back({"WDB_UNROLL", p->unroll});
  D.push_back({"WDB_MODE", mode == WDB_DENSE_ZERO ? 2 : 0});
  D.push_back({"WDB_HAS_COND", has_cond ? 1 : 0});
  D.push_back({"WDB_BULK", p->variant == 2 ? 1 : 0});
  D.push_back({"WDB_PRUNE", prune ? 1 : 0});
  if (p->variant == 2) {
    D.push_back({"WDB_TILE", p->tile});
    D.push_back({"WDB_STAGES", p->stages});
  }
  p->spec.fns.push_back({"expr", "float", expr});
  if (has_cond) p->spec.fns.push_back({"cond", "bool", cond});
  p->spec.bodies = {k_src_project};
  p->entry = p->variant == 2 ? "wdb_project_bulk" : "wdb_project";
  return 0;
}

int run_project(Device *d, cudaStream_t stream, const wdb_col_t *cols, int ncols, const char *expr,
                       const char *cond, float *d_out, int64_t n, int mode, const unsigned char *zmask, int zshift) {
  ProjectPlan p;
  if (plan_project(cols, ncols, expr, cond, d_out, mode, &p, true, zmask != nullptr)) return 1;
  Kernel k;
  if (get_kernel(d, gen_source(p.spec), "wdb_project.cu", p.entry, &k)) return 1;
  if (n == 0) return 0;
  std::vector<const void *> ptrs;
  for (const auto &u : p.spec.used) ptrs.push_back(cols[u.table_index].dptr);
  if (ptrs.empty()) ptrs.push_back(nullptr);
  long long nn = n;
  void *args[] = {ptrs.data(), &d_out, &nn, &zmask, &zshift};   // the last two only exist in pruned instantiations
  if (p.variant == 2) {
    size_t row_bytes = 4;
    for (const auto &u : p.spec.used) row_bytes += dtype_size(u.dtype);
    size_t smem = 128 + (size_t)p.stages * p.tile * row_bytes;
    if (smem > d->smem_optin) return fail("project.tile*stages needs %zu B of shared memory (max %zu)", smem, d->smem_optin);
    WDB_CUDA(cudaFuncSetAttribute((const void *)k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t ntiles = n / p.tile;
    unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)d->num_sms * opt("project.ctas_per_sm", 1)));
    return launch(k, grid, p.block, smem, stream, args);
  }
  int64_t tile_rows = (int64_t)p.block * p.unroll * p.vec;
  int64_t ntiles = (n + tile_rows - 1) / tile_rows;
  int64_t grid = ntiles;
  if (p.variant == 1) grid = std::min<int64_t>(ntiles, (int64_t)d->num_sms * opt("project.ctas_per_sm", 8));
  grid = std::max<int64_t>(1, std::min<int64_t>(grid, 0x7fffffff));
  return launch(k, (unsigned)grid, p.block, 0, stream, args);
}

// source a call of the given kind would compile (no device needed): used by wdb_debug_compile
int gen_compact_source(const wdb_col_t *cols, int ncols, const char *expr, const char *expr2, const char *cond,
                       bool assume_aligned, std::string *src);
int gen_group_source(const wdb_col_t *cols, int ncols, const char *val, const char *key, const char *cond, int agg, std::string *src);
int gen_keyrange_source(const wdb_col_t *cols, int ncols, const char *key, std::string *src);
int gen_topk_source(const wdb_col_t *cols, int ncols, const char *key, const char *val, const char *cond, int desc, std::string *src);
int gen_kernel_source(const std::string &kind, const wdb_col_t *cols, int ncols, const char *a, const char *b,
                      const char *cond, int mode, std::string *src, std::string *name) {
  if (!a || !*a) return fail("empty expression");
  if (kind == "project" || kind == "filter") {
    ProjectPlan p;
    if (plan_project(cols, ncols, a, cond, nullptr, mode, &p, false)) return 1;
    *src = gen_source(p.spec);
    *name = "wdb_project.cu";
    return 0;
  }
  if (kind == "compact") {
    *name = "wdb_compact.cu";
    return gen_compact_source(cols, ncols, a, b, cond, true, src);
  }
  if (kind == "keyrange") {
    *name = "wdb_keyrange.cu";
    return gen_keyrange_source(cols, ncols, a, src);
  }
  if (kind == "group") {
    *name = "wdb_group.cu";
    return gen_group_source(cols, ncols, a, b, cond, mode, src);   // mode = aggregation type
  }
  if (kind == "topk") {
    *name = "wdb_topk.cu";
    return gen_topk_source(cols, ncols, a, b, cond, mode, src);    // mode = descending
  }
  return fail("wdb_debug_compile: unknown kernel kind '%s'", kind.c_str());
}
}  // namespace wdb

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int wdb_abi_version(void) { return WDB_ABI_VERSION; }
const char *wdb_last_error(void) { return g_error.c_str(); }

int wdb_init(int device) {
  Device *d;
  return get_device(device, &d);
}
int wdb_shutdown(void) {
  for (int i = 0; i < kMaxDevices; ++i) {
    Device *d = &g_dev[i];
    if (!d->ready) continue;
    std::lock_guard<std::mutex> l(d->mu);
    cudaSetDevice(i);
    cudaDeviceSynchronize();
    for (auto &kv : d->cache)
      if (kv.second.lib) cudaLibraryUnload(kv.second.lib);
    d->cache.clear();
    d->ready = false;
  }
  return 0;
}
int wdb_device_count(int *out) {
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) { cudaGetLastError(); c = 0; }
  *out = c;
  return 0;
}
int wdb_set_udf_source(const char *src) {
  std::lock_guard<std::mutex> l(g_mu);
  g_udf = src ? src : "";
  return 0;
}
int wdb_set_option(const char *key, int64_t value) {
  if (!key) return fail("null option key");
  std::lock_guard<std::mutex> l(g_mu);
  if (value == INT64_MIN) g_opts.erase(key);   // back to the built-in default
  else g_opts[key] = value;
  return 0;
}
int wdb_get_option(const char *key, int64_t *value) {
  std::lock_guard<std::mutex> l(g_mu);
  auto it = g_opts.find(key ? key : "");
  if (it == g_opts.end()) return fail("option %s is not set", key ? key : "(null)");
  *value = it->second;
  return 0;
}
int wdb_get_stats(wdb_stats_t *out) {
  out->kernels_compiled = g_stats.compiled;
  out->cache_hits = g_stats.hits;
  out->launches = g_stats.launches;
  out->last_compile_ms = g_stats.last_compile_ms;
  return 0;
}

int wdb_project_filter(int device, void *stream, const wdb_col_t *cols, int ncols, const char *expr, const char *cond,
                       float *d_out, int64_t n, int mode, int64_t *d_count, int64_t *h_count) {
  if (!expr || !*expr) return fail("empty expression");
  if (n < 0) return fail("negative row count");
  if (mode != WDB_DENSE && mode != WDB_COMPACT && mode != WDB_DENSE_ZERO) return fail("invalid mode %d", mode);
  Device *d;
  if (get_device(device, &d)) return 1;
  cudaStream_t s = (cudaStream_t)stream;
  const bool has_cond = cond && *cond;
  if (mode == WDB_COMPACT && has_cond)
    return run_compact(d, s, cols, ncols, expr, nullptr, cond, d_out, nullptr, n, d_count, h_count);
  if (run_project(d, s, cols, ncols, expr, cond, d_out, n, mode == WDB_COMPACT ? WDB_DENSE : mode)) return 1;
  long long nn = n;
  if (d_count) WDB_CUDA(cudaMemcpyAsync(d_count, &nn, sizeof nn, cudaMemcpyHostToDevice, s));
  if (h_count) {
    *h_count = n;
    WDB_CUDA(cudaStreamSynchronize(s));
  }
  return 0;
}

int wdb_shard_range(int64_t n, int ndev, int dev, int64_t *start, int64_t *end) {  // src/multi_gpu_utils.cpp:24-31
  if (ndev < 1 || dev < 0) return fail("invalid shard request");
  int64_t chunk = (n + ndev - 1) / ndev;
  int64_t s = std::min<int64_t>((int64_t)dev * chunk, n), e = std::min<int64_t>(s + chunk, n);
  *start = s;
  *end = e;
  return 0;
}

int wdb_synth_f32(int device, void *stream, float *d_out, int64_t n, uint64_t seed, float lo, float hi, int64_t row0) {
  Device *d;
  if (get_device(device, &d)) return 1;
  if (n <= 0) return 0;
  unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)d->num_sms * 16);
  wdb_synth_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_out, n, seed, lo, hi - lo, row0);
  g_stats.launches++;
  WDB_CUDA(cudaGetLastError());
  return 0;
}
int wdb_synth_i32(int device, void *stream, int32_t *d_out, int64_t n, uint64_t seed, int32_t lo, int32_t hi_excl, int64_t row0) {
  Device *d;
  if (get_device(device, &d)) return 1;
  if (n <= 0) return 0;
  if (hi_excl <= lo) return fail("empty integer range");
  unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)d->num_sms * 16);
  wdb_synth_i32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_out, n, seed, lo, (unsigned long long)((long long)hi_excl - (long long)lo), row0);
  g_stats.launches++;
  WDB_CUDA(cudaGetLastError());
  return 0;
}

void wdb_free(void *p) { free(p); }

int wdb_debug_compile(const char *kind, const wdb_col_t *cols, int ncols, const char *expr_a, const char *expr_b,
                      const char *cond, int mode, const char *arch, char **out_source, void **out_cubin,
                      size_t *out_cubin_size) {
  std::string src, name;
  if (gen_kernel_source(kind ? kind : "", cols, ncols, expr_a, expr_b, cond, mode, &src, &name)) return 1;
  if (out_source) *out_source = strdup(src.c_str());
  if (out_cubin) {
    std::string cubin;
    if (compile_to_cubin(src, name, arch && *arch ? arch : "sm_100a", &cubin)) return 1;
    *out_cubin = malloc(cubin.size());
    memcpy(*out_cubin, cubin.data(), cubin.size());
    if (out_cubin_size) *out_cubin_size = cubin.size();
  }
  return 0;
}

}  // extern "C"
