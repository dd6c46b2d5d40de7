// ref_jit_shim.cpp -- C ABI + symbol interposition around the REFERENCE's NVRTC execution core
// (src/jit.cpp, src/multi_gpu_utils.cpp, src/csv_loader.cpp are compiled unmodified from
// /root/reference by oracle/Makefile into oracle/_ref/libref_jit.so).  TEST/BENCH INFRASTRUCTURE:
// this is the "reference's own NVRTC build on the same box" baseline, never a product path.
//
// Interposition (link-time, inside this DSO only; the reference's code is not edited):
//   * cuLaunchKernel   -- brackets the reference's launch (src/jit.cpp:169) with CUDA events so
//                         the kernel time can be reported apart from NVRTC/ctx overhead.
//   * cuCtxCreate_v2 / cuCtxDestroy_v2 -- when ref_set_primary_ctx(1) was called, the context
//                         the reference creates per call (src/jit.cpp:155) is replaced by the
//                         device's primary context: pointers allocated through the runtime
//                         (cudaMalloc / torch) live there and are not addressable from a fresh
//                         context.  Mode 0 leaves the reference's behaviour as is.
#include <cuda.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "jit.hpp"
#include "multi_gpu_utils.hpp"

namespace {
void *g_libcuda = nullptr;
template <class F> F real(const char *name) {
  if (!g_libcuda) g_libcuda = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!g_libcuda) return nullptr;
  return reinterpret_cast<F>(dlsym(g_libcuda, name));
}
int g_primary = 0;
float g_last_ms = -1.0f;
int g_launches = 0;
}  // namespace

extern "C" {

CUresult CUDAAPI cuCtxCreate_v2(CUcontext *pctx, unsigned int flags, CUdevice dev) {
  if (g_primary) {
    auto retain = real<CUresult(CUDAAPI *)(CUcontext *, CUdevice)>("cuDevicePrimaryCtxRetain");
    auto setcur = real<CUresult(CUDAAPI *)(CUcontext)>("cuCtxSetCurrent");
    if (!retain || !setcur) return CUDA_ERROR_NOT_INITIALIZED;
    CUresult r = retain(pctx, dev);
    if (r != CUDA_SUCCESS) return r;
    return setcur(*pctx);
  }
  auto f = real<CUresult(CUDAAPI *)(CUcontext *, unsigned int, CUdevice)>("cuCtxCreate_v2");
  return f ? f(pctx, flags, dev) : CUDA_ERROR_NOT_INITIALIZED;
}
CUresult CUDAAPI cuCtxDestroy_v2(CUcontext ctx) {
  if (g_primary) {
    // the reference destroys "its" context after every call (src/jit.cpp:140-145); in primary
    // mode that must become a release of the retain above
    auto getdev = real<CUresult(CUDAAPI *)(CUdevice *)>("cuCtxGetDevice");
    auto release = real<CUresult(CUDAAPI *)(CUdevice)>("cuDevicePrimaryCtxRelease_v2");
    CUdevice dev = 0;
    if (getdev) getdev(&dev);
    return release ? release(dev) : CUDA_ERROR_NOT_INITIALIZED;
  }
  auto f = real<CUresult(CUDAAPI *)(CUcontext)>("cuCtxDestroy_v2");
  return f ? f(ctx) : CUDA_ERROR_NOT_INITIALIZED;
}
CUresult CUDAAPI cuLaunchKernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                                unsigned smem, CUstream s, void **params, void **extra) {
  auto launch = real<CUresult(CUDAAPI *)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                         CUstream, void **, void **)>("cuLaunchKernel");
  auto evc = real<CUresult(CUDAAPI *)(CUevent *, unsigned)>("cuEventCreate");
  auto evr = real<CUresult(CUDAAPI *)(CUevent, CUstream)>("cuEventRecord");
  auto evs = real<CUresult(CUDAAPI *)(CUevent)>("cuEventSynchronize");
  auto eve = real<CUresult(CUDAAPI *)(float *, CUevent, CUevent)>("cuEventElapsedTime");
  auto evd = real<CUresult(CUDAAPI *)(CUevent)>("cuEventDestroy_v2");
  if (!launch) return CUDA_ERROR_NOT_INITIALIZED;
  CUevent a = nullptr, b = nullptr;
  bool timed = evc && evr && evs && eve && evd && evc(&a, 0) == CUDA_SUCCESS && evc(&b, 0) == CUDA_SUCCESS;
  if (timed) evr(a, s);
  CUresult r = launch(f, gx, gy, gz, bx, by, bz, smem, s, params, extra);
  g_launches++;
  if (timed) {
    evr(b, s);
    if (r == CUDA_SUCCESS && evs(b) == CUDA_SUCCESS) eve(&g_last_ms, a, b);
    else g_last_ms = -1.0f;
    evd(a);
    evd(b);
  }
  return r;
}

void ref_set_primary_ctx(int on) { g_primary = on; }
float ref_last_kernel_ms(void) { return g_last_ms; }
int ref_launches(void) { return g_launches; }

static int fill_table(Table &t, const char **names, const int *dtypes, void **dptrs, int ncols, int num_rows) {
  t.num_rows = num_rows;
  for (int i = 0; i < ncols; ++i) t.columns.push_back({names[i], static_cast<DataType>(dtypes[i]), dptrs[i], num_rows});
  return 0;
}
#define REF_TRY(stmt)                                                    \
  try { stmt; } catch (const std::exception &e) {                        \
    if (err && errlen > 0) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; } \
    return 1;                                                            \
  }                                                                      \
  return 0;

// jit_compile_and_launch (include/jit.hpp:7-10)
int ref_jit_compile_and_launch(const char *expr, const char *cond, const char **names, const int *dtypes, void **dptrs,
                               int ncols, int num_rows, float *d_out, int device, char *err, int errlen) {
  Table t;
  fill_table(t, names, dtypes, dptrs, ncols, num_rows);
  REF_TRY(jit_compile_and_launch(expr, cond ? cond : "", t, d_out, device))
}
// jit_group_sum (include/jit.hpp:15-18)
int ref_jit_group_sum(const char *val, const char *key, float *d_price, int *d_quantity, float *d_out_vals, int *d_out_keys,
                      int *d_count, int n, int device, char *err, int errlen) {
  REF_TRY(jit_group_sum(val, key, d_price, d_quantity, d_out_vals, d_out_keys, d_count, n, device))
}
// jit_sort_pairs / jit_sort_float (include/jit.hpp:22-27)
int ref_jit_sort_pairs(int *d_keys, float *d_vals, int count, int ascending, int device, char *err, int errlen) {
  REF_TRY(jit_sort_pairs(d_keys, d_vals, count, ascending != 0, device))
}
int ref_jit_sort_float(float *d_vals, int count, int ascending, int device, char *err, int errlen) {
  REF_TRY(jit_sort_float(d_vals, count, ascending != 0, device))
}
}  // extern "C"
