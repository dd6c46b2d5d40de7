// prelude.cuh -- common device helpers prepended to every NVRTC-instantiated kernel template.
// Self-contained (no CUDA headers: NVRTC builtins only).  Target: sm_100a.
//
// Macros supplied by the host code generator (warpcore.cu: gen_header):
//   WDB_VEC        rows per vector access for 4-byte columns: 4 (128-bit) or 8 (256-bit LDG/STG,
//                  new on sm_100: LDG.E.256 / STG.E.256)
//   WDB_ALIGNED    1 when every used column pointer and the output are WDB_VEC*4-byte aligned
//   WDB_LD_HINT    0 plain .nc.L1::no_allocate | 1 + L2::256B prefetch | 2 + L2::evict_first (256-bit only)
//   WDB_ST_HINT    0 default | 1 L1::no_allocate | 2 .cs | 3 L2::evict_first (256-bit only)
typedef long long i64;
typedef unsigned long long u64;
typedef unsigned int u32;

// `price[idx]` in a reference-style expression (include/expression.hpp:45) resolves to the
// register copy of the row's value through this proxy; its C++ type is the column's type, so the
// usual arithmetic conversions are those of the reference kernel (src/jit.cpp:75-83).
struct wdb_idx_t {};
template <class T> struct wdb_cell {
  T v;
  __device__ __forceinline__ T &operator[](wdb_idx_t) { return v; }
  // a bare column name denotes the current row's value too: the reference's own JIT tests pass
  // "price" and "price + 1" as expression code (tests/jit_arch_test.cpp:23, jit_error_test.cpp:28),
  // which its kernel text would reject (a float* assigned to a float)
  __device__ __forceinline__ operator T() const { return v; }
};

#define WDB_FULL_MASK 0xffffffffu

// ORDER BY keys: one NaN policy for every path (register top-k, threshold compaction, radix sort):
// a NaN key orders as the worst key of the direction, after every number (fmaxf / fminf return the
// non-NaN operand); among themselves and against a real -inf / +inf such rows keep row order.
__device__ __forceinline__ float wdb_nanlast_d(float k) { return fmaxf(k, __int_as_float(0xff800000)); }
__device__ __forceinline__ float wdb_nanlast_a(float k) { return fminf(k, __int_as_float(0x7f800000)); }

__device__ __forceinline__ u32 wdb_lane() { u32 l; asm("mov.u32 %0, %%laneid;" : "=r"(l)); return l; }
__device__ __forceinline__ u32 wdb_lanemask_lt() { u32 m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// ---- streaming global loads ------------------------------------------------------------------
__device__ __forceinline__ void wdb_ldg16(const void *p, u32 (&r)[4]) {
#if WDB_LD_HINT == 1
  asm("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];"
#else
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
#endif
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(p));
}
// H: 0 plain | 1 L2::256B prefetch | 2 L2::evict_first (stream: do not displace what should stay)
//    | 3 L2::evict_last (park: this line will be read again soon).  The L2 eviction qualifiers exist
//    only for the 256-bit forms on sm_100.
template <int H> __device__ __forceinline__ void wdb_ldg32h(const void *p, u32 (&r)[8]) {
  if constexpr (H == 2)
    asm("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
  else if constexpr (H == 3)
    asm("ld.global.nc.L1::no_allocate.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
  else if constexpr (H == 1)
    asm("ld.global.nc.L1::no_allocate.L2::256B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
  else
    asm("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ void wdb_ldg32(const void *p, u32 (&r)[8]) { wdb_ldg32h<WDB_LD_HINT>(p, r); }
// scalar store with an L2 evict-first policy (output streams that must not displace parked input)
__device__ __forceinline__ void wdb_st_f32_stream(float *p, float v) {
  u64 pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(p), "f"(v), "l"(pol) : "memory");
}
// non-blocking prefetch of the line holding `p` into the L2 (no destination register)
__device__ __forceinline__ void wdb_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(p)); }
__device__ __forceinline__ void wdb_stg16(void *p, const u32 (&r)[4]) {
#if WDB_ST_HINT == 1
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
#elif WDB_ST_HINT == 2
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};"
#else
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};"
#endif
               :: "l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void wdb_stg32(void *p, const u32 (&r)[8]) {
#if WDB_ST_HINT == 3
  asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
#elif WDB_ST_HINT == 1
  asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
#elif WDB_ST_HINT == 2
  asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
#else
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
#endif
               :: "l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

template <class T> __device__ __forceinline__ T wdb_from_bits32(u32 b) { return *reinterpret_cast<T *>(&b); }
template <class T> __device__ __forceinline__ T wdb_from_bits64(u32 lo, u32 hi) {
  u64 b = ((u64)hi << 32) | lo;
  return *reinterpret_cast<T *>(&b);
}

// Load WDB_VEC consecutive rows of one column starting at row index `row` (a multiple of WDB_VEC
// on the aligned path).
template <int H, class T> __device__ __forceinline__ void wdb_load_vec_h(const T *__restrict__ p, i64 row, T (&o)[WDB_VEC]) {
#if WDB_ALIGNED
  if constexpr (sizeof(T) == 4) {
#if WDB_VEC == 8
    u32 r[8];
    wdb_ldg32h<H>(p + row, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = wdb_from_bits32<T>(r[j]);
#else
    u32 r[4];
    wdb_ldg16(p + row, r);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = wdb_from_bits32<T>(r[j]);
#endif
  } else {
#if WDB_VEC == 8
    u32 a[8], b[8];
    wdb_ldg32h<H>(p + row, a);
    wdb_ldg32h<H>(p + row + 4, b);
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[j] = wdb_from_bits64<T>(a[2 * j], a[2 * j + 1]); o[4 + j] = wdb_from_bits64<T>(b[2 * j], b[2 * j + 1]); }
#else
    u32 a[4], b[4];
    wdb_ldg16(p + row, a);
    wdb_ldg16(p + row + 2, b);
    o[0] = wdb_from_bits64<T>(a[0], a[1]); o[1] = wdb_from_bits64<T>(a[2], a[3]);
    o[2] = wdb_from_bits64<T>(b[0], b[1]); o[3] = wdb_from_bits64<T>(b[2], b[3]);
#endif
  }
#else
#pragma unroll
  for (int j = 0; j < WDB_VEC; ++j) o[j] = __ldg(p + row + j);
#endif
}
template <class T> __device__ __forceinline__ void wdb_load_vec(const T *__restrict__ p, i64 row, T (&o)[WDB_VEC]) { wdb_load_vec_h<WDB_LD_HINT>(p, row, o); }
// coherent (read-write data) vector load of WDB_VEC floats: used to blend into the output buffer
__device__ __forceinline__ void wdb_load_out_vec(const float *p, i64 row, float (&o)[WDB_VEC]) {
#if WDB_ALIGNED && WDB_VEC == 8
  u32 r[8];
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p + row) : "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(r[j]);
#elif WDB_ALIGNED
  u32 r[4];
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(p + row) : "memory");
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = __uint_as_float(r[j]);
#else
#pragma unroll
  for (int j = 0; j < WDB_VEC; ++j) o[j] = p[row + j];
#endif
}
__device__ __forceinline__ void wdb_store_vec(float *__restrict__ p, i64 row, const float (&v)[WDB_VEC]) {
#if WDB_ALIGNED
#if WDB_VEC == 8
  u32 r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(v[j]);
  wdb_stg32(p + row, r);
#else
  u32 r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) r[j] = __float_as_uint(v[j]);
  wdb_stg16(p + row, r);
#endif
#else
#pragma unroll
  for (int j = 0; j < WDB_VEC; ++j) p[row + j] = v[j];
#endif
}

// ---- mbarrier / bulk async copy (TMA 1-D) helpers ----------------------------------------------
__device__ __forceinline__ u32 wdb_smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wdb_mbar_init(u64 *bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(wdb_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void wdb_mbar_expect_tx(u64 *bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(wdb_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wdb_mbar_wait(u64 *bar, u32 parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WDB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WDB_DONE;\n"
      "bra WDB_WAIT;\n"
      "WDB_DONE:\n"
      "}\n" :: "r"(wdb_smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void wdb_fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void wdb_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// global -> shared bulk copy completing on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void wdb_bulk_g2s(void *smem_dst, const void *gsrc, u32 bytes, u64 *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(wdb_smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(wdb_smem_addr(bar)) : "memory");
}
// shared -> global bulk copy tracked by bulk groups
__device__ __forceinline__ void wdb_bulk_s2g(void *gdst, const void *smem_src, u32 bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"(wdb_smem_addr(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wdb_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wdb_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void wdb_bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }
