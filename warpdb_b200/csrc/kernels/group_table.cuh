// group_table.cuh -- device-side open-addressing aggregation table shared by the NVRTC consume
// kernel (group.cuh) and the statically compiled merge/export kernels (ops_group.cu).
// Self-contained: compiles under both nvcc and NVRTC.
//
// Layout (struct of arrays, capacity cap = mask+1 slots plus one "special" slot at index cap that
// holds the key equal to the empty-slot sentinel INT_MIN):
//   keys[cap+1] int32 (INT_MIN = empty)   sums[cap+1] f64   counts[cap+1] u64
//   mins/maxs[cap+1] order-preserving i64 encodings of f64   first[cap+1] i64 (smallest row id)
//   meta[0] = groups inserted, meta[1] = overflow flag, meta[2] = special slot used,
//   meta[3] = rows that bypassed the shared-memory accumulators, meta[4] = hash entries found outside
//   the side table's range when the table was folded for a cross-GPU merge (stale statistics)
#ifndef WDB_GROUP_TABLE_CUH
#define WDB_GROUP_TABLE_CUH

#define WDB_KEY_EMPTY ((int)0x80000000)
#define WDB_NEED_SUM_BIT 1
#define WDB_NEED_CNT_BIT 2
#define WDB_NEED_MINMAX_BIT 4
#define WDB_NEED_FIRST_BIT 8

struct wdb_table {
  int *keys;
  double *sums;
  unsigned long long *counts;
  long long *mins;
  long long *maxs;
  long long *first;
  unsigned int *meta;
  unsigned int mask;
  unsigned int shift;   // 32 - log2(capacity): the slot of a key is the TOP bits of its hash, so keys
                        // whose hashes share a prefix live in one contiguous region of the table
  // direct-addressed side table for integer keys with a known range (optimizer statistics): key k
  // lives at index k - dlo, no probe and no CAS.  dsums start as -0.0 (WDB_DENSE_EMPTY): adding any
  // value other than -0.0 changes the bit pattern, so an untouched slot is recognisable without a
  // second atomic; -0.0 addends are added as +0.0.  dmins / dmaxs hold order-preserving encodings
  // (+inf / -inf when untouched).  Only the arrays the table's needs call for exist.  dspan == 0: not
  // in use.  Identical layouts on every GPU make the cross-GPU merge one all-reduce per array
  // (sum / sum / min / max; -0.0 + -0.0 = -0.0 keeps the untouched marker).
  double *dsums;
  unsigned long long *dcnts;
  long long *dmins;
  long long *dmaxs;
  int dlo;
  unsigned int dspan;
};
#define WDB_DENSE_EMPTY 0x8000000000000000ull

// monotone map double -> signed 64-bit (a < b  <=>  enc(a) < enc(b) for non-NaN values)
__device__ __forceinline__ long long wdb_f64_enc(double d) {
  long long b = __double_as_longlong(d);
  return b >= 0 ? b : (long long)(0x8000000000000000ull - (unsigned long long)b);
}
__device__ __forceinline__ double wdb_f64_dec(long long e) {
  long long b = e >= 0 ? e : (long long)(0x8000000000000000ull - (unsigned long long)e);
  return __longlong_as_double(b);
}
#define WDB_ENC_PLUS_INF 0x7ff0000000000000ll
#define WDB_ENC_MINUS_INF (-0x7ff0000000000000ll)

// one row (or one partial aggregate) into entry di of the direct-addressed side table
template <int NEEDS>
__device__ __forceinline__ void wdb_dense_add(const wdb_table &T, unsigned int di, double sum, unsigned long long cnt,
                                              long long mn_enc, long long mx_enc) {
  if (NEEDS & WDB_NEED_SUM_BIT) atomicAdd(&T.dsums[di], sum + 0.0);
  if (NEEDS & WDB_NEED_CNT_BIT) atomicAdd(&T.dcnts[di], cnt);
  if (NEEDS & WDB_NEED_MINMAX_BIT) { atomicMin(&T.dmins[di], mn_enc); atomicMax(&T.dmaxs[di], mx_enc); }
}
// has entry i been touched?  (every row updates every accumulator the table tracks)
template <int NEEDS> __device__ __forceinline__ bool wdb_dense_present(const wdb_table &T, long long i) {
  if (NEEDS & WDB_NEED_CNT_BIT) return T.dcnts[i] != 0ull;
  if (NEEDS & WDB_NEED_SUM_BIT) return (unsigned long long)__double_as_longlong(T.dsums[i]) != WDB_DENSE_EMPTY;
  return T.dmins[i] != WDB_ENC_PLUS_INF || T.dmaxs[i] != WDB_ENC_MINUS_INF;
}

__device__ __forceinline__ unsigned int wdb_hash32(int key) {
  unsigned int x = (unsigned int)key;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// find or claim the slot of `key` (hash `hsh`); returns -1 when the table is full (overflow flag raised)
__device__ __forceinline__ long long wdb_table_slot_h(const wdb_table &T, int key, unsigned int hsh) {
  if (key == WDB_KEY_EMPTY) {
    if (T.meta[2] == 0u) atomicExch(&T.meta[2], 1u);
    return (long long)T.mask + 1;
  }
  unsigned int h = hsh >> T.shift;
  const unsigned int limit = T.mask < 65535u ? T.mask + 1u : 65536u;
  for (unsigned int p = 0; p < limit; ++p) {
    int k = T.keys[h];
    if (k == key) return h;
    if (k == WDB_KEY_EMPTY) {
      const int prev = atomicCAS(&T.keys[h], WDB_KEY_EMPTY, key);
      if (prev == WDB_KEY_EMPTY) { atomicAdd(&T.meta[0], 1u); return h; }
      if (prev == key) return h;
    }
    h = (h + 1u) & T.mask;
  }
  atomicExch(&T.meta[1], 1u);
  return -1;
}

__device__ __forceinline__ long long wdb_table_slot(const wdb_table &T, int key) { return wdb_table_slot_h(T, key, wdb_hash32(key)); }

// fold one partial aggregate into slot s
template <int NEEDS>
__device__ __forceinline__ void wdb_table_add(const wdb_table &T, long long s, double sum, unsigned long long cnt,
                                              long long mn_enc, long long mx_enc, long long first_row) {
  if (NEEDS & WDB_NEED_SUM_BIT) atomicAdd(&T.sums[s], sum);
  if (NEEDS & WDB_NEED_CNT_BIT) atomicAdd(&T.counts[s], cnt);
  if (NEEDS & WDB_NEED_MINMAX_BIT) { atomicMin(&T.mins[s], mn_enc); atomicMax(&T.maxs[s], mx_enc); }
  if (NEEDS & WDB_NEED_FIRST_BIT) atomicMin(&T.first[s], first_row);
}
#endif
