// aux_kernels.cu -- the small kernels around the scan: synthetic fill, finite
// check, row packing, mask ops, K4 prefix search + scatter, K5 gathered merge.
#include "aux_kernels.cuh"
#include "common.cuh"
#include "scan.cuh"

namespace tss {

// ---- synthetic corpus (bit-identical to oracle/oracle.cpp:gen_row) ------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float synth_from_u32(uint32_t h) {
  int v = (int)(h & 0xFFFFu) + (int)(h >> 16) - 65535;
  return (float)v * (1.0f / 65536.0f);
}
__device__ __forceinline__ uint16_t f32_to_bf16_rne(float x) {
  uint32_t u = __float_as_uint(x);
  if ((u & 0x7F800000u) != 0x7F800000u) u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

__global__ void synth_fill_kernel(void* dst, uint64_t row_begin, uint64_t nrows, uint32_t dim,
                                  uint32_t stride_elems, int bf16, uint64_t seed) {
  const uint32_t pairs = stride_elems >> 1;
  const uint64_t total = nrows * pairs;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = i / pairs;
    uint32_t pr = (uint32_t)(i - r * pairs);
    uint32_t j = pr * 2;
    float a = 0.f, b = 0.f;
    if (j < dim) {
      uint64_t h = mix64((((row_begin + r) << 16) | (uint64_t)pr) + seed * 0x9E3779B97F4A7C15ull);
      a = synth_from_u32((uint32_t)h);
      if (j + 1 < dim) b = synth_from_u32((uint32_t)(h >> 32));
    }
    if (bf16) {
      uint32_t packed = (uint32_t)f32_to_bf16_rne(a) | ((uint32_t)f32_to_bf16_rne(b) << 16);
      reinterpret_cast<uint32_t*>(dst)[r * pairs + pr] = packed;
    } else {
      reinterpret_cast<float2*>(dst)[r * pairs + pr] = make_float2(a, b);
    }
  }
}

cudaError_t launch_synth_fill(void* dst, uint64_t row_begin, uint64_t nrows, uint32_t dim,
                              uint32_t stride_elems, bool bf16, uint64_t seed, cudaStream_t st) {
  if (!nrows) return cudaSuccess;
  synth_fill_kernel<<<148 * 8, 256, 0, st>>>(dst, row_begin, nrows, dim, stride_elems, bf16 ? 1 : 0,
                                             seed);
  return cudaGetLastError();
}

__global__ void check_finite_kernel(const float* src, uint64_t count, int* flag) {
  int bad = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x)
    bad |= !isfinite(src[i]);
  if (__any_sync(FULL_MASK, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}
cudaError_t launch_check_finite(const float* src, uint64_t count, int* flag, cudaStream_t st) {
  if (!count) return cudaSuccess;
  check_finite_kernel<<<148 * 4, 256, 0, st>>>(src, count, flag);
  return cudaGetLastError();
}

__global__ void pack_rows_kernel(const float* src, void* dst, uint64_t nrows, uint32_t dim,
                                 uint32_t stride_elems, int bf16) {
  const uint64_t total = nrows * stride_elems;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = i / stride_elems;
    uint32_t j = (uint32_t)(i - r * stride_elems);
    float v = j < dim ? src[r * dim + j] : 0.f;
    if (bf16)
      reinterpret_cast<uint16_t*>(dst)[i] = f32_to_bf16_rne(v);
    else
      reinterpret_cast<float*>(dst)[i] = v;
  }
}
cudaError_t launch_pack_rows(const float* src, void* dst, uint64_t nrows, uint32_t dim,
                             uint32_t stride_elems, bool bf16, cudaStream_t st) {
  if (!nrows) return cudaSuccess;
  pack_rows_kernel<<<148 * 8, 256, 0, st>>>(src, dst, nrows, dim, stride_elems, bf16 ? 1 : 0);
  return cudaGetLastError();
}

// one warp per row; the row is read twice (the second read hits L1/L2).  *max_err (float bits,
// >= 0 so unsigned order == float order) ends up holding an upper bound of
// max over rows of | e_n - e / |e| |: how far bf16 rounding moved any unit row -- the quantity the
// re-scoring margins are built from (about 0.43 * 2^-8 on ordinary data, 2^-8 in the worst case).
__global__ void normalize_rows_kernel(const void* src, int src_bf16, uint16_t* dst, uint64_t nrows,
                                      uint32_t stride_elems, unsigned int* max_err) {
  const uint64_t r = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  const float* sf = reinterpret_cast<const float*>(src) + r * stride_elems;
  const uint16_t* sh = reinterpret_cast<const uint16_t*>(src) + r * stride_elems;
  auto at = [&](uint32_t j) -> float {
    return src_bf16 ? __uint_as_float((uint32_t)sh[j] << 16) : sf[j];
  };
  float ss = 0.f;
  for (uint32_t j = lane; j < stride_elems; j += 32) {
    const float v = at(j);
    ss = fmaf(v, v, ss);
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, m);
  const float inv = ss > 0.f && isfinite(ss) ? rsqrtf(ss) : 0.f;
  uint16_t* d = dst + r * stride_elems;
  float err2 = 0.f;
  for (uint32_t j = lane; j < stride_elems; j += 32) {
    const float x = at(j) * inv;
    const uint16_t h = f32_to_bf16_rne(x);
    d[j] = h;
    const float dx = __uint_as_float((uint32_t)h << 16) - x;
    err2 = fmaf(dx, dx, err2);
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) err2 += __shfl_xor_sync(FULL_MASK, err2, m);
  // (+2^-9 relative: the fp32 sums above and rsqrtf's 2 ulp are far inside that)
  if (lane == 0 && max_err) atomicMax(max_err, __float_as_uint(sqrtf(err2) * 1.002f));
}
cudaError_t launch_normalize_rows(const void* src, bool src_bf16, void* dst_bf16, uint64_t nrows,
                                  uint32_t stride_elems, unsigned int* d_max_err, cudaStream_t st) {
  if (d_max_err) {
    cudaError_t e = cudaMemsetAsync(d_max_err, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
  }
  if (!nrows) return cudaSuccess;
  const uint64_t blocks = (nrows * 32 + 255) / 256;
  normalize_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, src_bf16 ? 1 : 0,
                                                         reinterpret_cast<uint16_t*>(dst_bf16), nrows,
                                                         stride_elems, d_max_err);
  return cudaGetLastError();
}

__global__ void unpack_rows_kernel(const void* src, float* dst, uint64_t nrows, uint32_t dim,
                                   uint32_t stride_elems, int bf16) {
  const uint64_t total = nrows * dim;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = i / dim;
    uint32_t j = (uint32_t)(i - r * dim);
    uint64_t si = r * stride_elems + j;
    dst[i] = bf16 ? __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(src)[si] << 16)
                  : reinterpret_cast<const float*>(src)[si];
  }
}
cudaError_t launch_unpack_rows(const void* src, float* dst, uint64_t nrows, uint32_t dim,
                               uint32_t stride_elems, bool bf16, cudaStream_t st) {
  if (!nrows) return cudaSuccess;
  unpack_rows_kernel<<<148 * 8, 256, 0, st>>>(src, dst, nrows, dim, stride_elems, bf16 ? 1 : 0);
  return cudaGetLastError();
}

// ---- masks -----------------------------------------------------------------------------
__global__ void mask_set_rows_kernel(uint32_t* words, uint64_t nbits, const uint32_t* rows,
                                     uint64_t n, uint64_t row_base, int set) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = (uint64_t)rows[i] - row_base;  // wraps to huge if below the shard
    if (r >= nbits) continue;
    if (set) atomicOr(words + (r >> 5), 1u << (uint32_t)(r & 31));
    else atomicAnd(words + (r >> 5), ~(1u << (uint32_t)(r & 31)));
  }
}
cudaError_t launch_mask_set_rows(uint32_t* words, uint64_t nbits, const uint32_t* rows, uint64_t n,
                                 uint64_t row_base, bool set, cudaStream_t st) {
  if (!n) return cudaSuccess;
  int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  mask_set_rows_kernel<<<grid, 256, 0, st>>>(words, nbits, rows, n, row_base, set ? 1 : 0);
  return cudaGetLastError();
}

__global__ void mask_set_rows_inline_kernel(uint32_t* words, uint64_t nbits,
                                            const __grid_constant__ InlineRows rows, uint64_t row_base,
                                            int set) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows.n) return;
  uint64_t r = (uint64_t)rows.rows[i] - row_base;
  if (r >= nbits) return;
  if (set) atomicOr(words + (r >> 5), 1u << (uint32_t)(r & 31));
  else atomicAnd(words + (r >> 5), ~(1u << (uint32_t)(r & 31)));
}
cudaError_t launch_mask_set_rows_inline(uint32_t* words, uint64_t nbits, const InlineRows& rows,
                                        uint64_t row_base, bool set, cudaStream_t st) {
  if (!rows.n) return cudaSuccess;
  mask_set_rows_inline_kernel<<<(rows.n + 127) / 128, 128, 0, st>>>(words, nbits, rows, row_base,
                                                                   set ? 1 : 0);
  return cudaGetLastError();
}

// set (or clear) the bits of the rows named by packed keys (0 = empty slot)
__global__ void mask_update_from_keys_kernel(uint32_t* words, uint64_t nbits, const uint64_t* keys,
                                             uint32_t n, uint64_t row_base, int set) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint64_t key = keys[i];
    if (!key) continue;
    uint64_t r = (uint64_t)(0xFFFFFFFFu - (uint32_t)key) - row_base;
    if (r >= nbits) continue;
    if (set) atomicOr(words + (r >> 5), 1u << (uint32_t)(r & 31));
    else atomicAnd(words + (r >> 5), ~(1u << (uint32_t)(r & 31)));
  }
}
cudaError_t launch_mask_update_from_keys(uint32_t* words, uint64_t nbits, const uint64_t* keys,
                                         uint32_t n, uint64_t row_base, bool set, cudaStream_t st) {
  if (!n) return cudaSuccess;
  mask_update_from_keys_kernel<<<(n + 255) / 256, 256, 0, st>>>(words, nbits, keys, n, row_base,
                                                                set ? 1 : 0);
  return cudaGetLastError();
}

// N3: columnar metadata pre-filter.  One thread builds one 32-row mask word: a row passes when
// its date lies in [lo,hi] and (no court list, or its court id is in the 65 536-bit allow set).
__global__ void filter_mask_kernel(const uint16_t* court, const int32_t* date, uint64_t nrows,
                                   const uint32_t* allow_bits, int any_court, int32_t lo, int32_t hi,
                                   uint32_t* words, int combine_and) {
  __shared__ uint32_t s_allow[2048];
  if (!any_court)
    for (uint32_t i = threadIdx.x; i < 2048; i += blockDim.x) s_allow[i] = allow_bits[i];
  __syncthreads();
  const uint64_t nwords = (nrows + 31) / 32;
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < nwords;
       w += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t bits = 0;
    const uint64_t r0 = w * 32;
#pragma unroll 4
    for (uint32_t j = 0; j < 32; ++j) {
      const uint64_t r = r0 + j;
      if (r >= nrows) break;
      const int32_t d = date[r];
      bool ok = d >= lo && d <= hi;
      if (ok && !any_court) {
        const uint32_t c = court[r];
        ok = (s_allow[c >> 5] >> (c & 31)) & 1u;
      }
      bits |= (uint32_t)ok << j;
    }
    words[w] = combine_and ? (words[w] & bits) : bits;
  }
}
cudaError_t launch_filter_mask(const uint16_t* court, const int32_t* date, uint64_t nrows,
                               const uint32_t* allow_bits, bool any_court, int32_t lo, int32_t hi,
                               uint32_t* words, bool combine_and, cudaStream_t st) {
  if (!nrows) return cudaSuccess;
  const uint64_t nwords = (nrows + 31) / 32;
  int grid = (int)((nwords + 255) / 256 < 148 * 8 ? (nwords + 255) / 256 : 148 * 8);
  filter_mask_kernel<<<grid, 256, 0, st>>>(court, date, nrows, allow_bits, any_court ? 1 : 0, lo, hi,
                                           words, combine_and ? 1 : 0);
  return cudaGetLastError();
}

__global__ void mask_popcount_kernel(const uint32_t* words, uint64_t nwords,
                                     unsigned long long* out) {
  unsigned long long c = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nwords;
       i += (uint64_t)gridDim.x * blockDim.x)
    c += __popc(words[i]);
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) c += __shfl_xor_sync(FULL_MASK, c, m);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
cudaError_t launch_mask_popcount(const uint32_t* words, uint64_t nwords, unsigned long long* out,
                                 cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess || !nwords) return e;
  mask_popcount_kernel<<<148 * 2, 256, 0, st>>>(words, nwords, out);
  return cudaGetLastError();
}

// ---- K4: prefix search ------------------------------------------------------------------
// term i (bytes pool[term_off[i], term_off[i+1])) < key ?  (unsigned byte order;
// a proper prefix sorts first)
__device__ __forceinline__ bool term_less(const TermsDev& t, uint64_t i, const char* key,
                                          uint32_t klen) {
  uint64_t b = t.term_off[i], e = t.term_off[i + 1];
  uint32_t tlen = (uint32_t)(e - b);
  uint32_t n = tlen < klen ? tlen : klen;
  const unsigned char* tp = reinterpret_cast<const unsigned char*>(t.pool) + b;
  const unsigned char* kp = reinterpret_cast<const unsigned char*>(key);
  for (uint32_t j = 0; j < n; ++j) {
    unsigned char a = tp[j], c = kp[j];
    if (a != c) return a < c;
  }
  return tlen < klen;
}

// ---- K4 fused: search + clear, grid barrier, scatter (+ unique-row list) -----------------------
// 256-ary cooperative lower bound: every round each thread probes one pivot and the CTA counts
// the pivots below the key (the predicate is monotone over sorted terms): ceil(log256 T) rounds
// of two dependent loads instead of log2 T.
__device__ uint64_t cta_lower_bound(const TermsDev& t, const char* key, uint32_t klen) {
  uint64_t lo = 0, hi = t.nterms;
  const uint32_t nt = blockDim.x;
  while (hi - lo > nt) {
    const uint64_t step = (hi - lo + nt - 1) / nt;
    const uint64_t pv = lo + (uint64_t)threadIdx.x * step;
    const bool less = pv < hi && term_less(t, pv, key, klen);
    const int c = __syncthreads_count(less);
    if (c == 0) return lo;
    const uint64_t nlo = lo + (uint64_t)(c - 1) * step + 1;
    uint64_t nhi = lo + (uint64_t)c * step;
    if (nhi > hi) nhi = hi;
    lo = nlo;
    hi = nhi;
  }
  const uint64_t pv = lo + threadIdx.x;
  const bool less = pv < hi && term_less(t, pv, key, klen);
  return lo + (uint64_t)__syncthreads_count(less);
}

constexpr int kPrefixThreads = 256;
__global__ void __launch_bounds__(kPrefixThreads)
prefix_mask_kernel(const PrefixMaskArgs a, const __grid_constant__ PrefixKeys keys) {
  const uint32_t G = gridDim.x;
  // ---- phase A: bounds (CTAs 0..3) | clear (the rest) ----
  if (blockIdx.x < 4) {
    const int b = blockIdx.x;
    const char* kb = a.d_keybytes ? a.d_keybytes : keys.bytes;
    uint64_t r;
    if (keys.fixed[b] >= 0) r = (uint64_t)keys.fixed[b];
    else if (keys.fixed[b] == -2) r = a.t.nterms;
    else r = cta_lower_bound(a.t, kb + keys.off[b], keys.off[b + 1] - keys.off[b]);
    if (threadIdx.x == 0) {
      a.bounds[b] = r;
      if (b == 0 && a.list_count) *a.list_count = 0u;
    }
  }
  if (a.clear_nwords) {
    // every CTA clears a share (search CTAs after their search: they are a small minority and
    // the barrier below waits for them anyway)
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nth = (uint64_t)G * blockDim.x;
    uint4* w4 = reinterpret_cast<uint4*>(a.words);  // cudaMalloc'ed: 256-byte aligned
    const uint64_t n4 = a.clear_nwords >> 2;
    for (uint64_t i = tid; i < n4; i += nth) w4[i] = make_uint4(0u, 0u, 0u, 0u);
    for (uint64_t i = (n4 << 2) + tid; i < a.clear_nwords; i += nth) a.words[i] = 0u;
  }
  // ---- grid barrier (all G <= #SMs CTAs are co-resident; the words are left zeroed) ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&a.sync[0], 1u);
    while (*reinterpret_cast<volatile unsigned int*>(&a.sync[0]) < G) __nanosleep(20);
    __threadfence();
    if (atomicAdd(&a.sync[1], 1u) == G - 1) {  // last one through: reset for the next launch
      a.sync[1] = 0u;
      __threadfence();
      a.sync[0] = 0u;
    }
  }
  __syncthreads();
  // ---- phase B: scatter the two posting ranges ----
  uint64_t pb[2], pe[2], total = 0;
#pragma unroll
  for (int rng = 0; rng < 2; ++rng) {
    const uint64_t tlo = __ldcg(a.bounds + 2 * rng), thi = __ldcg(a.bounds + 2 * rng + 1);
    pb[rng] = pe[rng] = 0;
    if (thi > tlo) {
      pb[rng] = a.t.post_off[tlo];
      pe[rng] = a.t.post_off[thi];
    }
    total += pe[rng] - pb[rng];
  }
  const bool use_list = a.list && total <= a.list_cap;
  const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nth = (uint64_t)G * blockDim.x;
  if (gtid == 0) {
    a.bounds[4] = total;
    if (a.list_count && !use_list) *a.list_count = 0xFFFFFFFFu;
  }
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ uint32_t s_wcnt[kPrefixThreads / 32];
  __shared__ uint32_t s_base;
#pragma unroll
  for (int rng = 0; rng < 2; ++rng) {
    if (!use_list) {
      for (uint64_t i = pb[rng] + gtid; i < pe[rng]; i += nth) {
        const uint64_t r = (uint64_t)a.t.post_rows[i] - a.row_base;
        if (r < a.nbits) atomicOr(a.words + (r >> 5), 1u << (uint32_t)(r & 31));
      }
      continue;
    }
    // with a row list: whole CTAs iterate together, and the threads whose atomicOr flipped the
    // bit (the row is new) reserve their list slots with ONE atomicAdd per CTA per trip
    for (uint64_t c0 = pb[rng] + (uint64_t)blockIdx.x * blockDim.x; c0 < pe[rng]; c0 += nth) {
      const uint64_t i = c0 + threadIdx.x;
      const bool in = i < pe[rng];
      const uint64_t r = in ? (uint64_t)a.t.post_rows[i] - a.row_base : ~0ull;
      const bool ok = in && r < a.nbits;
      const uint32_t bit = 1u << (uint32_t)(r & 31);
      uint32_t old = 0xFFFFFFFFu;
      if (ok) old = atomicOr(a.words + (r >> 5), bit);
      const bool win = ok && !(old & bit);
      const unsigned m = __ballot_sync(FULL_MASK, win);
      if (lane == 0) s_wcnt[warp] = __popc(m);
      __syncthreads();
      uint32_t before = 0, total_w = 0;
#pragma unroll
      for (uint32_t w = 0; w < kPrefixThreads / 32; ++w) {
        if (w < warp) before += s_wcnt[w];
        total_w += s_wcnt[w];
      }
      if (threadIdx.x == 0 && total_w) s_base = atomicAdd(a.list_count, total_w);
      __syncthreads();
      if (win) {
        const uint32_t pos = s_base + before + __popc(m & ((1u << lane) - 1u));
        if (pos < a.list_cap) a.list[pos] = (uint32_t)r;
      }
      // (no third barrier: s_wcnt is rewritten in the next trip only by warps that have passed
      // the barrier above, after every read of it; s_base only after the next trip's first
      // barrier, which every thread reaches after reading it here)
    }
  }
}
cudaError_t launch_prefix_mask(const PrefixMaskArgs& a, const PrefixKeys& keys, int num_sms,
                               cudaStream_t st) {
  // one CTA per SM at most (the grid barrier needs every CTA resident); a small mask does not
  // need them all
  int grid = num_sms > 0 ? num_sms : 148;
  const uint64_t want = 4 + (a.clear_nwords / 4 + kPrefixThreads - 1) / kPrefixThreads;
  if (a.clear_nwords && want < (uint64_t)grid && want >= 8) grid = (int)want;
  if (grid < 8) grid = 8;
  if (grid > num_sms && num_sms > 0) grid = num_sms;
  // a cooperative launch: the grid barrier needs every CTA resident at once, and this is the
  // launch form that guarantees it (or fails instead of hanging)
  void* args[] = {const_cast<PrefixMaskArgs*>(&a), const_cast<PrefixKeys*>(&keys)};
  return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(prefix_mask_kernel), dim3((unsigned)grid),
                                     dim3(kPrefixThreads), args, 0, st);
}

__global__ void terms_validate_kernel(TermsDev t, uint64_t pool_bytes, uint64_t nposts,
                                      unsigned int* bad) {
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (tid == 0) {
    if (t.term_off[0] != 0 || t.post_off[0] != 0 || t.term_off[t.nterms] != pool_bytes ||
        t.post_off[t.nterms] != nposts)
      atomicOr(bad, 1u);
  }
  for (uint64_t i = tid; i < t.nterms; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t b = t.term_off[i], e = t.term_off[i + 1];
    bool ok = e >= b && e <= pool_bytes && t.post_off[i + 1] >= t.post_off[i] &&
              t.post_off[i + 1] <= nposts;
    if (ok && i) {  // term[i-1] < term[i]
      const uint64_t pb = t.term_off[i - 1];
      ok = pb <= b;
      if (ok) {
        const uint64_t la = b - pb, lb = e - b, n = la < lb ? la : lb;
        const unsigned char* pa = reinterpret_cast<const unsigned char*>(t.pool) + pb;
        const unsigned char* pc = reinterpret_cast<const unsigned char*>(t.pool) + b;
        int c = 0;
        for (uint64_t j = 0; j < n && !c; ++j) c = (int)pa[j] - (int)pc[j];
        ok = c < 0 || (c == 0 && la < lb);
      }
    }
    if (!ok) atomicOr(bad, 1u);
  }
}
cudaError_t launch_terms_validate(const TermsDev& t, uint64_t pool_bytes, uint64_t nposts,
                                  unsigned int* d_bad, cudaStream_t st) {
  terms_validate_kernel<<<148 * 4, 256, 0, st>>>(t, pool_bytes, nposts, d_bad);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(1024)
redo_compact_kernel(const uint32_t* flags, uint32_t nq, uint32_t* d_list, uint32_t* d_count,
                    uint32_t* h_list, uint32_t* h_count, uint32_t fixups, unsigned int* h_sticky) {
  __shared__ uint32_t s_warp[32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool f = tid < nq && flags[tid] != 0;  // nq <= 1024: one thread per query
  const unsigned m = __ballot_sync(FULL_MASK, f);
  if (lane == 0) s_warp[warp] = __popc(m);
  __syncthreads();
  uint32_t before = 0, total = 0;
  for (uint32_t w = 0; w < 32; ++w) {
    if (w < warp) before += s_warp[w];
    total += s_warp[w];
  }
  if (f) {
    const uint32_t pos = before + __popc(m & ((1u << lane) - 1u));
    d_list[pos] = tid;
    h_list[pos] = tid;
  }
  if (tid == 0) {
    *d_count = total;
    *h_count = total;
    if (total > fixups) atomicOr(h_sticky, 2u);
  }
}
cudaError_t launch_redo_compact(const uint32_t* flags, uint32_t nq, uint32_t* d_list,
                                uint32_t* d_count, uint32_t* h_list, uint32_t* h_count,
                                uint32_t fixups, unsigned int* h_sticky, cudaStream_t st) {
  if (nq > 1024) return cudaErrorInvalidValue;
  redo_compact_kernel<<<1, 1024, 0, st>>>(flags, nq, d_list, d_count, h_list, h_count, fixups,
                                          h_sticky);
  return cudaGetLastError();
}

// ---- K5: merge of all-gathered per-rank top-k lists ---------------------------------------
// one warp per query: stage the P sorted lists of k keys in shared memory, then a warp
// tournament (scan.cuh) picks the k best.  in [P][nq][k] -> out [nq][k].
__global__ void merge_gathered_kernel(const uint64_t* in, uint64_t* out, uint32_t P, uint32_t nq,
                                      uint32_t k) {
  extern __shared__ uint64_t sk[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t wpc = blockDim.x >> 5;
  const uint32_t qi = blockIdx.x * wpc + warp;
  if (qi >= nq) return;
  uint64_t* mine = sk + (size_t)warp * P * k;
  for (uint32_t i = lane; i < P * k; i += 32) {
    uint32_t r = i / k, e = i - r * k;
    mine[i] = in[((size_t)r * nq + qi) * k + e];
  }
  __syncwarp();
  warp_tournament<2>(mine, k, P, k, k, out + (size_t)qi * k, lane);
}
cudaError_t launch_merge_gathered(const uint64_t* in, uint64_t* out, uint32_t P, uint32_t nq,
                                  uint32_t k, cudaStream_t st) {
  if (!nq) return cudaSuccess;
  if (P > 64) return cudaErrorInvalidConfiguration;
  size_t per_warp = (size_t)P * k * 8;
  if (per_warp > 48 * 1024) return cudaErrorInvalidConfiguration;
  uint32_t wpc = (uint32_t)(48 * 1024 / per_warp);
  if (wpc > 4) wpc = 4;
  if (wpc > nq) wpc = nq;
  merge_gathered_kernel<<<(nq + wpc - 1) / wpc, wpc * 32, per_warp * wpc, st>>>(in, out, P, nq, k);
  return cudaGetLastError();
}

}  // namespace tss
