// terms_build.cu -- N2: build the flattened trie on the device.
//
// The reference builds its tries one insert at a time (TrieNode::insert, src/trie.rs:211-221:
// one HashMap hop per token) and nothing ever calls it.  Here the same structure -- unique
// terms in byte order with their postings in insertion order -- is built from N tokenised
// postings at once:
//   1. postings are tuples of <= L token ids (ids follow the byte order of the vocabulary, and
//      no token holds a byte <= ' ', so tuple order == byte order of the ' '-joined strings);
//   2. L stable LSD radix passes (cub::DeviceRadixSort, plumbing) sort the posting indices
//      lexicographically by tuple;
//   3. adjacent-difference flags + a prefix sum give term ids; heads scatter the CSR offsets;
//   4. term byte lengths -> exclusive scan -> one thread per term writes the joined string.
// Everything stays in HBM and becomes a tss_terms without a host round trip.
#include "terms_build.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <utility>
#include <vector>

namespace tss {

namespace {

__global__ void iota_kernel(uint32_t* idx, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    idx[i] = (uint32_t)i;
}
__global__ void gather_word_kernel(const uint32_t* ids, const uint32_t* idx, uint32_t L, uint32_t j,
                                   uint32_t* keys, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    keys[i] = ids[(uint64_t)idx[i] * L + j];
}
// head[i] = 1 when sorted posting i starts a new term
__global__ void heads_kernel(const uint32_t* ids, const uint32_t* idx, uint32_t L, uint64_t n,
                             uint32_t* head) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t h = 1;
    if (i) {
      const uint32_t* a = ids + (uint64_t)idx[i] * L;
      const uint32_t* b = ids + (uint64_t)idx[i - 1] * L;
      h = 0;
      for (uint32_t j = 0; j < L; ++j)
        if (a[j] != b[j]) {
          h = 1;
          break;
        }
    }
    head[i] = h;
  }
}
// per sorted posting: its row; per head: CSR offset, first posting index and byte length of its term
__global__ void scatter_kernel(const uint32_t* ids, const uint32_t* idx, const uint32_t* rows,
                               const uint32_t* head, const uint32_t* term_of, uint32_t L, uint64_t n,
                               const uint64_t* vocab_off, uint32_t* post_rows, uint64_t* post_off,
                               uint32_t* first_posting, uint64_t* term_len) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t src = idx[i];
    post_rows[i] = rows[src];
    if (head[i]) {
      const uint32_t t = term_of[i] - 1;  // inclusive scan -> 1-based
      post_off[t] = i;
      first_posting[t] = src;
      uint64_t len = 0;
      uint32_t ntok = 0;
      for (uint32_t j = 0; j < L; ++j) {
        uint32_t id = ids[(uint64_t)src * L + j];
        if (!id) break;
        len += vocab_off[id] - vocab_off[id - 1];
        ++ntok;
      }
      term_len[t] = len + (ntok ? ntok - 1 : 0);
    }
  }
}
__global__ void write_pool_kernel(const uint32_t* ids, const uint32_t* first_posting, uint32_t L,
                                  uint64_t nterms, const char* vocab_pool, const uint64_t* vocab_off,
                                  const uint64_t* term_off, char* pool) {
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < nterms;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t* tup = ids + (uint64_t)first_posting[t] * L;
    char* dst = pool + term_off[t];
    for (uint32_t j = 0; j < L; ++j) {
      uint32_t id = tup[j];
      if (!id) break;
      if (j) *dst++ = ' ';
      for (uint64_t b = vocab_off[id - 1]; b < vocab_off[id]; ++b) *dst++ = vocab_pool[b];
    }
  }
}

struct Scratch {
  std::vector<void*> ptrs;
  ~Scratch() {
    for (void* p : ptrs) cudaFree(p);
  }
  template <class T>
  cudaError_t alloc(T** out, size_t count) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (count ? count : 1) * sizeof(T));
    if (e == cudaSuccess) ptrs.push_back(p);
    *out = static_cast<T*>(p);
    return e;
  }
};

}  // namespace

#define TRY(expr)                      \
  do {                                 \
    cudaError_t e__ = (expr);          \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

cudaError_t build_terms_device(const char* vocab_pool, const uint64_t* vocab_off, uint32_t vocab_size,
                               const uint32_t* token_ids, uint32_t L, const uint32_t* rows,
                               uint64_t n, cudaStream_t st, BuiltTerms* out) {
  Scratch s;
  const int grid = 148 * 4, block = 256;
  char* d_vpool;
  uint64_t* d_voff;
  uint32_t *d_ids, *d_rows, *d_idx, *d_idx2, *d_keys, *d_keys2, *d_head, *d_termof, *d_first;
  const uint64_t vbytes = vocab_off[vocab_size];
  TRY(s.alloc(&d_vpool, vbytes));
  TRY(s.alloc(&d_voff, (size_t)vocab_size + 1));
  TRY(s.alloc(&d_ids, n * L));
  TRY(s.alloc(&d_rows, n));
  TRY(s.alloc(&d_idx, n));
  TRY(s.alloc(&d_idx2, n));
  TRY(s.alloc(&d_keys, n));
  TRY(s.alloc(&d_keys2, n));
  TRY(s.alloc(&d_head, n));
  TRY(s.alloc(&d_termof, n));
  if (vbytes) TRY(cudaMemcpyAsync(d_vpool, vocab_pool, vbytes, cudaMemcpyHostToDevice, st));
  TRY(cudaMemcpyAsync(d_voff, vocab_off, ((size_t)vocab_size + 1) * 8, cudaMemcpyHostToDevice, st));
  if (n) {
    TRY(cudaMemcpyAsync(d_ids, token_ids, n * L * 4, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_rows, rows, n * 4, cudaMemcpyHostToDevice, st));
  }
  uint64_t nterms = 0;
  if (n) {
    iota_kernel<<<grid, block, 0, st>>>(d_idx, n);
    int bits = 1;
    while (bits < 32 && (1ull << bits) <= vocab_size) ++bits;
    size_t tmp_bytes = 0;
    TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_idx, d_idx2, (int)n, 0,
                                        bits, st));
    size_t scan_bytes = 0;
    TRY(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, d_head, d_termof, (int)n, st));
    if (scan_bytes > tmp_bytes) tmp_bytes = scan_bytes;
    uint8_t* d_tmp;
    TRY(s.alloc(&d_tmp, tmp_bytes + 16));
    for (int j = (int)L - 1; j >= 0; --j) {  // stable LSD passes: last token first
      gather_word_kernel<<<grid, block, 0, st>>>(d_ids, d_idx, L, (uint32_t)j, d_keys, n);
      TRY(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_idx, d_idx2, (int)n, 0,
                                          bits, st));
      std::swap(d_idx, d_idx2);
    }
    heads_kernel<<<grid, block, 0, st>>>(d_ids, d_idx, L, n, d_head);
    TRY(cub::DeviceScan::InclusiveSum(d_tmp, tmp_bytes, d_head, d_termof, (int)n, st));
    uint32_t last = 0;
    TRY(cudaMemcpyAsync(&last, d_termof + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
    nterms = last;
  }
  // results (owned by the caller on success)
  BuiltTerms r;
  r.nterms = nterms;
  r.nposts = n;
  uint64_t* d_len = nullptr;
  TRY(s.alloc(&d_first, nterms));
  TRY(s.alloc(&d_len, nterms + 1));
  TRY(cudaMalloc(&r.d_term_off, (nterms + 1) * 8));
  TRY(cudaMalloc(&r.d_post_off, (nterms + 1) * 8));
  TRY(cudaMalloc(&r.d_post_rows, (n + 1) * 4));
  cudaError_t e = cudaSuccess;
  auto fail = [&](cudaError_t err) {
    cudaFree(r.d_term_off);
    cudaFree(r.d_post_off);
    cudaFree(r.d_post_rows);
    cudaFree(r.d_pool);
    return err;
  };
  if ((e = cudaMemsetAsync(d_len, 0, (nterms + 1) * 8, st)) != cudaSuccess) return fail(e);
  if (n) {
    scatter_kernel<<<grid, block, 0, st>>>(d_ids, d_idx, d_rows, d_head, d_termof, L, n, d_voff,
                                           r.d_post_rows, r.d_post_off, d_first, d_len);
    if ((e = cudaMemcpyAsync(r.d_post_off + nterms, &n, 8, cudaMemcpyHostToDevice, st)) != cudaSuccess)
      return fail(e);
  } else {
    uint64_t zero = 0;
    if ((e = cudaMemcpyAsync(r.d_post_off, &zero, 8, cudaMemcpyHostToDevice, st)) != cudaSuccess)
      return fail(e);
  }
  {  // term_off = exclusive scan of the lengths (nterms + 1 entries -> last = pool bytes)
    size_t scan_bytes = 0;
    if ((e = cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_len, r.d_term_off, (int)(nterms + 1),
                                           st)) != cudaSuccess)
      return fail(e);
    uint8_t* d_tmp2;
    if ((e = s.alloc(&d_tmp2, scan_bytes + 16)) != cudaSuccess) return fail(e);
    if ((e = cub::DeviceScan::ExclusiveSum(d_tmp2, scan_bytes, d_len, r.d_term_off, (int)(nterms + 1),
                                           st)) != cudaSuccess)
      return fail(e);
  }
  uint64_t pool_bytes = 0;
  if ((e = cudaMemcpyAsync(&pool_bytes, r.d_term_off + nterms, 8, cudaMemcpyDeviceToHost, st)) !=
          cudaSuccess ||
      (e = cudaStreamSynchronize(st)) != cudaSuccess)
    return fail(e);
  r.pool_bytes = pool_bytes;
  if ((e = cudaMalloc(&r.d_pool, pool_bytes + 16)) != cudaSuccess) return fail(e);
  if (nterms) {
    write_pool_kernel<<<grid, block, 0, st>>>(d_ids, d_first, L, nterms, d_vpool, d_voff, r.d_term_off,
                                              r.d_pool);
  }
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(e);
  if ((e = cudaGetLastError()) != cudaSuccess) return fail(e);
  *out = r;
  return cudaSuccess;
}

}  // namespace tss
