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

// Core: everything already resident in HBM (d_vpool/d_voff: the byte-sorted vocabulary; d_ids:
// n x L token ids, 0 = padding; d_rows: n).  The inputs are only read.
cudaError_t build_terms_core(const char* d_vpool, const uint64_t* d_voff, uint32_t vocab_size,
                             const uint32_t* d_ids, uint32_t L, const uint32_t* d_rows, uint64_t n,
                             cudaStream_t st, BuiltTerms* out) {
  Scratch s;
  const int grid = 148 * 4, block = 256;
  uint32_t *d_idx, *d_idx2, *d_keys, *d_keys2, *d_head, *d_termof, *d_first;
  TRY(s.alloc(&d_idx, n));
  TRY(s.alloc(&d_idx2, n));
  TRY(s.alloc(&d_keys, n));
  TRY(s.alloc(&d_keys2, n));
  TRY(s.alloc(&d_head, n));
  TRY(s.alloc(&d_termof, n));
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

cudaError_t build_terms_device(const char* vocab_pool, const uint64_t* vocab_off, uint32_t vocab_size,
                               const uint32_t* token_ids, uint32_t L, const uint32_t* rows,
                               uint64_t n, cudaStream_t st, BuiltTerms* out) {
  Scratch s;
  char* d_vpool;
  uint64_t* d_voff;
  uint32_t *d_ids, *d_rows;
  const uint64_t vbytes = vocab_off[vocab_size];
  TRY(s.alloc(&d_vpool, vbytes));
  TRY(s.alloc(&d_voff, (size_t)vocab_size + 1));
  TRY(s.alloc(&d_ids, n * L));
  TRY(s.alloc(&d_rows, n));
  if (vbytes) TRY(cudaMemcpyAsync(d_vpool, vocab_pool, vbytes, cudaMemcpyHostToDevice, st));
  TRY(cudaMemcpyAsync(d_voff, vocab_off, ((size_t)vocab_size + 1) * 8, cudaMemcpyHostToDevice, st));
  if (n) {
    TRY(cudaMemcpyAsync(d_ids, token_ids, n * L * 4, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_rows, rows, n * 4, cudaMemcpyHostToDevice, st));
  }
  return build_terms_core(d_vpool, d_voff, vocab_size, d_ids, L, d_rows, n, st, out);
}


// ---- N2 from raw strings: tokenise + dictionary-encode on the device ---------------------------
// What feeds TrieIndex::insert_* in the reference: `split_whitespace()` over the phrase and, for
// case names and content, `to_lowercase()` per token (src/trie.rs:147,158,171,177; citations keep
// their case, :190,196).  Here n phrases arrive as one byte buffer + offsets:
//   1. one thread per byte flags token starts (ASCII whitespace 0x09-0x0D, 0x20 separates; other
//      bytes < 0x20 are rejected, multi-byte UTF-8 passes through untouched -- case folding and
//      whitespace are ASCII, the deviation DESIGN.md section 5 records);
//   2. an inclusive scan numbers the tokens; each start records its position, length, phrase
//      (binary search of the offsets) and ordinal inside the phrase;
//   3. the tokens are sorted as strings by stable LSD radix passes over 8-byte big-endian chunks
//      (zero padded: no token holds a byte <= 0x20, so a proper prefix sorts first) -- as many
//      passes as the longest token needs;
//   4. adjacent-difference + scan rank the unique tokens = the byte-sorted vocabulary; ranks are
//      scattered into the n x L id matrix and the vocabulary pool is written;
//   5. build_terms_core (above) does the rest.
namespace {

constexpr uint32_t kMaxTokenBytes = 128;

__device__ __forceinline__ bool is_ws(unsigned char b) { return b == 0x20 || (b >= 0x09 && b <= 0x0D); }
__device__ __forceinline__ unsigned char fold(unsigned char b, int lower) {
  return (lower && b >= 'A' && b <= 'Z') ? (unsigned char)(b + 32) : b;
}
// largest p with off[p] <= pos  (off is non-decreasing, off[0] = 0 <= pos < off[n])
__device__ __forceinline__ uint64_t phrase_of(const uint64_t* off, uint64_t n, uint64_t pos) {
  uint64_t lo = 0, hi = n;  // invariant: off[lo] <= pos < off[hi]
  while (hi - lo > 1) {
    uint64_t mid = (lo + hi) >> 1;
    if (off[mid] <= pos) lo = mid;
    else hi = mid;
  }
  return lo;
}

__global__ void phrase_start_kernel(const uint64_t* off, uint64_t n, uint8_t* pstart) {
  for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < n;
       p += (uint64_t)gridDim.x * blockDim.x)
    if (off[p] < off[p + 1]) pstart[off[p]] = 1;
}
// err bits: 1 = control byte, 2 = token too long, 4 = too many tokens in a phrase
__global__ void token_start_kernel(const unsigned char* text, uint64_t total, const uint8_t* pstart,
                                   uint32_t* start, uint32_t* err) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned char b = text[i];
    if (b < 0x20 && !is_ws(b)) atomicOr(err, 1u);
    start[i] = (!is_ws(b) && (pstart[i] || is_ws(text[i - (i ? 1 : 0)]) || i == 0)) ? 1u : 0u;
  }
}
__global__ void token_info_kernel(const unsigned char* text, uint64_t total, const uint32_t* start,
                                  const uint32_t* tok_no, const uint64_t* off, uint64_t n, uint32_t L,
                                  uint64_t* tok_pos, uint32_t* tok_len, uint32_t* tok_slot,
                                  uint32_t* max_len, uint32_t* err) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    if (!start[i]) continue;
    const uint32_t t = tok_no[i] - 1;
    const uint64_t p = phrase_of(off, n, i);
    const uint64_t end = off[p + 1];
    uint32_t len = 1;
    while (i + len < end && !is_ws(text[i + len]) && len <= kMaxTokenBytes) ++len;
    if (len > kMaxTokenBytes) {
      atomicOr(err, 2u);
      len = kMaxTokenBytes;
    }
    const uint32_t first = off[p] ? tok_no[off[p] - 1] : 0u;  // tokens before this phrase
    const uint32_t ord = t - first;
    if (ord >= L) {
      atomicOr(err, 4u);
      tok_slot[t] = 0xFFFFFFFFu;
    } else {
      tok_slot[t] = (uint32_t)(p * L + ord);
    }
    tok_pos[t] = i;
    tok_len[t] = len;
    atomicMax(max_len, len);
  }
}
// key of token idx[j] for chunk c: bytes [8c, 8c+8) big-endian, folded, zero padded
__global__ void chunk_key_kernel(const unsigned char* text, const uint64_t* tok_pos,
                                 const uint32_t* tok_len, const uint32_t* idx, uint64_t m, uint32_t c,
                                 int lower, uint64_t* keys) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < m;
       j += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t t = idx[j];
    const unsigned char* s = text + tok_pos[t];
    const uint32_t len = tok_len[t];
    uint64_t k = 0;
#pragma unroll
    for (uint32_t b = 0; b < 8; ++b) {
      const uint32_t o = 8 * c + b;
      k = (k << 8) | (o < len ? (uint64_t)fold(s[o], lower) : 0ull);
    }
    keys[j] = k;
  }
}
__global__ void token_heads_kernel(const unsigned char* text, const uint64_t* tok_pos,
                                   const uint32_t* tok_len, const uint32_t* idx, uint64_t m, int lower,
                                   uint32_t* head) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < m;
       j += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t h = 1;
    if (j) {
      const uint32_t a = idx[j], b = idx[j - 1];
      if (tok_len[a] == tok_len[b]) {
        const unsigned char* sa = text + tok_pos[a];
        const unsigned char* sb = text + tok_pos[b];
        h = 0;
        for (uint32_t o = 0; o < tok_len[a]; ++o)
          if (fold(sa[o], lower) != fold(sb[o], lower)) {
            h = 1;
            break;
          }
      }
    }
    head[j] = h;
  }
}
// sorted token j has vocabulary rank rank[j] (1-based): write it into its slot of the id matrix;
// heads also record their token's length and index for the vocabulary pool
__global__ void assign_ids_kernel(const uint32_t* idx, const uint32_t* head, const uint32_t* rank,
                                  const uint32_t* tok_slot, const uint32_t* tok_len, uint64_t m,
                                  uint32_t* ids, uint64_t* vlen, uint32_t* vfirst) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < m;
       j += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t t = idx[j], r = rank[j];
    if (tok_slot[t] != 0xFFFFFFFFu) ids[tok_slot[t]] = r;
    if (head[j]) {
      vlen[r - 1] = tok_len[t];
      vfirst[r - 1] = t;
    }
  }
}
__global__ void write_vocab_kernel(const unsigned char* text, const uint64_t* tok_pos,
                                   const uint32_t* tok_len, const uint32_t* vfirst, const uint64_t* voff,
                                   uint64_t v, int lower, char* vpool) {
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < v;
       w += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t t = vfirst[w];
    const unsigned char* s = text + tok_pos[t];
    char* dst = vpool + voff[w];
    for (uint32_t o = 0; o < tok_len[t]; ++o) dst[o] = (char)fold(s[o], lower);
  }
}

}  // namespace

cudaError_t build_terms_from_text(const char* text, const uint64_t* phrase_off, const uint32_t* rows,
                                  uint64_t n, bool lowercase, uint32_t L, cudaStream_t st,
                                  BuiltTerms* out, uint32_t* err_bits, uint64_t* n_tokens,
                                  uint32_t* vocab_size_out) {
  Scratch s;
  const int grid = 148 * 4, block = 256;
  const int lower = lowercase ? 1 : 0;
  const uint64_t total = phrase_off[n];
  *err_bits = 0;
  unsigned char* d_text;
  uint64_t* d_off;
  uint32_t *d_rows, *d_ids, *d_start, *d_tokno, *d_flags;
  uint8_t* d_pstart;
  TRY(s.alloc(&d_text, total + 8));
  TRY(s.alloc(&d_off, n + 1));
  TRY(s.alloc(&d_rows, n));
  TRY(s.alloc(&d_ids, n * L));
  TRY(s.alloc(&d_pstart, total + 8));
  TRY(s.alloc(&d_start, total));
  TRY(s.alloc(&d_tokno, total));
  TRY(s.alloc(&d_flags, 4));  // [0] error bits, [1] longest token
  if (total) TRY(cudaMemcpyAsync(d_text, text, total, cudaMemcpyHostToDevice, st));
  TRY(cudaMemcpyAsync(d_off, phrase_off, (n + 1) * 8, cudaMemcpyHostToDevice, st));
  if (n) TRY(cudaMemcpyAsync(d_rows, rows, n * 4, cudaMemcpyHostToDevice, st));
  TRY(cudaMemsetAsync(d_ids, 0, (n * L ? n * L : 1) * 4, st));
  TRY(cudaMemsetAsync(d_pstart, 0, total + 8, st));
  TRY(cudaMemsetAsync(d_flags, 0, 16, st));
  uint64_t m = 0;  // tokens
  if (total && n) {
    phrase_start_kernel<<<grid, block, 0, st>>>(d_off, n, d_pstart);
    token_start_kernel<<<grid, block, 0, st>>>(d_text, total, d_pstart, d_start, d_flags);
    size_t scan_bytes = 0;
    TRY(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, d_start, d_tokno, (int)total, st));
    uint8_t* d_tmp;
    TRY(s.alloc(&d_tmp, scan_bytes + 16));
    TRY(cub::DeviceScan::InclusiveSum(d_tmp, scan_bytes, d_start, d_tokno, (int)total, st));
    uint32_t last = 0;
    TRY(cudaMemcpyAsync(&last, d_tokno + (total - 1), 4, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
    m = last;
  }
  *n_tokens = m;
  uint64_t* d_tpos;
  uint32_t *d_tlen, *d_tslot, *d_idx, *d_idx2, *d_head, *d_rank, *d_vfirst;
  uint64_t *d_keys, *d_keys2, *d_vlen, *d_voff;
  TRY(s.alloc(&d_tpos, m));
  TRY(s.alloc(&d_tlen, m));
  TRY(s.alloc(&d_tslot, m));
  TRY(s.alloc(&d_idx, m));
  TRY(s.alloc(&d_idx2, m));
  TRY(s.alloc(&d_keys, m));
  TRY(s.alloc(&d_keys2, m));
  TRY(s.alloc(&d_head, m));
  TRY(s.alloc(&d_rank, m));
  uint32_t v = 0;
  if (m) {
    token_info_kernel<<<grid, block, 0, st>>>(d_text, total, d_start, d_tokno, d_off, n, L, d_tpos,
                                              d_tlen, d_tslot, d_flags + 1, d_flags);
    uint32_t fl[2] = {0, 0};
    TRY(cudaMemcpyAsync(fl, d_flags, 8, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
    if (fl[0]) {
      *err_bits = fl[0];
      return cudaSuccess;  // the caller reports which rule the text broke
    }
    const uint32_t chunks = (fl[1] + 7) / 8;
    iota_kernel<<<grid, block, 0, st>>>(d_idx, m);
    size_t sort_bytes = 0, scan_bytes = 0;
    TRY(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, d_keys, d_keys2, d_idx, d_idx2, (int)m, 0,
                                        64, st));
    TRY(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, d_head, d_rank, (int)m, st));
    if (scan_bytes > sort_bytes) sort_bytes = scan_bytes;
    uint8_t* d_tmp;
    TRY(s.alloc(&d_tmp, sort_bytes + 16));
    for (int c = (int)chunks - 1; c >= 0; --c) {  // stable LSD over the chunks: last chunk first
      chunk_key_kernel<<<grid, block, 0, st>>>(d_text, d_tpos, d_tlen, d_idx, m, (uint32_t)c, lower,
                                               d_keys);
      TRY(cub::DeviceRadixSort::SortPairs(d_tmp, sort_bytes, d_keys, d_keys2, d_idx, d_idx2, (int)m, 0,
                                          64, st));
      std::swap(d_idx, d_idx2);
    }
    token_heads_kernel<<<grid, block, 0, st>>>(d_text, d_tpos, d_tlen, d_idx, m, lower, d_head);
    TRY(cub::DeviceScan::InclusiveSum(d_tmp, sort_bytes, d_head, d_rank, (int)m, st));
    TRY(cudaMemcpyAsync(&v, d_rank + (m - 1), 4, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
  } else if (total && n) {
    uint32_t fl = 0;
    TRY(cudaMemcpyAsync(&fl, d_flags, 4, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
    if (fl) {
      *err_bits = fl;
      return cudaSuccess;
    }
  }
  *vocab_size_out = v;
  char* d_vpool = nullptr;
  TRY(s.alloc(&d_vlen, (size_t)v + 1));
  TRY(s.alloc(&d_voff, (size_t)v + 1));
  TRY(s.alloc(&d_vfirst, v));
  TRY(cudaMemsetAsync(d_vlen, 0, ((size_t)v + 1) * 8, st));
  if (m)
    assign_ids_kernel<<<grid, block, 0, st>>>(d_idx, d_head, d_rank, d_tslot, d_tlen, m, d_ids, d_vlen,
                                              d_vfirst);
  {
    size_t scan_bytes = 0;
    TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_vlen, d_voff, (int)(v + 1), st));
    uint8_t* d_tmp2;
    TRY(s.alloc(&d_tmp2, scan_bytes + 16));
    TRY(cub::DeviceScan::ExclusiveSum(d_tmp2, scan_bytes, d_vlen, d_voff, (int)(v + 1), st));
  }
  uint64_t vbytes = 0;
  TRY(cudaMemcpyAsync(&vbytes, d_voff + v, 8, cudaMemcpyDeviceToHost, st));
  TRY(cudaStreamSynchronize(st));
  TRY(s.alloc(&d_vpool, vbytes));
  if (v)
    write_vocab_kernel<<<grid, block, 0, st>>>(d_text, d_tpos, d_tlen, d_vfirst, d_voff, v, lower,
                                               d_vpool);
  TRY(cudaGetLastError());
  return build_terms_core(d_vpool, d_voff, v, d_ids, L, d_rows, n, st, out);
}

}  // namespace tss
