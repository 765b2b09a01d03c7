// aux_kernels.cuh -- host-callable wrappers of the small kernels around the scan.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tss {

// rows [row_begin,row_begin+nrows) of the seeded synthetic corpus written into
// padded storage (stride_elems elements per row, fp32 or bf16).
cudaError_t launch_synth_fill(void* dst, uint64_t row_begin, uint64_t nrows, uint32_t dim,
                              uint32_t stride_elems, bool bf16, uint64_t seed, cudaStream_t st);
// *flag |= 1 if any of count floats is NaN/Inf
cudaError_t launch_check_finite(const float* src, uint64_t count, int* flag, cudaStream_t st);
// unpadded fp32 [nrows][dim] -> padded storage rows (fp32 or bf16 RNE)
cudaError_t launch_pack_rows(const float* src, void* dst, uint64_t nrows, uint32_t dim,
                             uint32_t stride_elems, bool bf16, cudaStream_t st);
// padded storage rows (fp32 or bf16) -> bf16 rows of the same stride, each scaled to unit length
// first (fp32 sum of squares, rsqrt, fp32 multiply, RNE): the matrix the tensor cores read, so
// that a raw dot product IS the cosine numerator and the epilogue needs no per-row weight.  A
// zero row stays zero.
// *d_max_err (nullable) <- an upper bound of max over rows of |e_n - e/|e|| (float bits).
cudaError_t launch_normalize_rows(const void* src, bool src_bf16, void* dst_bf16, uint64_t nrows,
                                  uint32_t stride_elems, unsigned int* d_max_err, cudaStream_t st);
// padded storage -> unpadded fp32
cudaError_t launch_unpack_rows(const void* src, float* dst, uint64_t nrows, uint32_t dim,
                               uint32_t stride_elems, bool bf16, cudaStream_t st);

// masks
cudaError_t launch_mask_set_rows(uint32_t* words, uint64_t nbits, const uint32_t* rows, uint64_t n,
                                 uint64_t row_base, bool set, cudaStream_t st);
// the same for a short list that travels in the kernel parameters (no staging copy, no sync)
constexpr uint32_t kInlineRows = 896;
struct InlineRows {
  uint32_t n;
  uint32_t rows[kInlineRows];
};
cudaError_t launch_mask_set_rows_inline(uint32_t* words, uint64_t nbits, const InlineRows& rows,
                                        uint64_t row_base, bool set, cudaStream_t st);
// N3: metadata columns -> mask words (overwrite or AND)
cudaError_t launch_filter_mask(const uint16_t* court, const int32_t* date, uint64_t nrows,
                               const uint32_t* allow_bits, bool any_court, int32_t lo, int32_t hi,
                               uint32_t* words, bool combine_and, cudaStream_t st);
cudaError_t launch_mask_update_from_keys(uint32_t* words, uint64_t nbits, const uint64_t* keys,
                                         uint32_t n, uint64_t row_base, bool set, cudaStream_t st);
cudaError_t launch_mask_popcount(const uint32_t* words, uint64_t nwords, unsigned long long* out,
                                 cudaStream_t st);

// K4: prefix search over the flattened term array + posting scatter
struct TermsDev {
  const char* pool;
  const uint64_t* term_off;  // T+1
  const uint64_t* post_off;  // T+1
  const uint32_t* post_rows;
  uint64_t nterms;
};
constexpr uint32_t kPrefixInlineBytes = 1024;
struct PrefixKeys {  // up to 4 lower-bound probes, keys concatenated
  uint32_t off[5];
  int32_t fixed[4];  // >= 0: bound is this constant (no search); -1: search; -2: nterms
  // the concatenated key bytes ride in the kernel parameters when they fit (prefixes up to 255
  // bytes): no staging copy, so nothing on the host has to wait for the previous call
  char bytes[kPrefixInlineBytes];
};
// *bad |= 1 unless: offsets start at 0, are monotone and end on pool_bytes / nposts, and the terms
// are strictly increasing in byte order (what the prefix search relies on).  For loaded files.
cudaError_t launch_terms_validate(const TermsDev& t, uint64_t pool_bytes, uint64_t nposts,
                                  unsigned int* d_bad, cudaStream_t st);

// K4 as ONE launch (what tss_prefix_mask / tss_prefix_mask_fresh enqueue): CTAs 0..3 each find
// one lower bound with a 256-ary cooperative search while the other CTAs clear the mask
// (clear_nwords > 0); a grid barrier; then every CTA scatters the posting ranges.  When `list`
// is given and the ranges hold at most list_cap postings, the threads whose atomicOr flipped a
// bit also append the row to `list` (unique local rows, *list_count of them): the masked scan
// then fetches exactly those rows instead of walking the mask.  *list_count = 0xFFFFFFFF
// otherwise.  sync: two zeroed words used by the grid barrier (left zeroed).
struct PrefixMaskArgs {
  TermsDev t;
  const char* d_keybytes;   // null: keys.bytes
  uint64_t* bounds;         // [4] out (+ [4] = postings visited)
  uint32_t* words;          // the mask
  uint64_t clear_nwords;    // 0: OR into the mask as it is
  uint64_t nbits, row_base;
  uint32_t* list;           // nullable
  uint32_t* list_count;
  uint32_t list_cap;
  unsigned int* sync;
};
cudaError_t launch_prefix_mask(const PrefixMaskArgs& a, const PrefixKeys& keys, int num_sms,
                               cudaStream_t st);

// Device-side fix-up list of a K2 batch: the indices of the queries whose flag is set, compacted
// (ascending) into d_list / *d_count for the guarded scans that follow, mirrored into mapped
// host memory (h_list / *h_count: read by the host only when it synchronises anyway), and
// *h_sticky |= 2 when there are more of them than `fixups` guarded launches will serve.
cudaError_t launch_redo_compact(const uint32_t* flags, uint32_t nq, uint32_t* d_list,
                                uint32_t* d_count, uint32_t* h_list, uint32_t* h_count,
                                uint32_t fixups, unsigned int* h_sticky, cudaStream_t st);

// K5: merge P gathered lists of k keys per query: in [P][nq][k] -> out [nq][k]
cudaError_t launch_merge_gathered(const uint64_t* in, uint64_t* out, uint32_t P, uint32_t nq,
                                  uint32_t k, cudaStream_t st);

}  // namespace tss
