// terms_build.cuh -- N2: flattened-trie construction on the device (see terms_build.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tss {

struct BuiltTerms {  // device arrays owned by the caller after a successful build
  char* d_pool = nullptr;
  uint64_t* d_term_off = nullptr;  // [nterms + 1]
  uint64_t* d_post_off = nullptr;  // [nterms + 1]
  uint32_t* d_post_rows = nullptr; // [nposts]
  uint64_t nterms = 0, pool_bytes = 0, nposts = 0;
};

// token_ids: n x L u32 on the HOST (0 = padding, id = vocabulary index + 1); vocabulary byte-sorted.
cudaError_t build_terms_device(const char* vocab_pool, const uint64_t* vocab_off, uint32_t vocab_size,
                               const uint32_t* token_ids, uint32_t L, const uint32_t* rows,
                               uint64_t n, cudaStream_t st, BuiltTerms* out);

}  // namespace tss
