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

// N2 from raw strings: n phrases (text + phrase_off[n + 1], HOST), one posting each (rows[n]);
// tokens = maximal runs of non-whitespace bytes (ASCII whitespace), optionally ASCII lower-cased;
// at most L tokens per phrase and 128 bytes per token.  *err_bits != 0 (1 control byte, 2 token
// too long, 4 too many tokens) means the text broke a rule and nothing was built.
cudaError_t build_terms_from_text(const char* text, const uint64_t* phrase_off, const uint32_t* rows,
                                  uint64_t n, bool lowercase, uint32_t L, cudaStream_t st,
                                  BuiltTerms* out, uint32_t* err_bits, uint64_t* n_tokens,
                                  uint32_t* vocab_size_out);

}  // namespace tss
