/*
 * oracle.h -- CPU restatement of the trie-semantic-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker / CPU arm.
 *
 * PARITY UNPINNED: the reference's scoring path is an unimplemented stub
 * (src/vector.rs:190-207 -- add_vector drops its input, search returns an
 * empty Vec, size returns 0; src/vector.rs:173 embeds every text as 768
 * zeros), its arithmetic crates are commented out (Cargo.toml:37,40), the
 * crate does not compile, and its only tests are src/utils.rs:205-227
 * (unrelated).  So no golden vector exists for scoring.  What IS pinned by
 * reading the reference: the token-level trie (src/trie.rs:139-278), the
 * hybrid merge (src/search.rs:185-240,255-274), the result mapping
 * (src/vector.rs:128-150) and the defaults (src/lib.rs:122-145).  The
 * arithmetic (canonical fp32 reduction order, zero-norm rule, tie-break)
 * is defined in DESIGN.md section 3 and restated in oracle.cpp; that
 * restatement is itself held, bit for bit, to an evaluation of the
 * definition in exact rational arithmetic (tests/test_exact_arithmetic.py).
 */
#ifndef TSS_ORACLE_H
#define TSS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- synthetic data (same integer hash as the device generator) -------- */

/* rows [row_begin, row_begin+nrows) of the seeded synthetic corpus, written
 * unpadded row-major (dim floats per row). */
void orc_gen_rows(float* out, uint64_t row_begin, uint64_t nrows, uint32_t dim,
                  uint64_t seed);

/* ---- scoring + top-k (restates the contract of HnswIndex::search,
 *      src/vector.rs:195-202, with DESIGN.md section 3 arithmetic) ---------- */

enum { ORC_ORDER_CANONICAL = 0, ORC_ORDER_SEQUENTIAL = 1 };
enum { ORC_MASK_NONE = 0, ORC_MASK_INCLUDE = 1, ORC_MASK_EXCLUDE = 2 };

/* cosine score of every row against one query; out_scores has n floats. */
void orc_scores(const float* rows, uint64_t n, uint32_t dim, const float* query,
                float* out_scores, int order, int threads);

/* bf16 variant: rows are RNE-rounded to bf16 first (storage rounding of the
 * bf16 index), query stays fp32, fp32 accumulate in canonical order. */
void orc_scores_bf16(const float* rows, uint64_t n, uint32_t dim,
                     const float* query, float* out_scores, int threads);

/* exact top-k, (score desc, row asc).  mask_words: bit r of the bitmask is
 * (mask_words[r>>5] >> (r&31)) & 1, indexed by LOCAL row (0..n).  row_base is
 * added to the reported row ids.  Outputs are nq*k, unused slots are
 * row=0xFFFFFFFF, score=0; out_counts[q] = number of valid slots. */
int orc_cosine_topk(const float* rows, uint64_t n, uint32_t dim,
                    const float* queries, uint32_t nq, uint32_t k,
                    const uint32_t* mask_words, int mask_mode, uint64_t row_base,
                    uint32_t* out_rows, float* out_scores, uint32_t* out_counts,
                    int order, int storage_bf16, int threads);

/* top-k straight from the seeded generator (no matrix in memory): rows
 * [row_begin,row_begin+nrows) are generated tile by tile.  Used as the CPU
 * baseline on corpora too large to hold and for full-size spot checks. */
int orc_cosine_topk_synth(uint64_t row_begin, uint64_t nrows, uint32_t dim,
                          uint64_t seed, const float* queries, uint32_t nq,
                          uint32_t k, uint32_t* out_rows, float* out_scores,
                          uint32_t* out_counts, int threads);

/* 64-bit ordering key used everywhere: (orderable(score) << 32) | ~row. */
uint64_t orc_pack_key(float score, uint32_t row);

/* ---- token trie: literal restatement of src/trie.rs ---------------------- */

typedef struct orc_docref {
  uint8_t case_id[16];      /* Uuid, src/lib.rs:65 */
  uint64_t paragraph_index; /* src/lib.rs:72 */
  int64_t char_offset;      /* Option<usize>: -1 = None, src/lib.rs:74 */
} orc_docref;

typedef struct orc_trie_index orc_trie_index; /* TrieIndex, src/trie.rs:28-33 */

enum { ORC_TRIE_CASE_NAME = 0, ORC_TRIE_CONTENT = 1, ORC_TRIE_CITATION = 2 };

orc_trie_index* orc_trie_new(void);
void orc_trie_free(orc_trie_index*);
/* src/trie.rs:97-109 */
void orc_trie_insert_case_name(orc_trie_index*, const char* name, const uint8_t case_id[16]);
void orc_trie_insert_content(orc_trie_index*, const char* const* tokens, uint32_t ntokens,
                             const orc_docref* ref);
void orc_trie_insert_citation(orc_trie_index*, const char* citation, const orc_docref* ref);

typedef struct orc_trie_result {
  orc_docref* exact_matches; /* insertion order, duplicates kept */
  uint64_t n_exact;
  char** completions; /* <= limit strings, byte-sorted (reference order is HashMap order) */
  uint64_t n_completions;
  uint64_t n_completions_unlimited; /* how many the DFS would emit with no limit */
  uint64_t total_matches;           /* n_exact + n_completions, src/trie.rs:251 */
  uint32_t frequency;               /* of the node reached; 0 if miss */
} orc_trie_result;

/* TrieNode::search on ONE trie (src/trie.rs:223-255) with that trie's
 * tokenisation (src/trie.rs:147,158,171,177,190,196). */
orc_trie_result* orc_trie_search_one(const orc_trie_index*, int which, const char* query);
/* TrieIndex::search cascade (src/trie.rs:112-130). */
orc_trie_result* orc_trie_search(const orc_trie_index*, const char* query);
void orc_trie_result_free(orc_trie_result*);

/* prefix posting set: postings of the node reached by the
 * query tokens plus every terminal below it, no limit.  Returns count and
 * a malloc'ed array (free with orc_free). */
uint64_t orc_trie_prefix_postings(const orc_trie_index*, int which, const char* query,
                                  orc_docref** out);
void orc_free(void*);

/* ---- hybrid merge on integer case ids (src/search.rs:185-240) ------------- */

typedef struct orc_hit {
  uint64_t case_id;
  float score;
  int match_type; /* 0 = Exact, 2 = Semantic (MatchType, src/search.rs:71-82) */
} orc_hit;

/* exact_cases: case ids of trie exact matches in order; vec_cases/vec_scores:
 * vector hits best first.  Returns number of results written (<= cap). */
uint32_t orc_hybrid_merge(const uint64_t* exact_cases, uint32_t n_exact,
                          const uint64_t* vec_cases, const float* vec_scores,
                          uint32_t n_vec, int enable_prefix, int enable_semantic,
                          uint32_t cfg_max_results, int64_t query_max_results,
                          float min_similarity, float exact_match_weight,
                          orc_hit* out, uint32_t cap);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
