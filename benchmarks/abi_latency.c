/*
 * abi_latency.c -- what a compiled host (the reference's Rust, through ffi/tss.rs) sees when it
 * calls libtss: per-call latency of the C ABI with host pointers, no interpreter in the way.
 * bench.py's `e2e` figures are taken from Python through ctypes; this program times the same
 * calls from C, one JSON line on stdout.
 *
 *   make -C trie-semantic-search_b200/host      (or: gcc -O2 benchmarks/abi_latency.c -Iinclude \
 *       -Ltrie-semantic-search_b200 -ltss -Wl,-rpath,'$ORIGIN/../trie-semantic-search_b200' -lm
 *   trie-semantic-search_b200/build/abi_latency [rows=10000000] [iters=300]
 *
 * Legs (all k = 10, D = 384, fp32 index filled by the seeded device generator):
 *   blocking    tss_index_search, one query at a time
 *   pipelined   tss_index_search_submit / _collect, two in flight
 *   hybrid2     tss_prefix_mask_fresh + tss_index_search(INCLUDE), ~10k live rows
 *   hybrid1     tss_index_search_prefix (the same as one call)
 *   tiny        tss_index_search over a 1 000-row index: the fixed cost of a call
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "tss.h"

static double now_us(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

#define CK(call)                                                          \
  do {                                                                    \
    int rc_ = (call);                                                     \
    if (rc_ != TSS_OK) {                                                  \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, tss_last_error());   \
      return 1;                                                           \
    }                                                                     \
  } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void) {
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return (uint32_t)(rng_state >> 32);
}

int main(int argc, char** argv) {
  const uint64_t rows = argc > 1 ? strtoull(argv[1], NULL, 10) : 10000000ull;
  const int iters = argc > 2 ? atoi(argv[2]) : 300;
  const uint32_t dim = 384, k = 10;
  if (tss_device_count() < 1) {
    fprintf(stderr, "no CUDA device: libtss has no CPU path\n");
    return 1;
  }
  tss_index *ix = NULL, *tiny = NULL;
  CK(tss_index_create(&ix, dim, TSS_F32, 0));
  CK(tss_index_reserve(ix, rows));
  CK(tss_index_add_synthetic(ix, 0, rows, 0x5EED));
  CK(tss_index_finalize(ix));
  CK(tss_index_create(&tiny, dim, TSS_F32, 0));
  CK(tss_index_add_synthetic(tiny, 0, 1000, 0x5EED));
  CK(tss_index_finalize(tiny));

  const int nq = 16;
  float* q = (float*)malloc(sizeof(float) * dim * nq);
  for (uint32_t i = 0; i < dim * nq; ++i) q[i] = (float)((int)(rnd() % 2001) - 1000) / 1000.0f;
  uint32_t out_rows[2][10], out_counts[2];
  float out_scores[2][10];

  /* flattened trie: 1000 first tokens x 10 second tokens, 1000 postings each (10M postings) */
  const uint64_t T = 10000, per = 1000;
  char* pool = (char*)malloc(T * 8 + 1); /* sprintf ends each term with a NUL the next one overwrites */
  uint64_t* term_off = (uint64_t*)malloc(sizeof(uint64_t) * (T + 1));
  uint64_t* post_off = (uint64_t*)malloc(sizeof(uint64_t) * (T + 1));
  uint32_t* post_rows = (uint32_t*)malloc(sizeof(uint32_t) * T * per);
  uint64_t pb = 0;
  for (uint64_t t = 0; t < T; ++t) {
    term_off[t] = pb;
    pb += (uint64_t)sprintf(pool + pb, "w%04d t%d", (int)(t / 10), (int)(t % 10));
    post_off[t] = t * per;
    for (uint64_t j = 0; j < per; ++j) post_rows[t * per + j] = (uint32_t)(rnd() % rows);
  }
  term_off[T] = pb;
  post_off[T] = T * per;
  tss_terms* terms = NULL;
  tss_mask* mask = NULL;
  CK(tss_terms_create(&terms, pool, term_off, post_off, post_rows, T, 0));
  CK(tss_mask_create(&mask, rows, 0));
  CK(tss_terms_bind_stream(terms, ix));
  const char* prefix = "w0042";
  tss_prefix_stats st;
  CK(tss_prefix_mask_fresh(terms, prefix, 5, TSS_PREFIX_TOKEN, mask, 0, &st));
  uint64_t live = 0;
  CK(tss_mask_popcount(mask, &live));

  double t0, blocking, pipelined, hybrid2, hybrid1, tiny_us;
  /* blocking */
  for (int i = 0; i < 5; ++i)
    CK(tss_index_search(ix, q + (i % nq) * dim, 1, k, NULL, TSS_MASK_NONE, out_rows[0], out_scores[0], out_counts));
  t0 = now_us();
  for (int i = 0; i < iters; ++i)
    CK(tss_index_search(ix, q + (i % nq) * dim, 1, k, NULL, TSS_MASK_NONE, out_rows[0], out_scores[0], out_counts));
  blocking = (now_us() - t0) / iters;
  /* two in flight */
  uint64_t tk[2] = {0, 0};
  t0 = now_us();
  for (int i = 0; i < iters; ++i) {
    if (i >= 2) CK(tss_index_search_collect(ix, tk[i & 1], out_rows[i & 1], out_scores[i & 1], out_counts + (i & 1)));
    CK(tss_index_search_submit(ix, q + (i % nq) * dim, 1, k, NULL, TSS_MASK_NONE, &tk[i & 1]));
  }
  for (int i = iters; i < iters + 2 && i >= 2; ++i)
    CK(tss_index_search_collect(ix, tk[i & 1], out_rows[i & 1], out_scores[i & 1], out_counts + (i & 1)));
  pipelined = (now_us() - t0) / iters;
  /* hybrid, two calls */
  const int hit = iters * 4;
  for (int i = 0; i < 5; ++i) {
    CK(tss_prefix_mask_fresh(terms, prefix, 5, TSS_PREFIX_TOKEN, mask, 0, NULL));
    CK(tss_index_search(ix, q, 1, k, mask, TSS_MASK_INCLUDE, out_rows[0], out_scores[0], out_counts));
  }
  t0 = now_us();
  for (int i = 0; i < hit; ++i) {
    CK(tss_prefix_mask_fresh(terms, prefix, 5, TSS_PREFIX_TOKEN, mask, 0, NULL));
    CK(tss_index_search(ix, q + (i % nq) * dim, 1, k, mask, TSS_MASK_INCLUDE, out_rows[0], out_scores[0], out_counts));
  }
  hybrid2 = (now_us() - t0) / hit;
  uint32_t keep[10];
  memcpy(keep, out_rows[0], sizeof(keep));
  /* hybrid, one call */
  t0 = now_us();
  for (int i = 0; i < hit; ++i)
    CK(tss_index_search_prefix(ix, terms, prefix, 5, TSS_PREFIX_TOKEN, mask, q + (i % nq) * dim, 1, k,
                               out_rows[0], out_scores[0], out_counts));
  hybrid1 = (now_us() - t0) / hit;
  const int same = memcmp(keep, out_rows[0], sizeof(keep)) == 0;
  /* fixed cost of a call */
  for (int i = 0; i < 5; ++i)
    CK(tss_index_search(tiny, q, 1, k, NULL, TSS_MASK_NONE, out_rows[0], out_scores[0], out_counts));
  t0 = now_us();
  for (int i = 0; i < hit; ++i)
    CK(tss_index_search(tiny, q + (i % nq) * dim, 1, k, NULL, TSS_MASK_NONE, out_rows[0], out_scores[0], out_counts));
  tiny_us = (now_us() - t0) / hit;

  printf("{\"what\": \"C-ABI latency from a compiled host (C), host pointers\", \"rows\": %llu, "
         "\"dim\": %u, \"k\": %u, \"iters\": %d, \"blocking_us\": %.2f, \"blocking_qps\": %.1f, "
         "\"pipelined2_us\": %.2f, \"pipelined2_qps\": %.1f, \"hybrid_live_rows\": %llu, "
         "\"hybrid_postings\": %llu, \"hybrid_two_calls_us\": %.2f, \"hybrid_one_call_us\": %.2f, "
         "\"hybrid_results_equal\": %s, \"tiny_index_call_us\": %.2f}\n",
         (unsigned long long)rows, dim, k, iters, blocking, 1e6 / blocking, pipelined, 1e6 / pipelined,
         (unsigned long long)live, (unsigned long long)st.npostings, hybrid2, hybrid1,
         same ? "true" : "false", tiny_us);
  tss_terms_bind_stream(terms, NULL);
  tss_terms_destroy(terms);
  tss_mask_destroy(mask);
  tss_index_destroy(tiny);
  tss_index_destroy(ix);
  free(q); free(pool); free(term_off); free(post_off); free(post_rows);
  return same ? 0 : 2;
}
