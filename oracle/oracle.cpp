/*
 * oracle.cpp -- CPU restatement of the trie-semantic-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED for scoring: the
 * reference has no implementation and no golden vectors on this path
 * (src/vector.rs:190-207 are stubs).  Trie and merge semantics follow the
 * reference source line by line; citations are into /root/reference.
 *
 * Build: see oracle/Makefile (g++ -O3 -mavx2 -mfma -ffp-contract=off -fopenmp).
 * -ffp-contract=off matters: every fused multiply-add below is an explicit
 * fmaf() so the rounding sequence is exactly the one DESIGN.md section 3 states
 * and the CUDA kernels reproduce.
 */
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <queue>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

/* ------------------------------------------------------------------------ */
/* synthetic generator: one splitmix64 finaliser per element pair; each     */
/* element is a sum of two 16-bit uniforms, centred, scaled by 2^-16 --     */
/* every value is a 17-bit integer times a power of two, exact in fp32,     */
/* so host and device produce identical bits with no libm involved.         */
/* ------------------------------------------------------------------------ */
inline uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline float synth_from_u32(uint32_t h) {
  int v = (int)(h & 0xFFFFu) + (int)(h >> 16) - 65535;
  return (float)v * (1.0f / 65536.0f);
}
inline void synth_pair(uint64_t seed, uint64_t row, uint32_t pair, float* a, float* b) {
  uint64_t ctr = (row << 16) | (uint64_t)pair;
  uint64_t h = mix64(ctr + seed * 0x9E3779B97F4A7C15ull);
  *a = synth_from_u32((uint32_t)h);
  *b = synth_from_u32((uint32_t)(h >> 32));
}
void gen_row(float* out, uint64_t row, uint32_t dim, uint64_t seed) {
  for (uint32_t j = 0; j < dim; j += 2) {
    float a, b;
    synth_pair(seed, row, j >> 1, &a, &b);
    out[j] = a;
    if (j + 1 < dim) out[j + 1] = b;
  }
}

/* ------------------------------------------------------------------------ */
/* canonical fp32 reduction.  A row is viewed as stripes of  */
/* 128 elements (zero padded).  "Lane" l in 0..31 owns elements 4l..4l+3 of */
/* every stripe; its four sub-accumulators each take one element per stripe */
/* with fmaf, in stripe order; they are combined (a0+a1)+(a2+a3); the 32    */
/* lane partials go through the xor-butterfly 16,8,4,2,1.                    */
/* ------------------------------------------------------------------------ */
inline float butterfly32(const float* p) {
  float t[16];
  for (int l = 0; l < 16; ++l) t[l] = p[l] + p[l + 16];
  for (int l = 0; l < 8; ++l) t[l] = t[l] + t[l + 8];
  for (int l = 0; l < 4; ++l) t[l] = t[l] + t[l + 4];
  for (int l = 0; l < 2; ++l) t[l] = t[l] + t[l + 2];
  return t[0] + t[1];
}

inline float bf16_round(float x) { /* RNE to bf16, returned widened */
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) { /* inf/nan: truncate */
    u &= 0xFFFF0000u;
  } else {
    u += 0x7FFFu + ((u >> 16) & 1u);
    u &= 0xFFFF0000u;
  }
  float r;
  memcpy(&r, &u, 4);
  return r;
}

/* dot and squared norm of one row (dim floats, read unpadded) against a
 * padded query qp (stripes*128 floats). */
inline void canon_dot_norm(const float* e, uint32_t dim, const float* qp, float* dot,
                           float* nrm, bool bf16) {
  float ad[128], an[128];
  for (int i = 0; i < 128; ++i) ad[i] = 0.0f, an[i] = 0.0f;
  uint32_t full = dim / 128;
  for (uint32_t s = 0; s < full; ++s) {
    const float* es = e + 128 * s;
    const float* qs = qp + 128 * s;
    if (bf16) {
      for (int i = 0; i < 128; ++i) {
        float v = bf16_round(es[i]);
        ad[i] = __builtin_fmaf(qs[i], v, ad[i]);
        an[i] = __builtin_fmaf(v, v, an[i]);
      }
    } else {
      for (int i = 0; i < 128; ++i) {
        ad[i] = __builtin_fmaf(qs[i], es[i], ad[i]);
        an[i] = __builtin_fmaf(es[i], es[i], an[i]);
      }
    }
  }
  uint32_t rem = dim - full * 128;
  if (rem) { /* zero padded tail stripe: fmaf(q,0,acc) still executes */
    const float* es = e + 128 * full;
    const float* qs = qp + 128 * full;
    for (int i = 0; i < 128; ++i) {
      float v = (uint32_t)i < rem ? es[i] : 0.0f;
      if (bf16) v = bf16_round(v);
      ad[i] = __builtin_fmaf(qs[i], v, ad[i]);
      an[i] = __builtin_fmaf(v, v, an[i]);
    }
  }
  float pd[32], pn[32];
  for (int l = 0; l < 32; ++l) {
    pd[l] = (ad[4 * l] + ad[4 * l + 1]) + (ad[4 * l + 2] + ad[4 * l + 3]);
    pn[l] = (an[4 * l] + an[4 * l + 1]) + (an[4 * l + 2] + an[4 * l + 3]);
  }
  *dot = butterfly32(pd);
  *nrm = butterfly32(pn);
}

inline float canon_self_norm(const float* qp, uint32_t stripes) {
  float an[128];
  for (int i = 0; i < 128; ++i) an[i] = 0.0f;
  for (uint32_t s = 0; s < stripes; ++s)
    for (int i = 0; i < 128; ++i) an[i] = __builtin_fmaf(qp[128 * s + i], qp[128 * s + i], an[i]);
  float pn[32];
  for (int l = 0; l < 32; ++l)
    pn[l] = (an[4 * l] + an[4 * l + 1]) + (an[4 * l + 2] + an[4 * l + 3]);
  return butterfly32(pn);
}

/* score rule: zero / non-finite denominators and non-finite
 * quotients give 0.0; -0.0 is canonicalised to +0.0. */
inline float finish_score(float dot, float nq2, float ne2) {
  float den = sqrtf(nq2) * sqrtf(ne2);
  float s = dot / den;
  if (!(den > 0.0f) || !std::isfinite(s)) s = 0.0f;
  if (s == 0.0f) s = 0.0f;
  return s;
}

inline uint32_t orderable(float s) {
  uint32_t u;
  memcpy(&u, &s, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
inline uint64_t pack_key(float s, uint32_t row) {
  return ((uint64_t)orderable(s) << 32) | (uint64_t)(0xFFFFFFFFu - row);
}
inline float key_score(uint64_t key) {
  uint32_t o = (uint32_t)(key >> 32);
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  float s;
  memcpy(&s, &u, 4);
  return s;
}
inline uint32_t key_row(uint64_t key) { return 0xFFFFFFFFu - (uint32_t)key; }

struct QueryPrep {
  std::vector<float> qp; /* padded */
  float nq2;
};
QueryPrep prep_query(const float* q, uint32_t dim) {
  QueryPrep p;
  uint32_t stripes = (dim + 127) / 128;
  p.qp.assign((size_t)stripes * 128, 0.0f);
  memcpy(p.qp.data(), q, sizeof(float) * dim);
  p.nq2 = canon_self_norm(p.qp.data(), stripes);
  return p;
}

inline float seq_score(const float* e, uint32_t dim, const float* q) {
  /* plain left-to-right sums, no fma: the tolerance-only second opinion */
  float dot = 0.0f, nq2 = 0.0f, ne2 = 0.0f;
  for (uint32_t j = 0; j < dim; ++j) {
    float prod = q[j] * e[j];
    dot = dot + prod;
    float qq = q[j] * q[j];
    nq2 = nq2 + qq;
    float ee = e[j] * e[j];
    ne2 = ne2 + ee;
  }
  return finish_score(dot, nq2, ne2);
}

using MinHeap = std::priority_queue<uint64_t, std::vector<uint64_t>, std::greater<uint64_t>>;
inline void heap_offer(MinHeap& h, uint32_t k, uint64_t key) {
  if (h.size() < k)
    h.push(key);
  else if (key > h.top()) {
    h.pop();
    h.push(key);
  }
}

void emit_topk(std::vector<uint64_t>& keys, uint32_t k, uint32_t* out_rows, float* out_scores,
               uint32_t* out_count) {
  std::sort(keys.begin(), keys.end(), std::greater<uint64_t>());
  if (keys.size() > k) keys.resize(k);
  for (uint32_t i = 0; i < k; ++i) {
    if (i < keys.size()) {
      out_rows[i] = key_row(keys[i]);
      out_scores[i] = key_score(keys[i]);
    } else {
      out_rows[i] = 0xFFFFFFFFu;
      out_scores[i] = 0.0f;
    }
  }
  *out_count = (uint32_t)keys.size();
}

int resolve_threads(int threads) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  return threads;
}

/* ------------------------------------------------------------------------ */
/* Token trie -- literal restatement of src/trie.rs:51-57,201-278            */
/* ------------------------------------------------------------------------ */
struct Node {
  std::unordered_map<std::string, Node> children; /* HashMap<String,TrieNode> :53 */
  bool is_end_of_word = false;                    /* :54 */
  std::vector<orc_docref> document_refs;          /* :55 */
  uint32_t frequency = 0;                         /* :56 */
};

void node_insert(Node* root, const std::vector<std::string>& tokens, const orc_docref& ref) {
  Node* cur = root; /* :212 */
  for (const auto& t : tokens) cur = &cur->children[t]; /* entry().or_insert_with :215 */
  cur->is_end_of_word = true;                           /* :218 */
  cur->document_refs.push_back(ref);                    /* :219 (no de-dup) */
  cur->frequency += 1;                                  /* :220 */
}

std::string join(const std::vector<std::string>& v) {
  std::string s;
  for (size_t i = 0; i < v.size(); ++i) {
    if (i) s.push_back(' ');
    s += v[i];
  }
  return s;
}

/* collect_completions (:257-278) without the limit; the caller sorts and
 * truncates.  Emits terminals strictly below the start node (:266). */
void collect_all(const Node* node, const std::vector<std::string>& prefix,
                 std::vector<std::string>* out) {
  std::vector<std::pair<const Node*, std::vector<std::string>>> stack;
  stack.emplace_back(node, prefix);
  while (!stack.empty()) {
    auto item = std::move(stack.back());
    stack.pop_back();
    const Node* cur = item.first;
    const auto& path = item.second;
    if (cur->is_end_of_word && path.size() > prefix.size()) out->push_back(join(path));
    for (const auto& kv : cur->children) {
      auto np = path;
      np.push_back(kv.first);
      stack.emplace_back(&kv.second, std::move(np));
    }
  }
}

void collect_postings(const Node* node, std::vector<orc_docref>* out) {
  std::vector<const Node*> stack{node};
  while (!stack.empty()) {
    const Node* cur = stack.back();
    stack.pop_back();
    if (cur->is_end_of_word) out->insert(out->end(), cur->document_refs.begin(), cur->document_refs.end());
    for (const auto& kv : cur->children) stack.push_back(&kv.second);
  }
}

const Node* walk(const Node* root, const std::vector<std::string>& tokens) {
  const Node* cur = root;
  for (const auto& t : tokens) { /* :227-238 */
    auto it = cur->children.find(t);
    if (it == cur->children.end()) return nullptr;
    cur = &it->second;
  }
  return cur;
}

/* split_whitespace (ASCII whitespace; the reference is Unicode-aware --
 * non-ASCII whitespace / case folding are out of the synthetic corpus). */
std::vector<std::string> split_ws(const char* s, bool lower) {
  std::vector<std::string> out;
  std::string cur;
  for (const char* p = s; *p; ++p) {
    unsigned char c = (unsigned char)*p;
    if (c == ' ' || (c >= 9 && c <= 13)) {
      if (!cur.empty()) out.push_back(cur), cur.clear();
    } else {
      cur.push_back(lower && c >= 'A' && c <= 'Z' ? (char)(c + 32) : (char)c);
    }
  }
  if (!cur.empty()) out.push_back(cur);
  return out;
}

} /* namespace */

struct orc_trie_index {
  Node tries[3]; /* case_name, content, citation: src/trie.rs:28-33 */
};

namespace {
std::vector<std::string> tokens_for(int which, const char* query) {
  /* case-name and content lower-case (:147,158,171,177); citation keeps case (:190,196) */
  return split_ws(query, which != ORC_TRIE_CITATION);
}

orc_trie_result* search_node(const Node* root, const std::vector<std::string>& tokens) {
  auto* r = (orc_trie_result*)calloc(1, sizeof(orc_trie_result));
  const Node* cur = walk(root, tokens);
  if (!cur) return r; /* miss: all empty, total 0 (:232-236) */
  if (cur->is_end_of_word && !cur->document_refs.empty()) { /* :241-245 */
    r->n_exact = cur->document_refs.size();
    r->exact_matches = (orc_docref*)malloc(sizeof(orc_docref) * r->n_exact);
    memcpy(r->exact_matches, cur->document_refs.data(), sizeof(orc_docref) * r->n_exact);
  }
  r->frequency = cur->frequency;
  std::vector<std::string> all;
  collect_all(cur, tokens, &all); /* :248 */
  std::sort(all.begin(), all.end());
  r->n_completions_unlimited = all.size();
  size_t lim = std::min<size_t>(all.size(), 10); /* limit 10, :248 */
  r->n_completions = lim;
  if (lim) {
    r->completions = (char**)malloc(sizeof(char*) * lim);
    for (size_t i = 0; i < lim; ++i) r->completions[i] = strdup(all[i].c_str());
  }
  r->total_matches = r->n_exact + r->n_completions; /* :251 */
  return r;
}
} /* namespace */

extern "C" {

void orc_gen_rows(float* out, uint64_t row_begin, uint64_t nrows, uint32_t dim, uint64_t seed) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < (int64_t)nrows; ++r)
    gen_row(out + (size_t)r * dim, row_begin + (uint64_t)r, dim, seed);
}

uint64_t orc_pack_key(float score, uint32_t row) { return pack_key(score, row); }

int orc_num_threads(void) { return resolve_threads(0); }

void orc_scores(const float* rows, uint64_t n, uint32_t dim, const float* query,
                float* out_scores, int order, int threads) {
  threads = resolve_threads(threads);
  QueryPrep qp = prep_query(query, dim);
#pragma omp parallel for schedule(static) num_threads(threads)
  for (int64_t r = 0; r < (int64_t)n; ++r) {
    const float* e = rows + (size_t)r * dim;
    if (order == ORC_ORDER_SEQUENTIAL) {
      out_scores[r] = seq_score(e, dim, query);
    } else {
      float dot, ne2;
      canon_dot_norm(e, dim, qp.qp.data(), &dot, &ne2, false);
      out_scores[r] = finish_score(dot, qp.nq2, ne2);
    }
  }
}

void orc_scores_bf16(const float* rows, uint64_t n, uint32_t dim, const float* query,
                     float* out_scores, int threads) {
  threads = resolve_threads(threads);
  QueryPrep qp = prep_query(query, dim);
#pragma omp parallel for schedule(static) num_threads(threads)
  for (int64_t r = 0; r < (int64_t)n; ++r) {
    float dot, ne2;
    canon_dot_norm(rows + (size_t)r * dim, dim, qp.qp.data(), &dot, &ne2, true);
    out_scores[r] = finish_score(dot, qp.nq2, ne2);
  }
}

int orc_cosine_topk(const float* rows, uint64_t n, uint32_t dim, const float* queries,
                    uint32_t nq, uint32_t k, const uint32_t* mask_words, int mask_mode,
                    uint64_t row_base, uint32_t* out_rows, float* out_scores,
                    uint32_t* out_counts, int order, int storage_bf16, int threads) {
  if (!rows && n) return 1;
  if (!queries || !k || !dim) return 1;
  threads = resolve_threads(threads);
  for (uint32_t qi = 0; qi < nq; ++qi) {
    const float* q = queries + (size_t)qi * dim;
    QueryPrep qp = prep_query(q, dim);
    std::vector<MinHeap> heaps(threads);
#pragma omp parallel num_threads(threads)
    {
#ifdef _OPENMP
      int tid = omp_get_thread_num();
#else
      int tid = 0;
#endif
      MinHeap& h = heaps[tid];
#pragma omp for schedule(static)
      for (int64_t r = 0; r < (int64_t)n; ++r) {
        if (mask_mode != ORC_MASK_NONE && mask_words) {
          uint32_t bit = (mask_words[r >> 5] >> (r & 31)) & 1u;
          if (mask_mode == ORC_MASK_INCLUDE ? !bit : bit) continue;
        }
        const float* e = rows + (size_t)r * dim;
        float s;
        if (order == ORC_ORDER_SEQUENTIAL) {
          s = seq_score(e, dim, q);
        } else {
          float dot, ne2;
          canon_dot_norm(e, dim, qp.qp.data(), &dot, &ne2, storage_bf16 != 0);
          s = finish_score(dot, qp.nq2, ne2);
        }
        heap_offer(h, k, pack_key(s, (uint32_t)(row_base + (uint64_t)r)));
      }
    }
    std::vector<uint64_t> keys;
    for (auto& h : heaps)
      while (!h.empty()) keys.push_back(h.top()), h.pop();
    emit_topk(keys, k, out_rows + (size_t)qi * k, out_scores + (size_t)qi * k, out_counts + qi);
  }
  return 0;
}

int orc_cosine_topk_synth(uint64_t row_begin, uint64_t nrows, uint32_t dim, uint64_t seed,
                          const float* queries, uint32_t nq, uint32_t k, uint32_t* out_rows,
                          float* out_scores, uint32_t* out_counts, int threads) {
  if (!queries || !k || !dim) return 1;
  threads = resolve_threads(threads);
  std::vector<QueryPrep> qps;
  for (uint32_t qi = 0; qi < nq; ++qi) qps.push_back(prep_query(queries + (size_t)qi * dim, dim));
  std::vector<std::vector<MinHeap>> heaps(threads, std::vector<MinHeap>(nq));
#pragma omp parallel num_threads(threads)
  {
#ifdef _OPENMP
    int tid = omp_get_thread_num();
#else
    int tid = 0;
#endif
    std::vector<float> e(dim);
#pragma omp for schedule(static)
    for (int64_t r = 0; r < (int64_t)nrows; ++r) {
      uint64_t row = row_begin + (uint64_t)r;
      gen_row(e.data(), row, dim, seed);
      for (uint32_t qi = 0; qi < nq; ++qi) {
        float dot, ne2;
        canon_dot_norm(e.data(), dim, qps[qi].qp.data(), &dot, &ne2, false);
        heap_offer(heaps[tid][qi], k, pack_key(finish_score(dot, qps[qi].nq2, ne2), (uint32_t)row));
      }
    }
  }
  for (uint32_t qi = 0; qi < nq; ++qi) {
    std::vector<uint64_t> keys;
    for (int t = 0; t < threads; ++t) {
      auto& h = heaps[t][qi];
      while (!h.empty()) keys.push_back(h.top()), h.pop();
    }
    emit_topk(keys, k, out_rows + (size_t)qi * k, out_scores + (size_t)qi * k, out_counts + qi);
  }
  return 0;
}

/* ---- trie ---------------------------------------------------------------- */

orc_trie_index* orc_trie_new(void) { return new orc_trie_index(); }
void orc_trie_free(orc_trie_index* t) { delete t; }

void orc_trie_insert_case_name(orc_trie_index* t, const char* name, const uint8_t case_id[16]) {
  orc_docref ref; /* DocRef{case_id, paragraph_index:0, char_offset:None} :148-152 */
  memcpy(ref.case_id, case_id, 16);
  ref.paragraph_index = 0;
  ref.char_offset = -1;
  node_insert(&t->tries[ORC_TRIE_CASE_NAME], split_ws(name, true), ref);
}

void orc_trie_insert_content(orc_trie_index* t, const char* const* tokens, uint32_t ntokens,
                             const orc_docref* ref) {
  std::vector<std::string> toks; /* each token lower-cased, NOT re-split :171 */
  for (uint32_t i = 0; i < ntokens; ++i) {
    std::string s(tokens[i]);
    for (auto& c : s)
      if (c >= 'A' && c <= 'Z') c = (char)(c + 32);
    toks.push_back(std::move(s));
  }
  node_insert(&t->tries[ORC_TRIE_CONTENT], toks, *ref);
}

void orc_trie_insert_citation(orc_trie_index* t, const char* citation, const orc_docref* ref) {
  node_insert(&t->tries[ORC_TRIE_CITATION], split_ws(citation, false), *ref);
}

orc_trie_result* orc_trie_search_one(const orc_trie_index* t, int which, const char* query) {
  return search_node(&t->tries[which], tokens_for(which, query));
}

orc_trie_result* orc_trie_search(const orc_trie_index* t, const char* query) {
  /* cascade, src/trie.rs:112-130 */
  orc_trie_result* r = orc_trie_search_one(t, ORC_TRIE_CASE_NAME, query);
  if (r->n_exact) return r; /* :114-118 */
  orc_trie_result_free(r);
  r = orc_trie_search_one(t, ORC_TRIE_CITATION, query);
  if (r->n_exact) return r; /* :121-125 */
  orc_trie_result_free(r);
  /* :128-129: split_whitespace WITHOUT lower-casing, then search_tokens lower-cases */
  return orc_trie_search_one(t, ORC_TRIE_CONTENT, query);
}

void orc_trie_result_free(orc_trie_result* r) {
  if (!r) return;
  free(r->exact_matches);
  for (uint64_t i = 0; i < r->n_completions; ++i) free(r->completions[i]);
  free(r->completions);
  free(r);
}

uint64_t orc_trie_prefix_postings(const orc_trie_index* t, int which, const char* query,
                                  orc_docref** out) {
  *out = nullptr;
  const Node* cur = walk(&t->tries[which], tokens_for(which, query));
  if (!cur) return 0;
  std::vector<orc_docref> refs;
  collect_postings(cur, &refs);
  if (refs.empty()) return 0;
  *out = (orc_docref*)malloc(sizeof(orc_docref) * refs.size());
  memcpy(*out, refs.data(), sizeof(orc_docref) * refs.size());
  return refs.size();
}

void orc_free(void* p) { free(p); }

/* ---- hybrid merge (src/search.rs:185-240) -------------------------------- */

uint32_t orc_hybrid_merge(const uint64_t* exact_cases, uint32_t n_exact, const uint64_t* vec_cases,
                          const float* vec_scores, uint32_t n_vec, int enable_prefix,
                          int enable_semantic, uint32_t cfg_max_results,
                          int64_t query_max_results, float min_similarity,
                          float exact_match_weight, orc_hit* out, uint32_t cap) {
  std::vector<orc_hit> all;
  std::unordered_set<uint64_t> seen; /* seen_cases :187 */
  if (enable_prefix) {               /* :190 */
    for (uint32_t i = 0; i < n_exact; ++i)
      if (seen.insert(exact_cases[i]).second) /* :194 */
        all.push_back({exact_cases[i], exact_match_weight, 0});
  }
  if (enable_semantic && all.size() < cfg_max_results) { /* :209 */
    for (uint32_t i = 0; i < n_vec; ++i) {
      if (vec_scores[i] >= min_similarity) {      /* :212 */
        if (seen.insert(vec_cases[i]).second)     /* :214 */
          all.push_back({vec_cases[i], vec_scores[i], 2});
      }
    }
  }
  /* stable sort, score desc, incomparable == Equal (:230) */
  std::stable_sort(all.begin(), all.end(),
                   [](const orc_hit& a, const orc_hit& b) { return a.score > b.score; });
  size_t maxr = query_max_results >= 0 ? (size_t)query_max_results : cfg_max_results; /* :236 */
  if (all.size() > maxr) all.resize(maxr);                                            /* :237 */
  uint32_t n = (uint32_t)std::min<size_t>(all.size(), cap);
  for (uint32_t i = 0; i < n; ++i) out[i] = all[i];
  return n;
}

} /* extern "C" */
