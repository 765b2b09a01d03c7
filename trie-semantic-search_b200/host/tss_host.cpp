// tss_host.cpp -- see tss_host.hpp.  Reference citations are into /root/reference/src.
#include "tss_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <climits>
#include <ctime>
#include <unordered_set>

#include "../../include/tss.h"

namespace tss_host {

namespace {
[[noreturn]] void raise_tss(SearchError::Kind kind, const char* what, int rc) {
  throw SearchError(kind, std::string(what) + ": libtss status " + std::to_string(rc) + ": " +
                              tss_last_error());
}
bool is_ws(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }
}  // namespace

// ---- core types -----------------------------------------------------------------------------
CaseId CaseId::from_u64(uint64_t v) {
  CaseId c;
  for (int i = 0; i < 8; ++i) c.bytes[15 - i] = (uint8_t)(v >> (8 * i));
  return c;
}
std::string CaseId::to_string() const {
  char buf[40];
  const uint8_t* b = bytes.data();
  snprintf(buf, sizeof(buf),
           "%02x%02x%02x%02x-%02x%02x-%02x%02x-%02x%02x-%02x%02x%02x%02x%02x%02x", b[0], b[1],
           b[2], b[3], b[4], b[5], b[6], b[7], b[8], b[9], b[10], b[11], b[12], b[13], b[14],
           b[15]);
  return buf;
}
size_t CaseIdHash::operator()(const CaseId& c) const {
  uint64_t a, b;
  memcpy(&a, c.bytes.data(), 8);
  memcpy(&b, c.bytes.data() + 8, 8);
  return (size_t)(a * 0x9E3779B97F4A7C15ull ^ (b + 0x7F4A7C15ull + (a << 6) + (a >> 2)));
}

const char* SearchError::category() const {  // src/errors.rs:236-272
  switch (kind) {
    case VectorIndexFailed:
    case HnswSearchError: return "vector";
    case InvalidSearchQuery: return "search";
    case NotSupported: return "system";
  }
  return "unknown";
}

// ---- HnswIndex ------------------------------------------------------------------------------
HnswIndex::HnswIndex(const HnswConfig& config, size_t dimension, int device, bool bf16)
    : config_(config), dim_(dimension), device_(device) {
  int rc = tss_index_create(&ix_, (uint32_t)dimension, bf16 ? TSS_BF16 : TSS_F32, device);
  if (rc) raise_tss(SearchError::VectorIndexFailed, "HnswIndex::new", rc);
}
HnswIndex::~HnswIndex() { tss_index_destroy(ix_); }

void HnswIndex::add_vector(const DocRef& doc_ref, const std::vector<float>& embedding) {
  if (embedding.size() != dim_)
    throw SearchError(SearchError::VectorIndexFailed,
                      "embedding has " + std::to_string(embedding.size()) + " dims, index has " +
                          std::to_string(dim_));
  for (float v : embedding)
    if (!std::isfinite(v))
      throw SearchError(SearchError::VectorIndexFailed, "embedding contains NaN or Inf");
  if (row_docref_.size() >= config_.max_elements)
    throw SearchError(SearchError::VectorIndexFailed, "max_elements reached");
  uint32_t row = (uint32_t)row_docref_.size();
  row_docref_.push_back(doc_ref);
  case_rows_[doc_ref.case_id].push_back(row);
  pending_.insert(pending_.end(), embedding.begin(), embedding.end());
  ++pending_rows_;
  dirty_ = true;
  if (pending_rows_ >= 16384) flush();
}

void HnswIndex::flush() {
  if (pending_rows_) {
    int rc = tss_index_add(ix_, pending_.data(), pending_rows_);
    pending_.clear();
    pending_rows_ = 0;
    if (rc) {
      // the device refused the batch (tss_index_add leaves the index unchanged): the row ->
      // DocRef table must not run ahead of it, or later hits would name the wrong documents
      const size_t have = (size_t)tss_index_size(ix_);
      while (row_docref_.size() > have) {
        auto it = case_rows_.find(row_docref_.back().case_id);
        if (it != case_rows_.end()) {
          it->second.pop_back();
          if (it->second.empty()) case_rows_.erase(it);
        }
        row_docref_.pop_back();
      }
      raise_tss(SearchError::VectorIndexFailed, "HnswIndex::add_vector", rc);
    }
  }
}

size_t HnswIndex::size() const { return row_docref_.size(); }

void HnswIndex::make_searchable() {
  if (!dirty_) return;
  flush();
  int rc = tss_index_finalize(ix_);
  if (rc) raise_tss(SearchError::VectorIndexFailed, "HnswIndex finalize", rc);
  dirty_ = false;
}

namespace {
struct DocRefRecord {
  uint8_t case_id[16];
  uint64_t paragraph_index;
  int64_t char_offset;  // -1 = None
};
}  // namespace

void HnswIndex::save(const std::string& path) {
  make_searchable();
  int rc = tss_index_save(ix_, (path + ".tssidx").c_str());
  if (rc) raise_tss(SearchError::VectorIndexFailed, "HnswIndex::save", rc);
  FILE* f = fopen((path + ".docrefs").c_str(), "wb");
  if (!f) throw SearchError(SearchError::VectorIndexFailed, "cannot write " + path + ".docrefs");
  uint64_t n = row_docref_.size();
  bool ok = fwrite(&n, sizeof(n), 1, f) == 1;
  for (const DocRef& d : row_docref_) {
    DocRefRecord r;
    memcpy(r.case_id, d.case_id.bytes.data(), 16);
    r.paragraph_index = d.paragraph_index;
    r.char_offset = d.char_offset ? (int64_t)*d.char_offset : -1;
    ok = ok && fwrite(&r, sizeof(r), 1, f) == 1;
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok) throw SearchError(SearchError::VectorIndexFailed, "writing " + path + ".docrefs failed");
}

std::unique_ptr<HnswIndex> HnswIndex::load(const HnswConfig& config, const std::string& path,
                                           int device) {
  std::unique_ptr<HnswIndex> h(new HnswIndex());
  h->config_ = config;
  h->device_ = device;
  int rc = tss_index_load(&h->ix_, (path + ".tssidx").c_str(), device);
  if (rc) raise_tss(SearchError::VectorIndexFailed, "HnswIndex::load", rc);
  h->dim_ = tss_index_dim(h->ix_);
  FILE* f = fopen((path + ".docrefs").c_str(), "rb");
  if (!f) throw SearchError(SearchError::VectorIndexFailed, "cannot read " + path + ".docrefs");
  uint64_t n = 0;
  bool ok = fread(&n, sizeof(n), 1, f) == 1 && n == tss_index_size(h->ix_);
  for (uint64_t i = 0; ok && i < n; ++i) {
    DocRefRecord r;
    ok = fread(&r, sizeof(r), 1, f) == 1;
    if (!ok) break;
    DocRef d;
    memcpy(d.case_id.bytes.data(), r.case_id, 16);
    d.paragraph_index = (size_t)r.paragraph_index;
    if (r.char_offset >= 0) d.char_offset = (size_t)r.char_offset;
    h->case_rows_[d.case_id].push_back((uint32_t)i);
    h->row_docref_.push_back(d);
  }
  fclose(f);
  if (!ok) throw SearchError(SearchError::VectorIndexFailed, path + ".docrefs does not match the index");
  h->dirty_ = false;
  return h;
}

const std::vector<uint32_t>* HnswIndex::rows_of_case(const CaseId& id) const {
  auto it = case_rows_.find(id);
  return it == case_rows_.end() ? nullptr : &it->second;
}

void HnswIndex::set_batch_policy(uint32_t min_queries, bool build_shadow_now) {
  if (build_shadow_now) make_searchable();
  int rc = tss_index_set_batch_policy(ix_, min_queries, build_shadow_now ? 1 : 0);
  if (rc) raise_tss(SearchError::VectorIndexFailed, "HnswIndex set_batch_policy", rc);
}

std::vector<std::vector<std::pair<DocRef, float>>> HnswIndex::search_batch(
    const std::vector<std::vector<float>>& queries, size_t top_k) {
  std::vector<std::vector<std::pair<DocRef, float>>> out(queries.size());
  if (queries.empty()) return out;
  std::vector<float> flat;
  for (const auto& q : queries) {
    if (q.size() != dim_)
      throw SearchError(SearchError::HnswSearchError, "query dimension mismatch");
    flat.insert(flat.end(), q.begin(), q.end());
  }
  if (dirty_) {
    flush();
    int rc = tss_index_finalize(ix_);
    if (rc) raise_tss(SearchError::VectorIndexFailed, "HnswIndex finalize", rc);
    dirty_ = false;
  }
  const uint32_t nq = (uint32_t)queries.size(), k = (uint32_t)top_k;
  std::vector<uint32_t> rows((size_t)nq * k), counts(nq);
  std::vector<float> scores((size_t)nq * k);
  int rc = tss_index_search(ix_, flat.data(), nq, k, nullptr, TSS_MASK_NONE, rows.data(),
                            scores.data(), counts.data());
  if (rc) raise_tss(SearchError::HnswSearchError, "HnswIndex::search", rc);
  for (uint32_t qi = 0; qi < nq; ++qi)
    for (uint32_t j = 0; j < counts[qi]; ++j)
      out[qi].emplace_back(row_docref_[rows[(size_t)qi * k + j]],
                           1.0f - scores[(size_t)qi * k + j]);
  return out;
}

std::vector<std::pair<DocRef, float>> HnswIndex::search_masked(
    const std::vector<float>& query_embedding, size_t top_k, const tss_mask* mask, int mask_mode) {
  if (query_embedding.size() != dim_)
    throw SearchError(SearchError::HnswSearchError,
                      "query has " + std::to_string(query_embedding.size()) + " dims, index has " +
                          std::to_string(dim_));
  if (top_k == 0) return {};
  if (dirty_) {
    flush();
    int rc = tss_index_finalize(ix_);
    if (rc) raise_tss(SearchError::VectorIndexFailed, "HnswIndex finalize", rc);
    dirty_ = false;
  }
  const uint32_t k = (uint32_t)top_k;
  std::vector<uint32_t> rows(k);
  std::vector<float> scores(k);
  uint32_t count = 0;
  int rc = tss_index_search(ix_, query_embedding.data(), 1, k, mask, mask_mode, rows.data(),
                            scores.data(), &count);
  if (rc) raise_tss(SearchError::HnswSearchError, "HnswIndex::search", rc);
  std::vector<std::pair<DocRef, float>> out;
  out.reserve(count);
  // the ABI returns similarity; the reference signature carries a distance that the caller
  // turns back with 1.0 - distance (src/vector.rs:144)
  for (uint32_t j = 0; j < count; ++j) out.emplace_back(row_docref_[rows[j]], 1.0f - scores[j]);
  return out;
}

std::vector<std::pair<DocRef, float>> HnswIndex::search(const std::vector<float>& query_embedding,
                                                        size_t top_k) {
  return search_masked(query_embedding, top_k, nullptr, TSS_MASK_NONE);
}

// ---- embedding + cache -----------------------------------------------------------------------
EmbeddingResult EmbeddingModel::encode(const std::string& text) const {
  EmbeddingResult r;
  r.embedding = encoder_ ? encoder_(text) : std::vector<float>(dim_, 0.0f);  // src/vector.rs:173
  return r;
}
std::optional<std::vector<float>> VectorCache::get(const std::string& key) const {
  auto it = cache_.find(key);
  if (it == cache_.end()) return std::nullopt;
  return it->second;
}
void VectorCache::insert(const std::string& key, std::vector<float> value) {
  if (cache_.size() >= max_size_ && !cache_.empty()) cache_.erase(cache_.begin());  // :223-228
  cache_[key] = std::move(value);
}

// ---- VectorIndex -----------------------------------------------------------------------------
VectorIndex::VectorIndex(const VectorConfig& config)
    : config_(config),
      embedding_model_(config.model, config.dimension),
      hnsw_index_(new HnswIndex(config.hnsw, config.dimension, config.device, config.bf16_storage)),
      vector_cache_(1000) {}  // src/vector.rs:72
VectorIndex::VectorIndex(const VectorConfig& config, std::unique_ptr<HnswIndex> loaded)
    : config_(config),
      embedding_model_(config.model, config.dimension),
      hnsw_index_(std::move(loaded)),
      vector_cache_(1000) {}

void VectorIndex::save_to_disk(const std::string& path) { hnsw_index_->save(path); }
std::unique_ptr<VectorIndex> VectorIndex::load_from_disk(const VectorConfig& config,
                                                         const std::string& path) {
  auto h = HnswIndex::load(config.hnsw, path, config.device);
  if (h->dimension() != config.dimension)
    throw SearchError(SearchError::VectorIndexFailed, "index file dimension differs from config");
  return std::unique_ptr<VectorIndex>(new VectorIndex(config, std::move(h)));
}

EmbeddingResult VectorIndex::generate_embedding(const std::string& text) {
  if (auto cached = vector_cache_.get(text)) return EmbeddingResult{*cached, 0};  // :100-105
  EmbeddingResult r = embedding_model_.encode(text);                              // :108
  vector_cache_.insert(text, r.embedding);                                        // :111
  return r;
}
void VectorIndex::add_document(const DocRef& doc_ref, const std::string& text) {
  hnsw_index_->add_vector(doc_ref, generate_embedding(text).embedding);  // :122-123
}
void VectorIndex::add_embedding(const DocRef& doc_ref, const std::vector<float>& embedding) {
  hnsw_index_->add_vector(doc_ref, embedding);
}
std::vector<VectorSearchResult> VectorIndex::search_masked(const std::string& query, size_t top_k,
                                                           const tss_mask* mask, int mask_mode) {
  EmbeddingResult q = generate_embedding(query);                                      // :134
  auto neighbors = hnsw_index_->search_masked(q.embedding, top_k, mask, mask_mode);   // :137
  std::vector<VectorSearchResult> out;
  out.reserve(neighbors.size());
  for (auto& n : neighbors)  // order preserved; similarity = 1.0 - distance  :140-147
    out.push_back(VectorSearchResult{n.first, 1.0f - n.second, std::nullopt});
  return out;
}
std::vector<VectorSearchResult> VectorIndex::search(const std::string& query, size_t top_k) {
  return search_masked(query, top_k, nullptr, TSS_MASK_NONE);
}
VectorIndexStats VectorIndex::get_stats() const {
  return VectorIndexStats{hnsw_index_->size(), vector_cache_.size(), config_.dimension};  // :153-159
}

// ---- TokenTrie ------------------------------------------------------------------------------
std::vector<std::string> TokenTrie::tokenize(const std::string& text) const {
  std::vector<std::string> out;
  std::string cur;
  for (unsigned char c : text) {
    if (is_ws(c)) {
      if (!cur.empty()) out.push_back(cur), cur.clear();
    } else {
      cur.push_back(lowercase_ && c >= 'A' && c <= 'Z' ? (char)(c + 32) : (char)c);
    }
  }
  if (!cur.empty()) out.push_back(cur);
  return out;
}
std::string TokenTrie::normalise_join(const std::vector<std::string>& tokens) const {
  std::string s;
  for (size_t i = 0; i < tokens.size(); ++i) {
    if (tokens[i].empty())
      throw SearchError(SearchError::NotSupported, "empty token in a trie key");
    if (i) s.push_back(' ');
    for (unsigned char c : tokens[i]) {
      if (is_ws(c))  // the flattened form uses ' ' as the token separator
        throw SearchError(SearchError::NotSupported, "token contains whitespace");
      s.push_back(lowercase_ && c >= 'A' && c <= 'Z' ? (char)(c + 32) : (char)c);
    }
  }
  return s;
}
void TokenTrie::insert_tokens(const std::vector<std::string>& tokens, const DocRef& ref) {
  terms_[normalise_join(tokens)].push_back(ref);  // push, no de-dup: src/trie.rs:219
}
uint32_t TokenTrie::frequency(const std::vector<std::string>& tokens) const {
  auto it = terms_.find(normalise_join(tokens));
  return it == terms_.end() ? 0 : (uint32_t)it->second.size();  // += 1 per insert, :220
}
TrieSearchResult TokenTrie::search_tokens(const std::vector<std::string>& tokens) const {
  TrieSearchResult r;
  const std::string p = normalise_join(tokens);
  // does the walk (src/trie.rs:227-238) reach a node?  A node exists iff some term equals P
  // or continues it with a further token.
  auto exact = terms_.find(p);
  const std::string sub = p.empty() ? std::string() : p + " ";
  auto it = p.empty() ? terms_.begin() : terms_.lower_bound(sub);
  auto below = [&](const std::string& key) {
    return p.empty() ? !key.empty() : key.compare(0, sub.size(), sub) == 0;
  };
  if (exact != terms_.end()) r.exact_matches = exact->second;  // :241-245
  for (; it != terms_.end() && r.prefix_completions.size() < 10; ++it) {  // limit 10, :248
    if (p.empty() && it->first.empty()) continue;  // the root itself is not "strictly longer"
    if (!below(it->first)) break;
    r.prefix_completions.push_back(it->first);  // path.join(" "), strictly longer :266-267
  }
  r.total_matches = r.exact_matches.size() + r.prefix_completions.size();  // :251
  return r;
}

// ---- TrieIndex --------------------------------------------------------------------------------
TrieIndex::TrieIndex(const TrieConfig& config)
    : config_(config),
      tries_{TokenTrie(true), TokenTrie(true), TokenTrie(false)} {}  // :147,171 lower; :190 as is
TrieIndex::~TrieIndex() {
  for (auto* t : frozen_) tss_terms_destroy(t);
}
void TrieIndex::insert_case_name(const std::string& case_name, const CaseId& case_id) {
  DocRef ref{case_id, 0, std::nullopt};  // :148-152
  tries_[CaseName].insert_tokens(tries_[CaseName].tokenize(case_name), ref);
}
void TrieIndex::insert_content(const std::vector<std::string>& tokens, const DocRef& ref) {
  tries_[Content].insert_tokens(tokens, ref);  // tokens lower-cased, not re-split :171
}
void TrieIndex::insert_citation(const std::string& citation, const DocRef& ref) {
  tries_[Citation].insert_tokens(tries_[Citation].tokenize(citation), ref);
}
TrieSearchResult TrieIndex::search_one(Which w, const std::string& query) const {
  return tries_[w].search_tokens(tries_[w].tokenize(query));
}
TrieSearchResult TrieIndex::search(const std::string& query) const {
  TrieSearchResult r = search_one(CaseName, query);  // :114-118
  if (!r.exact_matches.empty()) return r;
  r = search_one(Citation, query);  // :121-125
  if (!r.exact_matches.empty()) return r;
  return search_one(Content, query);  // :128-129
}
std::vector<std::string> TrieIndex::get_completions(const std::string&, size_t) const {
  return {};  // TODO in the reference too, :133-136
}
TrieIndex::TrieIndex(TrieIndex&& o) noexcept
    : config_(std::move(o.config_)),
      tries_{std::move(o.tries_[0]), std::move(o.tries_[1]), std::move(o.tries_[2])} {
  for (int w = 0; w < 3; ++w) frozen_[w] = o.frozen_[w], o.frozen_[w] = nullptr;
}
TrieIndex& TrieIndex::operator=(TrieIndex&& o) noexcept {
  if (this != &o) {
    for (auto* t : frozen_) tss_terms_destroy(t);
    config_ = std::move(o.config_);
    for (int w = 0; w < 3; ++w) {
      tries_[w] = std::move(o.tries_[w]);
      frozen_[w] = o.frozen_[w];
      o.frozen_[w] = nullptr;
    }
  }
  return *this;
}

namespace {
constexpr char kTrieMagic[8] = {'T', 'S', 'S', 'T', 'R', 'I', 'E', '1'};
struct FileCloser {
  void operator()(FILE* f) const {
    if (f) fclose(f);
  }
};
template <class T>
bool put(FILE* f, const T& v) { return fwrite(&v, sizeof(T), 1, f) == 1; }
template <class T>
bool get(FILE* f, T& v) { return fread(&v, sizeof(T), 1, f) == 1; }
std::string frozen_path(const std::string& path, int w) { return path + "." + std::to_string(w) + ".terms"; }
}  // namespace

void TrieIndex::save_to_disk(const std::string& path) const {
  std::unique_ptr<FILE, FileCloser> f(fopen(path.c_str(), "wb"));
  if (!f) throw SearchError(SearchError::NotSupported, "Saving trie to disk: cannot open " + path);
  bool ok = fwrite(kTrieMagic, 8, 1, f.get()) == 1;
  for (int w = 0; w < 3 && ok; ++w) {
    ok = put<uint8_t>(f.get(), frozen_[w] ? 1 : 0) && put<uint64_t>(f.get(), tries_[w].terms().size());
    for (const auto& kv : tries_[w].terms()) {
      ok = ok && put<uint32_t>(f.get(), (uint32_t)kv.first.size()) &&
           (kv.first.empty() || fwrite(kv.first.data(), kv.first.size(), 1, f.get()) == 1) &&
           put<uint64_t>(f.get(), kv.second.size());
      for (const DocRef& d : kv.second) {
        const int64_t off = d.char_offset ? (int64_t)*d.char_offset : -1;
        ok = ok && fwrite(d.case_id.bytes.data(), 16, 1, f.get()) == 1 &&
             put<uint64_t>(f.get(), d.paragraph_index) && put<int64_t>(f.get(), off);
      }
      if (!ok) break;
    }
  }
  if (!ok) throw SearchError(SearchError::NotSupported, "Saving trie to disk: write to " + path + " failed");
  for (int w = 0; w < 3; ++w) {
    if (!frozen_[w]) continue;
    int rc = tss_terms_save(frozen_[w], frozen_path(path, w).c_str());
    if (rc) raise_tss(SearchError::VectorIndexFailed, "TrieIndex::save_to_disk", rc);
  }
}

TrieIndex TrieIndex::load_from_disk(const std::string& path, int device) {
  std::unique_ptr<FILE, FileCloser> f(fopen(path.c_str(), "rb"));
  char magic[8];
  if (!f || fread(magic, 8, 1, f.get()) != 1 || memcmp(magic, kTrieMagic, 8) != 0)
    throw SearchError(SearchError::NotSupported, "Loading trie from disk: " + path + " is not a saved TrieIndex");
  TrieIndex t;
  bool had_frozen[3] = {false, false, false};
  for (int w = 0; w < 3; ++w) {
    uint8_t fr = 0;
    uint64_t nterms = 0;
    bool ok = get(f.get(), fr) && get(f.get(), nterms);
    had_frozen[w] = fr != 0;
    auto& terms = t.tries_[w].mutable_terms();
    for (uint64_t i = 0; ok && i < nterms; ++i) {
      uint32_t len = 0;
      uint64_t nposts = 0;
      ok = get(f.get(), len) && len < (1u << 24);
      std::string term(len, '\0');
      ok = ok && (len == 0 || fread(&term[0], len, 1, f.get()) == 1) && get(f.get(), nposts) &&
           nposts < (1ull << 40);
      std::vector<DocRef> posts;
      for (uint64_t j = 0; ok && j < nposts; ++j) {
        DocRef d;
        uint64_t para = 0;
        int64_t off = -1;
        ok = fread(d.case_id.bytes.data(), 16, 1, f.get()) == 1 && get(f.get(), para) && get(f.get(), off);
        d.paragraph_index = (size_t)para;
        if (off >= 0) d.char_offset = (size_t)off;
        posts.push_back(d);
      }
      if (ok) terms.emplace_hint(terms.end(), std::move(term), std::move(posts));
    }
    if (!ok) throw SearchError(SearchError::NotSupported, "Loading trie from disk: " + path + " is truncated");
  }
  for (int w = 0; w < 3; ++w) {
    if (!had_frozen[w]) continue;
    int rc = tss_terms_load(&t.frozen_[w], frozen_path(path, w).c_str(), device);
    if (rc) raise_tss(SearchError::VectorIndexFailed, "TrieIndex::load_from_disk", rc);
  }
  return t;
}

void TrieIndex::freeze(Which w, int device, const RowsOf& rows_of) {
  std::string pool;
  std::vector<uint64_t> term_off{0}, post_off{0};
  std::vector<uint32_t> post_rows;
  for (const auto& kv : tries_[w].terms()) {  // std::map iterates in byte order
    pool += kv.first;
    term_off.push_back(pool.size());
    for (const DocRef& ref : kv.second)
      if (const auto* rows = rows_of(ref.case_id))
        post_rows.insert(post_rows.end(), rows->begin(), rows->end());
    post_off.push_back(post_rows.size());
  }
  tss_terms_destroy(frozen_[w]);
  frozen_[w] = nullptr;
  int rc = tss_terms_create(&frozen_[w], pool.data(), term_off.data(), post_off.data(),
                            post_rows.data(), term_off.size() - 1, device);
  if (rc) raise_tss(SearchError::VectorIndexFailed, "TrieIndex::freeze", rc);
}
void TrieIndex::prefix_mask(Which w, const std::string& query, tss_mask* mask, uint64_t row_base,
                            bool fresh) const {
  if (!frozen_[w]) throw SearchError(SearchError::NotSupported, "TrieIndex::prefix_mask before freeze");
  std::string p;
  for (const auto& t : tries_[w].tokenize(query)) {
    if (!p.empty()) p.push_back(' ');
    p += t;
  }
  // fresh: the mask is zeroed by the same enqueue (no separate clear, no host synchronisation)
  int rc = (fresh ? tss_prefix_mask_fresh : tss_prefix_mask)(frozen_[w], p.data(), (uint32_t)p.size(),
                                                            TSS_PREFIX_TOKEN, mask, row_base, nullptr);
  if (rc) raise_tss(SearchError::HnswSearchError, "TrieIndex::prefix_mask", rc);
}

// ---- SearchEngine ----------------------------------------------------------------------------
std::optional<CaseMetadata> MetadataStore::get_case_metadata(const CaseId& id) const {
  ++gets_;
  auto it = map_.find(id);
  if (it == map_.end()) return std::nullopt;
  return it->second;
}
std::vector<std::optional<CaseMetadata>> MetadataStore::multi_get(const std::vector<CaseId>& ids) const {
  ++multi_gets_;
  std::vector<std::optional<CaseMetadata>> out;
  out.reserve(ids.size());
  for (const CaseId& id : ids) {
    auto it = map_.find(id);
    out.push_back(it == map_.end() ? std::nullopt : std::optional<CaseMetadata>(it->second));
  }
  return out;
}

SearchEngine::SearchEngine(const VectorConfig& vc, const TrieConfig& tc,
                           const SearchEngineConfig& sc, std::shared_ptr<MetadataStore> storage)
    : config_(sc), trie_index_(tc), vector_index_(vc), storage_(std::move(storage)) {}
SearchEngine::~SearchEngine() {
  tss_mask_destroy(mask_);
  tss_columns_destroy(columns_);
}

void SearchEngine::freeze() {
  HnswIndex& h = vector_index_.hnsw();
  auto rows_of = [&h](const CaseId& id) { return h.rows_of_case(id); };
  for (int w = 0; w < 3; ++w) trie_index_.freeze((TrieIndex::Which)w, h.device(), rows_of);
  tss_mask_destroy(mask_);
  mask_ = nullptr;
  mask_bits_ = h.size();
  int rc = tss_mask_create(&mask_, mask_bits_ ? mask_bits_ : 1, h.device());
  if (rc) raise_tss(SearchError::VectorIndexFailed, "SearchEngine::freeze", rc);
  // N3 columns: court id (dictionary, 65535 = unknown) and decision date of every row's case
  std::vector<uint16_t> court(h.size(), 65535);
  std::vector<int32_t> date(h.size(), 0);
  court_ids_.clear();
  for (uint32_t r = 0; r < h.size(); ++r) {
    auto meta = storage_->get_case_metadata(h.doc_ref_of_row(r).case_id);
    if (!meta) continue;
    auto it = court_ids_.find(meta->court);
    if (it == court_ids_.end()) {
      if (court_ids_.size() >= 65535)
        throw SearchError(SearchError::NotSupported, "more than 65535 distinct courts");
      it = court_ids_.emplace(meta->court, (uint16_t)court_ids_.size()).first;
    }
    court[r] = it->second;
    date[r] = meta->decision_date;
  }
  tss_columns_destroy(columns_);
  columns_ = nullptr;
  rc = tss_columns_create(&columns_, court.data(), date.data(), h.size(), h.device());
  if (rc) raise_tss(SearchError::VectorIndexFailed, "SearchEngine::freeze (columns)", rc);
}

std::vector<SearchResult> SearchEngine::search(const std::string& query) {
  SearchQuery q;
  q.query = query;
  q.max_results = config_.default_max_results;  // :152
  return search_with_params(q);
}

void SearchEngine::validate_query(const SearchQuery& query) const {
  if (query.query.size() < config_.min_query_length)  // :285-290
    throw SearchError(SearchError::InvalidSearchQuery,
                      "Invalid search query: " + query.query + " - Query too short: minimum " +
                          std::to_string(config_.min_query_length) + " characters");
  if (query.query.size() > config_.max_query_length)  // :292-297
    throw SearchError(SearchError::InvalidSearchQuery,
                      "Invalid search query: " + query.query + " - Query too long: maximum " +
                          std::to_string(config_.max_query_length) + " characters");
}

std::vector<SearchResult> SearchEngine::search_with_params(const SearchQuery& query) {
  if (config_.enable_query_cache) {  // key = query string only, :164-168,303-306
    auto it = query_cache_.find(query.query);
    if (it != query_cache_.end() &&
        (int64_t)time(nullptr) - it->second.timestamp < (int64_t)config_.query_cache_ttl_seconds)
      return it->second.results;
  }
  validate_query(query);                                           // :171
  std::vector<SearchResult> results = execute_hybrid_search(query);  // :174
  if (config_.enable_query_cache) {                                // :177-179,365-378
    if (query_cache_.size() >= config_.query_cache_size && !query_cache_.empty())
      query_cache_.erase(query_cache_.begin());
    query_cache_[query.query] = Cached{results, (int64_t)time(nullptr)};
  }
  return results;
}

std::vector<std::vector<SearchResult>> SearchEngine::search_batch(
    const std::vector<SearchQuery>& queries) {
  auto snippet = [](const DocRef& d) {
    return "Snippet for case " + d.case_id.to_string() + " paragraph " +
           std::to_string(d.paragraph_index);
  };
  const size_t n = queries.size();
  std::vector<std::vector<SearchResult>> all(n);
  std::vector<std::unordered_set<CaseId, CaseIdHash>> seen(n);
  std::vector<size_t> need;  // queries that run the semantic pass
  std::vector<std::vector<float>> embeddings;
  // trie pass of every query first, then ONE multi-get hydrates all their exact hits
  std::vector<std::vector<DocRef>> exact(n);
  std::vector<CaseId> ids;
  for (size_t i = 0; i < n; ++i) {
    const SearchQuery& q = queries[i];
    validate_query(q);
    if (q.config.enable_prefix) {  // :190-206
      exact[i] = trie_index_.search(q.query).exact_matches;
      for (const DocRef& d : exact[i]) ids.push_back(d.case_id);
    }
  }
  auto metas = storage_->multi_get(ids);
  size_t at = 0;
  for (size_t i = 0; i < n; ++i) {
    const SearchQuery& q = queries[i];
    for (const DocRef& d : exact[i]) {
      const auto& meta = metas[at++];
      if (meta && seen[i].insert(d.case_id).second)
        all[i].push_back(SearchResult{*meta, q.config.exact_match_weight, MatchType::Exact, snippet(d)});
    }
    if (q.config.enable_semantic && all[i].size() < q.config.max_results) {  // :209
      need.push_back(i);
      embeddings.push_back(vector_index_.generate_embedding(q.query).embedding);
    }
  }
  auto hits = vector_index_.hnsw().search_batch(embeddings, kVectorTopK);  // one device call
  // ... and ONE multi-get for every vector hit of the batch that passes its query's threshold
  ids.clear();
  for (size_t j = 0; j < need.size(); ++j)
    for (const auto& h : hits[j])
      if (1.0f - h.second >= queries[need[j]].config.min_similarity) ids.push_back(h.first.case_id);
  metas = storage_->multi_get(ids);
  at = 0;
  for (size_t j = 0; j < need.size(); ++j) {
    const size_t i = need[j];
    const SearchQuery& q = queries[i];
    for (const auto& h : hits[j]) {
      const float sim = 1.0f - h.second;  // :144
      if (sim < q.config.min_similarity) continue;  // :212
      const auto& meta = metas[at++];
      if (meta && seen[i].insert(h.first.case_id).second)
        all[i].push_back(SearchResult{*meta, sim, MatchType::Semantic, snippet(h.first)});
    }
  }
  for (size_t i = 0; i < n; ++i) {
    std::stable_sort(all[i].begin(), all[i].end(),
                     [](const SearchResult& a, const SearchResult& b) { return a.score > b.score; });
    all[i] = apply_filters(std::move(all[i]), queries[i]);
    size_t max_results = queries[i].max_results.value_or(queries[i].config.max_results);
    if (all[i].size() > max_results) all[i].resize(max_results);
  }
  return all;
}

std::vector<SearchResult> SearchEngine::apply_filters(std::vector<SearchResult> results,
                                                      const SearchQuery& query) const {
  if (query.court_filter) {  // :261-263
    const auto& cf = *query.court_filter;
    results.erase(std::remove_if(results.begin(), results.end(),
                                 [&](const SearchResult& r) {
                                   return std::find(cf.begin(), cf.end(), r.case_metadata.court) ==
                                          cf.end();
                                 }),
                  results.end());
  }
  if (query.date_range) {  // :266-271
    auto [lo, hi] = *query.date_range;
    results.erase(std::remove_if(results.begin(), results.end(),
                                 [&](const SearchResult& r) {
                                   return r.case_metadata.decision_date < lo ||
                                          r.case_metadata.decision_date > hi;
                                 }),
                  results.end());
  }
  return results;
}

std::vector<SearchResult> SearchEngine::execute_hybrid_search(const SearchQuery& query) {
  std::vector<SearchResult> all_results;
  std::unordered_set<CaseId, CaseIdHash> seen_cases;  // :187
  auto snippet = [](const DocRef& d) {              // :277-281
    return "Snippet for case " + d.case_id.to_string() + " paragraph " +
           std::to_string(d.paragraph_index);
  };

  // 1. trie search for exact matches, :190-206
  if (query.config.enable_prefix) {
    TrieSearchResult tr = trie_index_.search(query.query);
    std::vector<CaseId> ids;
    for (const DocRef& d : tr.exact_matches) ids.push_back(d.case_id);
    auto metas = storage_->multi_get(ids);  // N4: one lookup for all of them (:193 is per hit)
    size_t at = 0;
    for (const DocRef& d : tr.exact_matches) {
      const auto& meta = metas[at++];
      if (!meta) continue;
      if (seen_cases.insert(d.case_id).second)  // :194
        all_results.push_back(
            SearchResult{*meta, query.config.exact_match_weight, MatchType::Exact, snippet(d)});
    }
  }

  // 2. vector search for semantic matches, :209-227
  if (query.config.enable_semantic && all_results.size() < query.config.max_results) {
    HnswIndex& h = vector_index_.hnsw();
    const tss_mask* mask = nullptr;
    int mode = TSS_MASK_NONE;
    const bool filtered = prefilter_ && (query.court_filter || query.date_range);
    if (policy_ != MaskPolicy::PostHoc || filtered) {
      if (!mask_ || mask_bits_ != h.size())
        throw SearchError(SearchError::NotSupported, "SearchEngine::freeze() not called after the last insert");
      // every mask operation below is stream-ordered by the library (include/tss.h, masks):
      // clear -> prefix scatter -> filter -> row edits -> the masked search, no host round trips
      int rc = TSS_OK;
      if (policy_ != MaskPolicy::PrefixFilter) {  // (the prefix path clears inside its first enqueue)
        rc = tss_mask_clear(mask_);
        if (rc) raise_tss(SearchError::HnswSearchError, "mask clear", rc);
      }
      std::vector<uint32_t> seen_rows;
      if (policy_ == MaskPolicy::ExcludeOnDevice)
        for (const CaseId& c : seen_cases)
          if (const auto* r = h.rows_of_case(c)) seen_rows.insert(seen_rows.end(), r->begin(), r->end());
      bool have_include = false;
      if (policy_ == MaskPolicy::PrefixFilter) {  // rows at or below the node the query reaches
        for (int w = 0; w < 3; ++w)
          trie_index_.prefix_mask((TrieIndex::Which)w, query.query, mask_, 0, /*fresh=*/w == 0);
        have_include = true;
      }
      if (filtered) {  // N3: the filter's rows, intersected with the prefix rows if any
        std::vector<uint16_t> allowed;
        if (query.court_filter) {
          for (const auto& name : *query.court_filter) {
            auto it = court_ids_.find(name);
            if (it != court_ids_.end()) allowed.push_back(it->second);
          }
          if (allowed.empty()) allowed.push_back(65535);  // no known court: nothing passes
        }
        int32_t lo = INT32_MIN, hi = INT32_MAX;
        if (query.date_range) lo = query.date_range->first, hi = query.date_range->second;
        rc = tss_filter_mask(columns_, allowed.data(), (uint32_t)allowed.size(), lo, hi, mask_,
                             have_include ? 1 : 0);
        if (rc) raise_tss(SearchError::HnswSearchError, "filter mask", rc);
        have_include = true;
      }
      if (have_include) {
        if (!seen_rows.empty()) {  // seen cases leave the include set
          rc = tss_mask_clear_rows(mask_, seen_rows.data(), seen_rows.size(), 0);
          if (rc) raise_tss(SearchError::HnswSearchError, "mask clear_rows", rc);
        }
        mask = mask_;
        mode = TSS_MASK_INCLUDE;
      } else if (!seen_rows.empty()) {  // ExcludeOnDevice alone
        rc = tss_mask_set_rows(mask_, seen_rows.data(), seen_rows.size(), 0);
        if (rc) raise_tss(SearchError::HnswSearchError, "mask set_rows", rc);
        mask = mask_;
        mode = TSS_MASK_EXCLUDE;
      }
    }
    auto vector_results = vector_index_.search_masked(query.query, kVectorTopK, mask, mode);  // :251
    std::vector<CaseId> ids;
    for (const VectorSearchResult& v : vector_results)
      if (v.similarity_score >= query.config.min_similarity) ids.push_back(v.doc_ref.case_id);
    auto metas = storage_->multi_get(ids);  // N4: one lookup for the <= 50 hits (:213 is per hit)
    size_t at = 0;
    for (const VectorSearchResult& v : vector_results) {
      if (v.similarity_score >= query.config.min_similarity) {       // :212
        const auto& meta = metas[at++];
        if (!meta) continue;
        if (seen_cases.insert(v.doc_ref.case_id).second)             // :214
          all_results.push_back(
              SearchResult{*meta, v.similarity_score, MatchType::Semantic, snippet(v.doc_ref)});
      }
    }
  }

  // 3. stable sort by score desc (partial_cmp, Equal on incomparable), :230
  std::stable_sort(all_results.begin(), all_results.end(),
                   [](const SearchResult& a, const SearchResult& b) { return a.score > b.score; });
  all_results = apply_filters(std::move(all_results), query);  // :233
  size_t max_results = query.max_results.value_or(query.config.max_results);  // :236
  if (all_results.size() > max_results) all_results.resize(max_results);      // :237
  return all_results;
}

// ---- QueryBatcher ------------------------------------------------------------------------------
QueryBatcher::QueryBatcher(SearchEngine& engine, size_t max_batch, uint64_t max_wait_us)
    : engine_(engine), max_batch_(max_batch ? max_batch : 1), max_wait_us_(max_wait_us),
      worker_([this] { run(); }) {}

QueryBatcher::~QueryBatcher() {
  {
    std::lock_guard<std::mutex> l(mu_);
    stop_ = true;
  }
  cv_.notify_all();
  worker_.join();
}

std::future<std::vector<SearchResult>> QueryBatcher::submit(SearchQuery query) {
  std::promise<std::vector<SearchResult>> p;
  auto f = p.get_future();
  {
    std::lock_guard<std::mutex> l(mu_);
    queue_.emplace_back(std::move(query), std::move(p));
  }
  cv_.notify_all();
  return f;
}

void QueryBatcher::run() {
  for (;;) {
    std::vector<SearchQuery> batch;
    std::vector<std::promise<std::vector<SearchResult>>> promises;
    {
      std::unique_lock<std::mutex> l(mu_);
      cv_.wait(l, [&] { return stop_ || !queue_.empty(); });
      if (queue_.empty()) return;  // stop_ and nothing left to answer
      if (queue_.size() < max_batch_ && max_wait_us_)  // give the batch a moment to fill
        cv_.wait_for(l, std::chrono::microseconds(max_wait_us_),
                     [&] { return stop_ || queue_.size() >= max_batch_; });
      while (!queue_.empty() && batch.size() < max_batch_) {
        batch.push_back(std::move(queue_.front().first));
        promises.push_back(std::move(queue_.front().second));
        queue_.pop_front();
      }
    }
    try {
      auto results = engine_.search_batch(batch);
      for (size_t i = 0; i < promises.size(); ++i) promises[i].set_value(std::move(results[i]));
    } catch (...) {
      // one bad query (e.g. too short) must not fail its batch mates: answer one by one
      for (size_t i = 0; i < promises.size(); ++i) {
        try {
          promises[i].set_value(engine_.search_batch({batch[i]})[0]);
        } catch (...) {
          promises[i].set_exception(std::current_exception());
        }
      }
    }
    ++batches_;
    queries_ += batch.size();
  }
}

}  // namespace tss_host
