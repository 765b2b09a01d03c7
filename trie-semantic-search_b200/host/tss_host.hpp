// tss_host.hpp -- C++ host side above the C ABI, mirroring the reference's
// vector.rs / trie.rs / search.rs interface for the hot path (same type and
// method names, argument meaning and error behaviour).
//
// The reference is Rust and this image has no Rust toolchain,
// so the host shim is C++17; the Rust binding a maintainer would add is shipped
// as source in ../ffi/tss.rs and shown in INTEGRATION.md.
//
// Everything that scores, selects or masks goes through libtss.so (include/tss.h);
// nothing here computes a similarity on the CPU.
#pragma once

#include <array>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <future>
#include <mutex>
#include <thread>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

extern "C" {
struct tss_index;
struct tss_mask;
struct tss_terms;
struct tss_columns;
}

namespace tss_host {

// ---- core types (reference src/lib.rs:65-76,122-145) ------------------------------------
struct CaseId {  // Uuid
  std::array<uint8_t, 16> bytes{};
  bool operator==(const CaseId& o) const { return bytes == o.bytes; }
  bool operator<(const CaseId& o) const { return bytes < o.bytes; }
  static CaseId from_u64(uint64_t v);
  std::string to_string() const;
};
struct CaseIdHash {
  size_t operator()(const CaseId& c) const;
};

struct DocRef {  // src/lib.rs:68-76
  CaseId case_id;
  size_t paragraph_index = 0;
  std::optional<size_t> char_offset;
  bool operator==(const DocRef& o) const {
    return case_id == o.case_id && paragraph_index == o.paragraph_index &&
           char_offset == o.char_offset;
  }
};

struct SearchConfig {  // src/lib.rs:122-145
  size_t max_results = 10;
  float min_similarity = 0.5f;
  float exact_match_weight = 2.0f;
  bool enable_semantic = true;
  bool enable_prefix = true;
};

struct CaseMetadata {  // the fields the hot path reads, src/lib.rs:88-118
  CaseId id;
  std::string name, citation, court;
  int32_t decision_date = 0;  // days since 1970-01-01 (NaiveDate)
};

// ---- errors (reference src/errors.rs:149-153,184-185,...) --------------------------------
struct SearchError : std::runtime_error {
  enum Kind { VectorIndexFailed, HnswSearchError, InvalidSearchQuery, NotSupported };
  Kind kind;
  SearchError(Kind k, const std::string& msg) : std::runtime_error(msg), kind(k) {}
  const char* category() const;  // src/errors.rs:236-272
};

// ---- configs (field names as in src/config.rs:180-240,272-287) ---------------------------
struct HnswConfig {  // m / ef_* are accepted and ignored: the index is exact
  size_t m = 16, ef_construction = 200, ef_search = 50, max_elements = 10'000'000;
  std::string index_path = "./data/vector_index";
};
struct EmbeddingModelConfig {
  std::string model_path = "./models/legal-bert.onnx", model_type = "legal-bert";
};
struct VectorConfig {
  EmbeddingModelConfig model;
  HnswConfig hnsw;
  size_t dimension = 768;  // src/config.rs:571 (README says 384; a runtime parameter here)
  float similarity_threshold = 0.5f;
  size_t max_ann_results = 100;
  int device = 0;          // new: which B200
  bool bf16_storage = false;  // new: TSS_BF16 rows
};
struct TrieConfig {
  bool use_fst = true, index_case_names = true, index_citations = true;
  size_t max_prefix_length = 50;
  std::string index_path = "./data/trie_index";
};
struct SearchEngineConfig {  // src/config.rs:272-287,587-595
  size_t default_max_results = 10;
  bool enable_query_cache = true;
  size_t query_cache_size = 10000;
  uint64_t query_cache_ttl_seconds = 3600;
  size_t min_query_length = 2, max_query_length = 1000;
};

// ---- HnswIndex: the four-method seam (src/vector.rs:40-44,184-208) ------------------------
class HnswIndex {
 public:
  HnswIndex(const HnswConfig& config, size_t dimension, int device, bool bf16);  // ::new :185
  ~HnswIndex();
  HnswIndex(const HnswIndex&) = delete;
  HnswIndex& operator=(const HnswIndex&) = delete;

  void add_vector(const DocRef& doc_ref, const std::vector<float>& embedding);  // :190-193
  // (DocRef, distance) best first; distance = 1 - cosine similarity   :195-202
  std::vector<std::pair<DocRef, float>> search(const std::vector<float>& query_embedding,
                                               size_t top_k);
  size_t size() const;  // :204-207

  // --- beyond the reference signature ---
  // same, with the rows of `exclude_cases` / only the rows in `include_mask` considered
  std::vector<std::pair<DocRef, float>> search_masked(const std::vector<float>& query_embedding,
                                                      size_t top_k, const tss_mask* mask,
                                                      int mask_mode);
  std::vector<std::vector<std::pair<DocRef, float>>> search_batch(
      const std::vector<std::vector<float>>& queries, size_t top_k);
  // from how many queries a call leaves the plain scan for the tensor-core path / the shadow
  // prefilter (results are bit-identical either way; tss_index_set_batch_policy in tss.h)
  void set_batch_policy(uint32_t min_queries, bool build_shadow_now);
  // on-disk form: <path>.tssidx (tss_index_save) + <path>.docrefs
  void save(const std::string& path);
  static std::unique_ptr<HnswIndex> load(const HnswConfig& config, const std::string& path,
                                         int device);
  const std::vector<uint32_t>* rows_of_case(const CaseId& id) const;
  const DocRef& doc_ref_of_row(uint32_t row) const { return row_docref_[row]; }
  tss_index* handle() { return ix_; }
  int device() const { return device_; }
  size_t dimension() const { return dim_; }

 private:
  HnswIndex() = default;
  void flush();
  void make_searchable();
  HnswConfig config_;
  size_t dim_ = 0;
  int device_ = 0;
  tss_index* ix_ = nullptr;
  std::vector<DocRef> row_docref_;  // row -> DocRef (stays on the host)
  std::unordered_map<CaseId, std::vector<uint32_t>, CaseIdHash> case_rows_;
  std::vector<float> pending_;  // rows staged on the host until the next search
  size_t pending_rows_ = 0;
  bool dirty_ = true;
};

// ---- embedding + cache (src/vector.rs:34-38,46-50,162-182,210-235) -------------------------
struct EmbeddingResult {
  std::vector<float> embedding;
  uint64_t processing_time_ms = 0;
};
class EmbeddingModel {
 public:
  using Encoder = std::function<std::vector<float>(const std::string&)>;
  EmbeddingModel(const EmbeddingModelConfig& c, size_t dimension) : config_(c), dim_(dimension) {}
  // default = the reference's stub: vec![0.0; dimension] (src/vector.rs:173, with the
  // configured dimension instead of the literal 768)
  EmbeddingResult encode(const std::string& text) const;
  void set_encoder(Encoder e) { encoder_ = std::move(e); }

 private:
  EmbeddingModelConfig config_;
  size_t dim_;
  Encoder encoder_;
};
class VectorCache {  // HashMap + evict-one-when-full, src/vector.rs:210-235
 public:
  explicit VectorCache(size_t max_size) : max_size_(max_size) {}
  std::optional<std::vector<float>> get(const std::string& key) const;
  void insert(const std::string& key, std::vector<float> value);
  size_t size() const { return cache_.size(); }

 private:
  std::unordered_map<std::string, std::vector<float>> cache_;
  size_t max_size_;
};

struct VectorSearchResult {  // src/vector.rs:53-58
  DocRef doc_ref;
  float similarity_score = 0.f;
  std::optional<std::vector<float>> embedding;
};
struct VectorIndexStats {  // src/vector.rs:238-243
  size_t total_vectors = 0, cache_size = 0, dimension = 0;
};

class VectorIndex {  // src/vector.rs:27-160
 public:
  explicit VectorIndex(const VectorConfig& config);  // ::new :69-80
  EmbeddingResult generate_embedding(const std::string& text);          // :98-114
  void add_document(const DocRef& doc_ref, const std::string& text);   // :117-125
  // beyond the reference: the embedding is supplied (no model in this build)
  void add_embedding(const DocRef& doc_ref, const std::vector<float>& embedding);
  std::vector<VectorSearchResult> search(const std::string& query, size_t top_k);  // :128-150
  std::vector<VectorSearchResult> search_masked(const std::string& query, size_t top_k,
                                                const tss_mask* mask, int mask_mode);
  VectorIndexStats get_stats() const;  // :153-159
  // src/vector.rs:83-95 are TODO stubs (save = Ok(()), load = Self::new); here they persist
  void save_to_disk(const std::string& path);
  static std::unique_ptr<VectorIndex> load_from_disk(const VectorConfig& config,
                                                     const std::string& path);
  EmbeddingModel& embedding_model() { return embedding_model_; }
  HnswIndex& hnsw() { return *hnsw_index_; }

 private:
  VectorIndex(const VectorConfig& config, std::unique_ptr<HnswIndex> loaded);
  VectorConfig config_;
  EmbeddingModel embedding_model_;
  std::unique_ptr<HnswIndex> hnsw_index_;
  VectorCache vector_cache_;
};

// ---- TrieIndex (src/trie.rs:27-278) ---------------------------------------------------------
// The reference walks HashMap<String, TrieNode> nodes.  Here each trie is the
// flattened form the GPU consumes: unique terms (tokens joined by ' ') in byte
// order with their postings -- an ordered map while building, exported to
// tss_terms by freeze().  Semantics follow trie.rs line by line; the one
// visible difference is that prefix_completions come back byte-sorted (the
// reference's order is HashMap iteration order, i.e. unspecified).
struct TrieSearchResult {  // src/trie.rs:60-65
  std::vector<DocRef> exact_matches;
  std::vector<std::string> prefix_completions;
  size_t total_matches = 0;
};

class TokenTrie {
 public:
  explicit TokenTrie(bool lowercase) : lowercase_(lowercase) {}
  void insert_tokens(const std::vector<std::string>& tokens, const DocRef& ref);  // :211-221
  TrieSearchResult search_tokens(const std::vector<std::string>& tokens) const;   // :223-255
  std::vector<std::string> tokenize(const std::string& text) const;  // split_whitespace (+lower)
  size_t num_terms() const { return terms_.size(); }
  uint32_t frequency(const std::vector<std::string>& tokens) const;  // :56,220
  const std::map<std::string, std::vector<DocRef>>& terms() const { return terms_; }
  std::map<std::string, std::vector<DocRef>>& mutable_terms() { return terms_; }  // load_from_disk

 private:
  std::string normalise_join(const std::vector<std::string>& tokens) const;
  bool lowercase_;
  std::map<std::string, std::vector<DocRef>> terms_;
};

class TrieIndex {
 public:
  enum Which { CaseName = 0, Content = 1, Citation = 2 };
  explicit TrieIndex(const TrieConfig& config = TrieConfig());
  ~TrieIndex();
  TrieIndex(TrieIndex&& o) noexcept;
  TrieIndex& operator=(TrieIndex&& o) noexcept;
  TrieIndex(const TrieIndex&) = delete;
  TrieIndex& operator=(const TrieIndex&) = delete;
  void insert_case_name(const std::string& case_name, const CaseId& case_id);      // :97-99,146-155
  void insert_content(const std::vector<std::string>& tokens, const DocRef& ref);  // :102-104
  void insert_citation(const std::string& citation, const DocRef& ref);            // :107-109
  TrieSearchResult search(const std::string& query) const;  // cascade :112-130
  TrieSearchResult search_one(Which w, const std::string& query) const;
  std::vector<std::string> get_completions(const std::string& prefix, size_t limit) const;  // :133-136
  // The reference's pair is a NotSupported stub and a no-op (:83-94).  Here `path` holds the
  // three tries (terms + DocRef postings) and, for every frozen trie, `path.<w>.terms` holds its
  // flattened device form (tss_terms_save); loading restores both without a re-freeze.  A
  // missing or foreign file raises NotSupported, the error the stub always raised.
  static TrieIndex load_from_disk(const std::string& path, int device = 0);
  void save_to_disk(const std::string& path) const;

  // --- device side (K4) ---
  // export trie `w` as a flattened term array on `device`; postings become the rows
  // `rows_of(doc_ref.case_id)` returns (de-dup in the merge is per case, src/search.rs:194,214).
  using RowsOf = std::function<const std::vector<uint32_t>*(const CaseId&)>;
  void freeze(Which w, int device, const RowsOf& rows_of);
  // OR the rows of every posting at or below the node `query` reaches into `mask`.
  void prefix_mask(Which w, const std::string& query, tss_mask* mask, uint64_t row_base = 0,
                   bool fresh = false) const;
  const TokenTrie& trie(Which w) const { return tries_[w]; }

 private:
  TrieConfig config_;
  TokenTrie tries_[3];
  tss_terms* frozen_[3] = {nullptr, nullptr, nullptr};
};

// ---- SearchEngine (src/search.rs:30-342) -----------------------------------------------------
enum class MatchType { Exact, Prefix, Semantic, CaseName, Citation };  // :71-82

struct SearchQuery {  // :40-52
  std::string query;
  std::optional<size_t> max_results;
  std::optional<std::vector<std::string>> court_filter;
  std::optional<std::pair<int32_t, int32_t>> date_range;  // inclusive, days since epoch
  SearchConfig config;
};
struct SearchResult {  // :55-67
  CaseMetadata case_metadata;
  float score = 0.f;
  MatchType match_type = MatchType::Semantic;
  std::string snippet;
};

// what execute_hybrid_search needs from StorageManager (src/storage.rs:118-132); the
// sled-backed store itself is out of scope.
class MetadataStore {
 public:
  void put(const CaseMetadata& m) { map_[m.id] = m; }
  std::optional<CaseMetadata> get_case_metadata(const CaseId& id) const;
  // N4: one call for all the cases a query (or a whole batch of queries) has to hydrate, in the
  // order asked, nullopt where the reference's `if let Ok(Some(..))` would skip (:193,213).  The
  // reference does one sled lookup + bincode decode per hit (src/storage.rs:118-132), up to ~60
  // per query, which dominates once scoring takes microseconds.
  std::vector<std::optional<CaseMetadata>> multi_get(const std::vector<CaseId>& ids) const;
  // calls served so far (tests: the merge issues two multi-gets per query or batch, no gets)
  size_t gets() const { return gets_; }
  size_t multi_gets() const { return multi_gets_; }

 private:
  std::unordered_map<CaseId, CaseMetadata, CaseIdHash> map_;
  mutable std::atomic<size_t> gets_{0}, multi_gets_{0};
};

class SearchEngine {
 public:
  // how the semantic pass uses the trie result
  enum class MaskPolicy {
    PostHoc,        // reference-faithful: top-50 then skip seen cases on the host (:211-226)
    ExcludeOnDevice,  // seen cases become a TSS_MASK_EXCLUDE mask: the 50 are never wasted
    PrefixFilter    // BASELINE config 4: only rows in the prefix's posting set are scored
  };
  SearchEngine(const VectorConfig& vc, const TrieConfig& tc, const SearchEngineConfig& sc,
               std::shared_ptr<MetadataStore> storage);
  ~SearchEngine();
  std::vector<SearchResult> search(const std::string& query);                 // :149-159
  std::vector<SearchResult> search_with_params(const SearchQuery& query);     // :162-182
  // N4: many queries at once.  The trie pass and the merge run per query exactly as in
  // search_with_params (PostHoc policy, no query cache); the semantic pass of all queries that
  // need one is a single batched tss_index_search, so the corpus is streamed once per 4
  // queries (K1) or once per batch (K2) instead of once per query behind the write lock
  // (src/search.rs:249-252).
  std::vector<std::vector<SearchResult>> search_batch(const std::vector<SearchQuery>& queries);
  TrieIndex& trie_index() { return trie_index_; }
  VectorIndex& vector_index() { return vector_index_; }
  void set_mask_policy(MaskPolicy p) { policy_ = p; }
  // N3: apply court_filter / date_range on the device BEFORE top-k (an include mask built from
  // per-row court-id / date columns) instead of after it; off = reference behaviour (:233)
  void set_prefilter(bool on) { prefilter_ = on; }
  // call after the last insert: exports the tries to the device
  void freeze();
  static constexpr size_t kVectorTopK = 50;  // hard-coded in search_vector, :251

 private:
  std::vector<SearchResult> execute_hybrid_search(const SearchQuery& query);  // :185-240
  void validate_query(const SearchQuery& query) const;                          // :284-300
  std::vector<SearchResult> apply_filters(std::vector<SearchResult> r,
                                          const SearchQuery& q) const;        // :255-274
  SearchEngineConfig config_;
  TrieIndex trie_index_;
  VectorIndex vector_index_;
  std::shared_ptr<MetadataStore> storage_;
  MaskPolicy policy_ = MaskPolicy::PostHoc;
  bool prefilter_ = false;
  tss_columns* columns_ = nullptr;
  std::unordered_map<std::string, uint16_t> court_ids_;
  tss_mask* mask_ = nullptr;
  uint64_t mask_bits_ = 0;
  struct Cached {
    std::vector<SearchResult> results;
    int64_t timestamp;
  };
  std::unordered_map<std::string, Cached> query_cache_;  // QueryCache :104-116,344-385
};

// ---- N4: micro-batcher in front of the engine ---------------------------------------------------
// The reference funnels every vector search through one tokio write lock
// (src/search.rs:249-252): concurrent requests queue and each streams the corpus on its own.
// QueryBatcher owns the engine on one worker thread; callers submit() from any thread and get
// a future; the worker drains up to max_batch queued queries (waiting at most max_wait_us for
// the batch to fill) and answers them with ONE SearchEngine::search_batch call.
class QueryBatcher {
 public:
  QueryBatcher(SearchEngine& engine, size_t max_batch = 64, uint64_t max_wait_us = 200);
  ~QueryBatcher();
  QueryBatcher(const QueryBatcher&) = delete;
  QueryBatcher& operator=(const QueryBatcher&) = delete;
  std::future<std::vector<SearchResult>> submit(SearchQuery query);
  size_t batches_run() const { return batches_; }
  size_t queries_run() const { return queries_; }

 private:
  void run();
  SearchEngine& engine_;
  size_t max_batch_;
  uint64_t max_wait_us_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::pair<SearchQuery, std::promise<std::vector<SearchResult>>>> queue_;
  bool stop_ = false;
  size_t batches_ = 0, queries_ = 0;
  std::thread worker_;
};

}  // namespace tss_host
