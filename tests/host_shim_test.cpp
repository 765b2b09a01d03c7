// host_shim_test.cpp -- the C++ host shim (VectorIndex / TrieIndex / SearchEngine mirrors of
// reference src/vector.rs, src/trie.rs, src/search.rs) against hand-derived KATs and the
// CPU oracle.  `host_shim_test cpu` needs no GPU; `host_shim_test gpu` runs the device paths.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <unistd.h>
#include <unordered_map>
#include <unordered_set>

#include "../include/tss.h"
#include "../oracle/oracle.h"
#include "../trie-semantic-search_b200/host/tss_host.hpp"

using namespace tss_host;

static int g_fail = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);               \
      ++g_fail;                                                            \
    }                                                                      \
  } while (0)

static CaseId cid(uint64_t v) { return CaseId::from_u64(v); }
static orc_docref to_orc(const DocRef& d) {
  orc_docref o;
  memcpy(o.case_id, d.case_id.bytes.data(), 16);
  o.paragraph_index = d.paragraph_index;
  o.char_offset = d.char_offset ? (int64_t)*d.char_offset : -1;
  return o;
}
static bool same_ref(const DocRef& d, const orc_docref& o) {
  orc_docref mine = to_orc(d);
  return memcmp(mine.case_id, o.case_id, 16) == 0 && mine.paragraph_index == o.paragraph_index &&
         mine.char_offset == o.char_offset;
}

static void trie_kats() {
  TrieIndex t;
  const char* names[] = {"Brown v. Board of Education", "Miranda v. Arizona", "Roe v. Wade"};
  const char* cits[] = {"347 U.S. 483 (1954)", "384 U.S. 436 (1966)", "410 U.S. 113 (1973)"};
  for (int i = 0; i < 3; ++i) {
    t.insert_case_name(names[i], cid(i + 1));
    t.insert_citation(cits[i], DocRef{cid(i + 1), 0, std::nullopt});
  }
  auto r = t.search_one(TrieIndex::CaseName, "Brown v. Board of Education");  // K1
  CHECK(r.exact_matches.size() == 1 && r.exact_matches[0] == (DocRef{cid(1), 0, std::nullopt}));
  CHECK(r.prefix_completions.empty() && r.total_matches == 1);
  r = t.search_one(TrieIndex::CaseName, "BROWN V.");  // K2
  CHECK(r.exact_matches.empty() && r.prefix_completions.size() == 1 &&
        r.prefix_completions[0] == "brown v. board of education" && r.total_matches == 1);
  r = t.search_one(TrieIndex::CaseName, "bro");  // K3
  CHECK(r.total_matches == 0 && r.exact_matches.empty() && r.prefix_completions.empty());
  CHECK(t.search_one(TrieIndex::CaseName, "brown board").total_matches == 0);  // K4
  r = t.search_one(TrieIndex::Citation, "347 U.S.");                           // K5
  CHECK(r.prefix_completions.size() == 1 && r.prefix_completions[0] == "347 U.S. 483 (1954)");
  CHECK(t.search_one(TrieIndex::Citation, "347 u.s.").total_matches == 0);
  {
    TrieIndex d;  // K6
    d.insert_case_name("Same Name", cid(10));
    d.insert_case_name("same name", cid(11));
    auto e = d.search_one(TrieIndex::CaseName, "SAME NAME").exact_matches;
    CHECK(e.size() == 2 && e[0].case_id == cid(10) && e[1].case_id == cid(11));
    CHECK(d.trie(TrieIndex::CaseName).frequency({"same", "name"}) == 2);
  }
  r = t.search("brown v.");  // K7: cascade loses the case-name completion
  CHECK(r.total_matches == 0);
  CHECK(t.search("roe v. wade").exact_matches.size() == 1);
  CHECK(t.search("384 U.S. 436 (1966)").exact_matches[0].case_id == cid(2));
  r = t.search_one(TrieIndex::CaseName, "");  // K8
  CHECK(r.exact_matches.empty() && r.prefix_completions.size() == 3);
  {
    TrieIndex many;
    for (int i = 0; i < 25; ++i) many.insert_case_name("state v. person" + std::to_string(100 + i), cid(i));
    auto m = many.search_one(TrieIndex::CaseName, "state v.");
    CHECK(m.prefix_completions.size() == 10 && m.total_matches == 10);  // limit 10
  }
  {
    TrieIndex c;  // content tokens are lower-cased, not re-split
    c.insert_content({"Equal", "Protection", "Clause"}, DocRef{cid(1), 4, 17});
    CHECK(c.search("equal PROTECTION").prefix_completions ==
          std::vector<std::string>{"equal protection clause"});
    CHECK(c.search("EQUAL protection clause").exact_matches[0] == (DocRef{cid(1), 4, 17}));
  }
  CHECK(t.get_completions("bro", 5).empty());  // TODO stub in the reference too
  bool threw = false;
  try {
    TrieIndex::load_from_disk("x");
  } catch (const SearchError& e) {
    threw = e.kind == SearchError::NotSupported;
  }
  CHECK(threw);
}

static void trie_vs_oracle_random() {
  std::mt19937 rng(42);
  TrieIndex t;
  orc_trie_index* o = orc_trie_new();
  const char* vocab[] = {"state", "v.", "People", "united", "States", "doe", "Roe", "in", "re",
                         "smith", "Jones", "co.", "inc.", "bank", "of", "america"};
  std::vector<std::string> names;
  for (int i = 0; i < 3000; ++i) {
    int nt = 1 + rng() % 4;
    std::string s;
    for (int j = 0; j < nt; ++j) s += std::string(j ? "  " : " ") + vocab[rng() % 16];
    names.push_back(s);
    CaseId c = cid(i);
    t.insert_case_name(s, c);
    orc_trie_insert_case_name(o, s.c_str(), c.bytes.data());
    DocRef ref{c, (size_t)(i % 7), i % 3 ? std::optional<size_t>(i) : std::nullopt};
    orc_docref oref = to_orc(ref);
    t.insert_citation(s, ref);
    orc_trie_insert_citation(o, s.c_str(), &oref);
  }
  std::vector<std::string> queries = {"", "state", "STATE v.", "people", "People", "zzz", "state state"};
  for (int i = 0; i < 200; ++i) queries.push_back(names[rng() % names.size()]);
  for (int i = 0; i < 100; ++i) {
    std::string s = names[rng() % names.size()];
    queries.push_back(s.substr(0, s.find_last_of(' ') == std::string::npos ? s.size() : s.find_last_of(' ')));
  }
  for (const auto& q : queries) {
    for (int w : {0, 2}) {
      auto mine = t.search_one((TrieIndex::Which)w, q);
      orc_trie_result* ref = orc_trie_search_one(o, w, q.c_str());
      CHECK(mine.exact_matches.size() == ref->n_exact);
      for (size_t i = 0; i < mine.exact_matches.size() && i < ref->n_exact; ++i)
        CHECK(same_ref(mine.exact_matches[i], ref->exact_matches[i]));
      CHECK(mine.prefix_completions.size() == ref->n_completions);
      for (size_t i = 0; i < mine.prefix_completions.size() && i < ref->n_completions; ++i)
        CHECK(mine.prefix_completions[i] == ref->completions[i]);
      CHECK(mine.total_matches == ref->total_matches);
      orc_trie_result_free(ref);
    }
    auto mine = t.search(q);
    orc_trie_result* ref = orc_trie_search(o, q.c_str());
    CHECK(mine.exact_matches.size() == ref->n_exact && mine.total_matches == ref->total_matches);
    orc_trie_result_free(ref);
  }
  orc_trie_free(o);
}

static void no_gpu_is_loud() {
  if (tss_device_count() > 0) return;
  bool threw = false;
  try {
    VectorConfig vc;
    vc.dimension = 384;
    VectorIndex v(vc);
  } catch (const SearchError& e) {
    threw = e.kind == SearchError::VectorIndexFailed && std::string(e.category()) == "vector";
  }
  CHECK(threw);
}

// ---- GPU -------------------------------------------------------------------------------------
static std::vector<float> synth(uint64_t row, uint32_t dim, uint64_t seed) {
  std::vector<float> v(dim);
  orc_gen_rows(v.data(), row, 1, dim, seed);
  return v;
}

static void hnsw_vs_oracle() {
  const uint32_t dim = 384, n = 5000, k = 50;
  HnswIndex h(HnswConfig(), dim, 0, false);
  std::vector<float> all((size_t)n * dim);
  orc_gen_rows(all.data(), 0, n, dim, 7);
  for (uint32_t i = 0; i < n; ++i)
    h.add_vector(DocRef{cid(i / 3), i % 3, std::nullopt},
                 std::vector<float>(all.begin() + (size_t)i * dim, all.begin() + (size_t)(i + 1) * dim));
  CHECK(h.size() == n);
  auto q = synth(1, dim, 9);
  auto got = h.search(q, k);
  std::vector<uint32_t> rows(k), counts(1);
  std::vector<float> scores(k);
  orc_cosine_topk(all.data(), n, dim, q.data(), 1, k, nullptr, ORC_MASK_NONE, 0, rows.data(),
                  scores.data(), counts.data(), ORC_ORDER_CANONICAL, 0, 0);
  CHECK(got.size() == counts[0]);
  for (size_t i = 0; i < got.size(); ++i) {
    CHECK(got[i].first.case_id == cid(rows[i] / 3) && got[i].first.paragraph_index == rows[i] % 3);
    CHECK(got[i].second == 1.0f - scores[i]);  // distance, src/vector.rs:144
  }
  // incremental add after a search
  h.add_vector(DocRef{cid(99999), 0, std::nullopt}, q);
  auto again = h.search(q, 3);
  CHECK(again[0].first.case_id == cid(99999) && std::fabs(again[0].second) < 1e-6f);
  bool threw = false;
  try {
    h.add_vector(DocRef{cid(1), 0, std::nullopt}, std::vector<float>(10, 1.f));
  } catch (const SearchError& e) {
    threw = e.kind == SearchError::VectorIndexFailed;
  }
  CHECK(threw);
  threw = false;
  try {
    h.search(std::vector<float>(7, 1.f), 3);
  } catch (const SearchError& e) {
    threw = e.kind == SearchError::HnswSearchError;
  }
  CHECK(threw);
  auto batch = h.search_batch({q, synth(2, dim, 9), synth(3, dim, 9)}, 5);
  CHECK(batch.size() == 3 && batch[0][0].first.case_id == cid(99999) && batch[2].size() == 5);
  // batch policy: single queries through the bf16 shadow answer exactly like the scan
  {
    auto plain = h.search(q, 5);
    h.set_batch_policy(1, true);
    auto pre = h.search(q, 5);
    h.set_batch_policy(0, false);
    CHECK(plain.size() == pre.size());
    for (size_t i = 0; i < plain.size() && i < pre.size(); ++i)
      CHECK(plain[i].first == pre[i].first && plain[i].second == pre[i].second);
  }
  // N1: save -> load -> same answers, DocRefs included
  const std::string path = "/tmp/tss_host_shim_test_index";
  h.save(path);
  auto h2 = HnswIndex::load(HnswConfig(), path, 0);
  CHECK(h2->size() == h.size() && h2->dimension() == dim);
  auto a = h.search(q, 20), b = h2->search(q, 20);
  CHECK(a.size() == b.size());
  for (size_t i = 0; i < a.size() && i < b.size(); ++i)
    CHECK(a[i].first == b[i].first && a[i].second == b[i].second);
  CHECK(h2->rows_of_case(cid(5)) && h2->rows_of_case(cid(5))->size() == 3);
  remove((path + ".tssidx").c_str());
  remove((path + ".docrefs").c_str());
}

static void stub_embedding_behaviour() {
  // the reference's EmbeddingModel returns zeros (src/vector.rs:173): every score is 0.0 and
  // the order falls back to ascending row id
  VectorConfig vc;
  vc.dimension = 384;
  VectorIndex v(vc);
  for (int i = 0; i < 20; ++i) v.add_embedding(DocRef{cid(i), 0, std::nullopt}, synth(i, 384, 3));
  auto r = v.search("anything", 5);
  CHECK(r.size() == 5);
  for (size_t i = 0; i < r.size(); ++i) CHECK(r[i].similarity_score == 0.0f && r[i].doc_ref.case_id == cid(i));
  auto st = v.get_stats();
  CHECK(st.total_vectors == 20 && st.cache_size == 1 && st.dimension == 384);
  v.add_document(DocRef{cid(77), 0, std::nullopt}, "some text");  // embeds as zeros, still indexed
  CHECK(v.get_stats().total_vectors == 21);
}

static void engine_hybrid() {
  const uint32_t dim = 384;
  VectorConfig vc;
  vc.dimension = dim;
  auto store = std::make_shared<MetadataStore>();
  SearchEngineConfig sc;
  sc.enable_query_cache = false;
  SearchEngine eng(vc, TrieConfig(), sc, store);
  const char* names[] = {"Brown v. Board of Education", "Miranda v. Arizona", "Roe v. Wade"};
  // 60 cases x 2 paragraphs; cases 0..2 are the simple_demo ones
  for (int c = 0; c < 60; ++c) {
    CaseMetadata m;
    m.id = cid(c);
    m.name = c < 3 ? names[c] : "case " + std::to_string(c);
    m.court = c % 2 ? "scotus" : "ca9";
    m.decision_date = 1000 + c;
    store->put(m);
    eng.trie_index().insert_case_name(m.name, m.id);
    for (int p = 0; p < 2; ++p)
      eng.vector_index().add_embedding(DocRef{m.id, (size_t)p, std::nullopt}, synth(c * 2 + p, dim, 5));
  }
  // query text "miranda v. arizona" embeds next to case 7's paragraph 1
  auto target = synth(7 * 2 + 1, dim, 5);
  eng.vector_index().embedding_model().set_encoder([&](const std::string&) { return target; });
  eng.freeze();

  SearchQuery q;
  q.query = "Miranda v. Arizona";
  const size_t gets0 = store->gets(), mg0 = store->multi_gets();
  auto r = eng.search_with_params(q);
  // N4: the merge hydrates with two multi-gets (trie hits, vector hits), never per hit
  CHECK(store->gets() == gets0 && store->multi_gets() == mg0 + 2);
  // M1 shape: trie exact hit first with weight 2.0, then semantic hits >= 0.5, de-duped by case
  CHECK(r.size() >= 2 && r[0].case_metadata.id == cid(1) && r[0].score == 2.0f &&
        r[0].match_type == MatchType::Exact);
  CHECK(r[1].case_metadata.id == cid(7) && r[1].match_type == MatchType::Semantic &&
        r[1].score > 0.99f);
  for (size_t i = 2; i < r.size(); ++i) CHECK(r[i].score >= 0.5f);
  CHECK(r[0].snippet.find("paragraph 0") != std::string::npos);
  // the three policies agree whenever top-50 is not exhausted by seen cases
  eng.set_mask_policy(SearchEngine::MaskPolicy::ExcludeOnDevice);
  auto r2 = eng.search_with_params(q);
  CHECK(r2.size() == r.size());
  for (size_t i = 0; i < r.size() && i < r2.size(); ++i)
    CHECK(r2[i].case_metadata.id == r[i].case_metadata.id && r2[i].score == r[i].score);
  // prefix filter: only rows of cases under the prefix "miranda" are scored
  eng.set_mask_policy(SearchEngine::MaskPolicy::PrefixFilter);
  q.query = "miranda";
  q.config.min_similarity = -1.0f;
  auto r3 = eng.search_with_params(q);
  CHECK(r3.size() == 1 && r3[0].case_metadata.id == cid(1) && r3[0].match_type == MatchType::Semantic);
  eng.set_mask_policy(SearchEngine::MaskPolicy::PostHoc);
  // filters are post-hoc (src/search.rs:233): court filter can empty the list
  q = SearchQuery();
  q.query = "Miranda v. Arizona";
  q.court_filter = std::vector<std::string>{"ca9"};
  for (auto& x : eng.search_with_params(q)) CHECK(x.case_metadata.court == "ca9");
  q.court_filter.reset();
  q.date_range = std::make_pair(1007, 1007);
  auto rd = eng.search_with_params(q);
  CHECK(rd.size() == 1 && rd[0].case_metadata.id == cid(7));
  // N3: the same court filter applied on the device BEFORE top-k cannot be starved: with
  // min_similarity -1 the semantic pass fills max_results with ca9 cases only
  {
    SearchQuery f;
    f.query = "nothing in the trie";
    f.court_filter = std::vector<std::string>{"ca9"};
    f.config.min_similarity = -1.0f;
    auto post = eng.search_with_params(f);          // post-hoc: top-50 then filter
    eng.set_prefilter(true);
    auto pre = eng.search_with_params(f);
    CHECK(pre.size() == 10);
    for (auto& x : pre) CHECK(x.case_metadata.court == "ca9");
    for (size_t i = 0; i + 1 < pre.size(); ++i) CHECK(pre[i].score >= pre[i + 1].score);
    CHECK(post.size() <= pre.size());
    if (!post.empty()) CHECK(pre[0].case_metadata.id == post[0].case_metadata.id);
    f.court_filter = std::vector<std::string>{"no such court"};
    CHECK(eng.search_with_params(f).empty());
    f.court_filter.reset();
    f.date_range = std::make_pair(1010, 1012);
    auto dr = eng.search_with_params(f);
    CHECK(dr.size() == 3);
    for (auto& x : dr) CHECK(x.case_metadata.decision_date >= 1010 && x.case_metadata.decision_date <= 1012);
    eng.set_mask_policy(SearchEngine::MaskPolicy::ExcludeOnDevice);
    f.query = "Miranda v. Arizona";  // exact trie hit (case 1, scotus, date 1001) + filter
    f.date_range = std::make_pair(1000, 1005);
    auto both = eng.search_with_params(f);
    CHECK(!both.empty() && both[0].match_type == MatchType::Exact);
    for (auto& x : both) CHECK(x.case_metadata.decision_date >= 1000 && x.case_metadata.decision_date <= 1005);
    size_t n1 = 0;
    for (auto& x : both) n1 += x.case_metadata.id == cid(1);
    CHECK(n1 == 1);
    eng.set_mask_policy(SearchEngine::MaskPolicy::PostHoc);
    eng.set_prefilter(false);
  }
  // N4: a batch of queries gives, per query, exactly what search_with_params gives
  {
    std::vector<SearchQuery> qs(7);
    const char* texts[] = {"Miranda v. Arizona", "Roe v. Wade", "nothing here", "case 12",
                           "brown v.", "CASE 33", "Miranda v. Arizona"};
    for (int i = 0; i < 7; ++i) qs[i].query = texts[i];
    qs[2].config.min_similarity = -1.0f;
    qs[3].court_filter = std::vector<std::string>{"ca9"};
    qs[5].config.enable_semantic = false;
    qs[6].max_results = 1;
    const size_t g0 = store->gets(), m0 = store->multi_gets();
    auto batch = eng.search_batch(qs);
    CHECK(batch.size() == 7);
    CHECK(store->gets() == g0 && store->multi_gets() == m0 + 2);  // the WHOLE batch: two multi-gets
    for (int i = 0; i < 7; ++i) {
      auto one = eng.search_with_params(qs[i]);
      CHECK(one.size() == batch[i].size());
      for (size_t j = 0; j < one.size() && j < batch[i].size(); ++j)
        CHECK(one[j].case_metadata.id == batch[i][j].case_metadata.id &&
              one[j].score == batch[i][j].score && one[j].match_type == batch[i][j].match_type);
    }
  }
  // N4: the micro-batcher answers concurrent submitters with batched device calls
  {
    QueryBatcher batcher(eng, 16, 2000);
    const char* texts[] = {"Miranda v. Arizona", "Roe v. Wade", "nothing here", "case 12", "x"};
    std::vector<std::thread> threads;
    std::vector<std::vector<std::future<std::vector<SearchResult>>>> futs(4);
    for (int t = 0; t < 4; ++t)
      threads.emplace_back([&, t] {
        for (int i = 0; i < 10; ++i) {
          SearchQuery q;
          q.query = texts[(t + i) % 5];
          futs[t].push_back(batcher.submit(q));
        }
      });
    for (auto& th : threads) th.join();
    size_t answered = 0, rejected = 0;
    for (int t = 0; t < 4; ++t)
      for (int i = 0; i < 10; ++i) {
        SearchQuery q;
        q.query = texts[(t + i) % 5];
        try {
          auto got = futs[t][i].get();
          auto want = eng.search_with_params(q);
          CHECK(got.size() == want.size());
          for (size_t j = 0; j < got.size() && j < want.size(); ++j)
            CHECK(got[j].case_metadata.id == want[j].case_metadata.id && got[j].score == want[j].score);
          ++answered;
        } catch (const SearchError& e) {
          CHECK(e.kind == SearchError::InvalidSearchQuery && q.query == "x");
          ++rejected;
        }
      }
    CHECK(answered == 32 && rejected == 8);
    CHECK(batcher.queries_run() == 40 && batcher.batches_run() < 40);
  }
  // M2: >= max_results exact hits skip the vector pass entirely
  q = SearchQuery();
  q.query = "Roe v. Wade";
  q.config.max_results = 1;
  auto rm = eng.search_with_params(q);
  CHECK(rm.size() == 1 && rm[0].match_type == MatchType::Exact);
  // validation, src/search.rs:284-300
  bool threw = false;
  try {
    eng.search("a");
  } catch (const SearchError& e) {
    threw = e.kind == SearchError::InvalidSearchQuery;
  }
  CHECK(threw);
  // semantic disabled
  q = SearchQuery();
  q.query = "Miranda v. Arizona";
  q.config.enable_semantic = false;
  CHECK(eng.search_with_params(q).size() == 1);
  q.config.enable_semantic = true;
  q.config.enable_prefix = false;
  auto rs = eng.search_with_params(q);
  CHECK(!rs.empty() && rs[0].case_metadata.id == cid(7) && rs[0].match_type == MatchType::Semantic);
}

// TrieIndex::save_to_disk / load_from_disk (stubs in the reference, src/trie.rs:83-94): the
// loaded index answers every search like the one that was saved; with a GPU the frozen
// (flattened, device-resident) form round-trips too and yields the same prefix masks.
static void trie_disk_round_trip(bool gpu) {
  TrieIndex t;
  std::mt19937 rng(7);
  const char* vocab[] = {"state", "v.", "People", "united", "States", "doe", "Roe", "in", "re", "smith"};
  std::vector<std::string> names;
  std::unordered_map<CaseId, std::vector<uint32_t>, CaseIdHash> rows;
  for (int i = 0; i < 500; ++i) {
    std::string s;
    int nt = 1 + rng() % 4;
    for (int j = 0; j < nt; ++j) s += std::string(j ? " " : "") + vocab[rng() % 10];
    names.push_back(s);
    t.insert_case_name(s, cid(i));
    t.insert_citation(std::to_string(100 + i % 37) + " U.S. " + std::to_string(i), DocRef{cid(i), 1, (size_t)i});
    t.insert_content({"Equal", vocab[rng() % 10]}, DocRef{cid(i), (size_t)(i % 5), std::nullopt});
    rows[cid(i)] = {(uint32_t)(2 * i), (uint32_t)(2 * i + 1)};
  }
  auto rows_of = [&](const CaseId& c) -> const std::vector<uint32_t>* {
    auto it = rows.find(c);
    return it == rows.end() ? nullptr : &it->second;
  };
  if (gpu)
    for (int w = 0; w < 3; ++w) t.freeze((TrieIndex::Which)w, 0, rows_of);
  char path[] = "/tmp/tss_trie_XXXXXX";
  int fd = mkstemp(path);
  CHECK(fd >= 0);
  if (fd >= 0) close(fd);
  t.save_to_disk(path);
  TrieIndex u = TrieIndex::load_from_disk(path, 0);
  for (int w = 0; w < 3; ++w) {
    CHECK(u.trie((TrieIndex::Which)w).terms() == t.trie((TrieIndex::Which)w).terms());
    remove((std::string(path) + "." + std::to_string(w) + ".terms").c_str());
  }
  for (const char* q : {"state v.", "People", "100 U.S.", "equal", "equal doe", "nope", ""}) {
    auto a = t.search(q), b = u.search(q);
    CHECK(a.exact_matches == b.exact_matches && a.prefix_completions == b.prefix_completions &&
          a.total_matches == b.total_matches);
  }
  if (gpu) {
    tss_mask *ma = nullptr, *mb = nullptr;
    CHECK(tss_mask_create(&ma, 1000, 0) == 0 && tss_mask_create(&mb, 1000, 0) == 0);
    std::vector<uint32_t> wa(32), wb(32);
    for (int w = 0; w < 3; ++w)
      for (const char* q : {"state", "state v.", "100 U.S.", "equal", ""}) {
        t.prefix_mask((TrieIndex::Which)w, q, ma, 0, true);
        u.prefix_mask((TrieIndex::Which)w, q, mb, 0, true);
        CHECK(tss_mask_download(ma, wa.data()) == 0 && tss_mask_download(mb, wb.data()) == 0);
        CHECK(wa == wb);
      }
    tss_mask_destroy(ma);
    tss_mask_destroy(mb);
  }
  remove(path);
  // a foreign file is refused with the error the reference's stub raises
  bool threw = false;
  try {
    TrieIndex::load_from_disk("/proc/self/cmdline");
  } catch (const SearchError& e) {
    threw = e.kind == SearchError::NotSupported;
  }
  CHECK(threw);
}


// S1/S2 end to end, randomized: SearchEngine (trie cascade + device scan + merge + multi-get
// hydration) against the oracle's literal pieces glued the way execute_hybrid_search glues them
// (src/search.rs:185-240): orc_trie_search -> exact case ids; orc_cosine_topk (top-50, the
// seen-case rows excluded first when the policy pushes de-dup onto the device) -> vector hits;
// orc_hybrid_merge.  Case ids, score bits and match types must agree for every query.
static uint64_t cid_u64(const CaseId& c) {
  uint64_t v = 0;
  for (int i = 0; i < 8; ++i) v |= (uint64_t)c.bytes[15 - i] << (8 * i);
  return v;
}
static void engine_vs_oracle_random() {
  const uint32_t dim = 128, ncases = 300;
  std::mt19937 rng(2024);
  VectorConfig vc;
  vc.dimension = dim;
  auto store = std::make_shared<MetadataStore>();
  SearchEngineConfig sc;
  sc.enable_query_cache = false;
  SearchEngine eng(vc, TrieConfig(), sc, store);
  orc_trie_index* otrie = orc_trie_new();
  const char* vocab[] = {"state", "v.", "People", "united", "States", "doe", "Roe", "in", "re", "smith"};
  std::vector<std::string> names;
  std::vector<float> all;          // every row's embedding, in row order
  std::vector<uint64_t> row_case;  // row -> case number
  for (uint32_t c = 0; c < ncases; ++c) {
    std::string name;
    int nt = 1 + rng() % 3;
    for (int j = 0; j < nt; ++j) name += std::string(j ? " " : "") + vocab[rng() % 10];
    names.push_back(name);  // few tokens from a small vocabulary: many cases share a name
    CaseMetadata m;
    m.id = cid(c);
    m.name = name;
    m.court = c % 3 ? "scotus" : "ca9";
    m.decision_date = 1000 + (int)c;
    if (c % 17 != 5) store->put(m);  // some cases have no metadata: skipped like `if let Ok(Some(..))`
    eng.trie_index().insert_case_name(name, m.id);
    orc_trie_insert_case_name(otrie, name.c_str(), m.id.bytes.data());
    const int paragraphs = 1 + rng() % 3;
    for (int p = 0; p < paragraphs; ++p) {
      // clusters: cases c and c+100 are close in embedding space, so vector hits pile up on few cases
      auto e = synth((c % 100) * 7 + p, dim, 5);
      auto noise = synth(c * 31 + p, dim, 11);
      for (uint32_t j = 0; j < dim; ++j) e[j] += 0.05f * noise[j];
      eng.vector_index().add_embedding(DocRef{m.id, (size_t)p, std::nullopt}, e);
      all.insert(all.end(), e.begin(), e.end());
      row_case.push_back(c);
    }
  }
  const uint32_t nrows = (uint32_t)row_case.size();
  std::unordered_map<std::string, std::vector<float>> embed_of;
  eng.vector_index().embedding_model().set_encoder(
      [&](const std::string& text) { return embed_of.at(text); });
  eng.freeze();
  std::vector<std::string> queries;
  for (int i = 0; i < 40; ++i) {
    std::string q = i % 2 ? names[rng() % ncases] : std::string(vocab[rng() % 10]) + " " + vocab[rng() % 10];
    if (i % 7 == 0) q = "  " + q + " ";  // whitespace / case noise for the tokeniser
    if (q.size() < 2) q += "xx";
    auto e = synth((rng() % 100) * 7 + rng() % 3, dim, 5);
    auto noise = synth(9000 + i, dim, 13);
    for (uint32_t j = 0; j < dim; ++j) e[j] += 0.2f * noise[j];
    embed_of[q] = e;
    queries.push_back(q);
  }
  for (auto policy : {SearchEngine::MaskPolicy::PostHoc, SearchEngine::MaskPolicy::ExcludeOnDevice}) {
    eng.set_mask_policy(policy);
    for (size_t qi = 0; qi < queries.size(); ++qi) {
      SearchQuery q;
      q.query = queries[qi];
      q.config.min_similarity = qi % 3 ? 0.5f : 0.2f;
      if (qi % 5 == 0) q.max_results = 3;
      auto got = eng.search_with_params(q);
      // --- the oracle's version ---
      orc_trie_result* tr = orc_trie_search(otrie, q.query.c_str());
      std::vector<uint64_t> exact;
      std::vector<uint32_t> mask_words((nrows + 31) / 32, 0);
      std::unordered_set<uint64_t> seen;
      for (uint64_t i = 0; i < tr->n_exact; ++i) {
        CaseId c;
        memcpy(c.bytes.data(), tr->exact_matches[i].case_id, 16);
        if (!store->get_case_metadata(c)) continue;  // :193 skips a case without metadata
        exact.push_back(cid_u64(c));
        seen.insert(cid_u64(c));
      }
      orc_trie_result_free(tr);
      int mode = ORC_MASK_NONE;
      if (policy == SearchEngine::MaskPolicy::ExcludeOnDevice && !seen.empty()) {
        for (uint32_t r = 0; r < nrows; ++r)
          if (seen.count(row_case[r])) mask_words[r >> 5] |= 1u << (r & 31);
        mode = ORC_MASK_EXCLUDE;
      }
      const uint32_t k = (uint32_t)SearchEngine::kVectorTopK;
      std::vector<uint32_t> rows(k), counts(1);
      std::vector<float> scores(k);
      orc_cosine_topk(all.data(), nrows, dim, embed_of[q.query].data(), 1, k, mask_words.data(), mode,
                      0, rows.data(), scores.data(), counts.data(), ORC_ORDER_CANONICAL, 0, 0);
      std::vector<uint64_t> vcase;
      std::vector<float> vscore;
      for (uint32_t i = 0; i < counts[0]; ++i) {
        if (!store->get_case_metadata(cid(row_case[rows[i]]))) continue;  // :213
        vcase.push_back(row_case[rows[i]]);
        vscore.push_back(1.0f - (1.0f - scores[i]));  // VectorIndex::search: 1 - distance, :144
      }
      std::vector<orc_hit> want(128);
      const uint32_t nw = orc_hybrid_merge(exact.data(), (uint32_t)exact.size(), vcase.data(), vscore.data(),
                                           (uint32_t)vcase.size(), 1, 1, (uint32_t)q.config.max_results,
                                           q.max_results ? (int64_t)*q.max_results : -1,
                                           q.config.min_similarity, q.config.exact_match_weight,
                                           want.data(), 128);
      CHECK(got.size() == nw);
      for (size_t i = 0; i < got.size() && i < nw; ++i) {
        CHECK(cid_u64(got[i].case_metadata.id) == want[i].case_id);
        CHECK(memcmp(&got[i].score, &want[i].score, 4) == 0);
        CHECK((got[i].match_type == MatchType::Exact) == (want[i].match_type == 0));
      }
    }
  }
  eng.set_mask_policy(SearchEngine::MaskPolicy::PostHoc);
  orc_trie_free(otrie);
}

int main(int argc, char** argv) {
  std::string mode = argc > 1 ? argv[1] : "cpu";
  try {
    trie_kats();
    trie_vs_oracle_random();
    trie_disk_round_trip(mode == "gpu");
    if (mode == "gpu") {
      hnsw_vs_oracle();
      stub_embedding_behaviour();
      engine_hybrid();
      engine_vs_oracle_random();
    } else {
      no_gpu_is_loud();
    }
  } catch (const std::exception& e) {
    printf("FAIL: exception: %s\n", e.what());
    return 2;
  }
  printf(g_fail ? "HOST_SHIM_FAILED (%d)\n" : "HOST_SHIM_OK\n", g_fail);
  return g_fail ? 1 : 0;
}
