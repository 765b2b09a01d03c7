/*
 * tss.h -- C ABI of libtss.so: the B200-native scoring / top-k / prefix-mask
 * path of trie-semantic-search.
 *
 * Every entry point is extern "C", takes plain pointers and sizes, never
 * throws or unwinds (the reference's release profile is panic = "abort",
 * Cargo.toml:77), and returns an int status (TSS_OK == 0).  A thread-local
 * message for the last failure is available from tss_last_error(); the Rust
 * shim maps it to SearchError::HnswSearchError{details} /
 * VectorIndexFailed{reason} (src/errors.rs:149-153).
 *
 * The reference has no FFI of its own.  The seam this library sits behind is
 * the private struct HnswIndex in src/vector.rs:184-208 (new / add_vector /
 * search / size) and, on the trie side, TrieIndex::{insert_*,search}
 * (src/trie.rs:97-130).  Each declaration below cites what it replaces.
 * There is NO CPU fallback: without a CUDA device every call that would
 * touch one returns TSS_ERR_CUDA.
 *
 * Threading: every tss_index entry point takes the handle's own lock, so any
 * number of host threads may call into one index (searches included) with
 * distinct output buffers, instead of queueing behind the caller's
 * process-wide write lock (the reference: src/search.rs:249-252).  A search of
 * up to TSS_PENDING_MAX_NQ queries holds the lock only while it is ENQUEUED
 * and waits for its result outside it (each call owns a result slot and an
 * event), so the scans of concurrent callers run back to back on the device;
 * tss_index_search_submit / _collect give one thread the same pipelining.
 * One scan already saturates HBM, so beyond hiding launch and wake-up latency
 * concurrency is turned into throughput by BATCHING (nq > 1: the corpus is
 * streamed once per 4 queries, or once per batch on the tensor-core path;
 * host/tss_host.hpp QueryBatcher does that for concurrent callers).
 * A sharded index is collective: every rank must issue its searches in the
 * same order, so drive it from one thread per rank.
 * tss_mask / tss_terms / tss_columns handles: one thread mutates a given
 * handle at a time; concurrent searches may read the same mask.  Distinct
 * handles may always be used from distinct threads.
 */
#ifndef TSS_H
#define TSS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* libtss is built with -fvisibility=hidden */
#endif

#define TSS_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------ */
enum {
  TSS_OK = 0,
  TSS_ERR_INVALID_ARG = 1, /* -> SearchError::InvalidSearchQuery / VectorIndexFailed */
  TSS_ERR_CUDA = 2,        /* -> SearchError::HnswSearchError{details} */
  TSS_ERR_NCCL = 3,
  TSS_ERR_OOM = 4,
  TSS_ERR_STATE = 5 /* call order violated (e.g. search before finalize) */
};

/* storage type of the embedding matrix in HBM */
enum { TSS_F32 = 0, TSS_BF16 = 1 };

/* polarity of a row mask handed to a search (SURVEY.md section 8a "semantic gap"):
 * EXCLUDE is the reference's seen_cases de-dup (src/search.rs:187,214),
 * INCLUDE is the prefix filter of BASELINE.json config 4. */
enum { TSS_MASK_NONE = 0, TSS_MASK_INCLUDE = 1, TSS_MASK_EXCLUDE = 2 };

/* how a prefix is matched against the flattened term array */
enum {
  TSS_PREFIX_TOKEN = 0, /* token-level, exactly TrieNode::search's walk (src/trie.rs:223-238) */
  TSS_PREFIX_CHAR = 1   /* byte prefix of the joined term (README.md:41-44 behaviour) */
};

#define TSS_MAX_K 1024u         /* largest k any search accepts */
#define TSS_MAX_FUSED_K 128u    /* k <= this runs the fused in-scan top-k */
#define TSS_ROW_NONE 0xFFFFFFFFu /* row id of an unused output slot */
#define TSS_MAX_PENDING 4u      /* tss_index_search_submit calls in flight per index */
#define TSS_PENDING_MAX_NQ 4u   /* queries per submitted search (one scan launch) */

typedef struct tss_index tss_index; /* replaces HnswIndex, src/vector.rs:40-44 */
typedef struct tss_mask tss_mask;   /* device bitmask over the rows of one shard */
typedef struct tss_terms tss_terms; /* flattened, byte-sorted term array + CSR postings */
typedef struct tss_comm tss_comm;   /* one rank of a row-sharded index (NCCL) */
typedef struct tss_columns tss_columns; /* per-row metadata columns of one shard (N3) */

/* ---- library --------------------------------------------------------------*/
int tss_abi_version(void);
const char* tss_last_error(void);
/* number of CUDA devices visible; 0 (not an error) when there is none. */
int tss_device_count(void);

/* ---- index lifecycle -------------------------------------------------------
 * tss_index_create  <- HnswIndex::new(HnswConfig)            src/vector.rs:185-188
 * tss_index_add     <- HnswIndex::add_vector(DocRef,Vec<f32>) src/vector.rs:190-193
 *                      (rows are copied before return; the DocRef stays with
 *                       the caller: row i of the index is the i-th row added)
 * tss_index_size    <- HnswIndex::size()                      src/vector.rs:204-207
 * tss_index_destroy <- Drop
 */
int tss_index_create(tss_index** out, uint32_t dim, int storage, int device);
/* optional: size the HBM allocation once (HnswConfig.max_elements, src/config.rs:568). */
int tss_index_reserve(tss_index* ix, uint64_t nrows);
/* nrows x dim fp32, row-major, host memory.  NaN / Inf anywhere -> TSS_ERR_INVALID_ARG
 * and the index is unchanged. */
int tss_index_add(tss_index* ix, const float* rows, uint64_t nrows);
/* test / bench utility: append rows [row_begin,row_begin+nrows) of the seeded
 * synthetic corpus, generated on the device (same integer hash as
 * oracle/oracle.cpp:gen_row). */
int tss_index_add_synthetic(tss_index* ix, uint64_t row_begin, uint64_t nrows, uint64_t seed);
/* completes pending uploads; searches are legal only after this.  More rows may
 * be added afterwards (finalize again before the next search). */
int tss_index_finalize(tss_index* ix);
uint64_t tss_index_size(const tss_index* ix);
uint32_t tss_index_dim(const tss_index* ix);
void tss_index_destroy(tss_index* ix);
/* On-disk form: a 64-byte header + the padded rows exactly as they sit in
 * HBM, so tss_index_load is file -> pinned buffer -> HBM with no repacking.  These fill in
 * VectorIndex::save_to_disk / load_from_disk (src/vector.rs:83-95, TODO stubs); the row ->
 * DocRef table stays with the caller.  The loaded index is finalized. */
int tss_index_save(tss_index* ix, const char* path);
int tss_index_load(tss_index** out, const char* path, int device);
/* copy stored rows back (fp32; bf16 storage is widened). Test utility. */
int tss_index_get_rows(tss_index* ix, uint64_t row_begin, uint64_t nrows, float* out);

/* ---- search ------------------------------------------------------------------
 * tss_index_search <- HnswIndex::search(&[f32], top_k) -> Vec<(DocRef, f32)>
 *                     src/vector.rs:195-202, batched over nq queries.
 * queries: nq x dim fp32 host memory.  Outputs (host, caller-allocated, nq*k
 * each): rows best-first under (score desc, row asc); scores are cosine
 * SIMILARITY (the shim returns 1 - s as "distance" to keep the reference
 * signature, src/vector.rs:144); out_counts[q] <= k valid slots, the rest are
 * TSS_ROW_NONE / 0.0.  mask may be NULL (mode TSS_MASK_NONE).  For a sharded
 * index (tss_index_set_shard) the call is collective over the comm and every
 * rank receives the merged global result; row ids are global.
 * A zero-norm query or row scores 0.0 (the reference's stub embeds every text
 * as zeros, src/vector.rs:173, so this is reachable).
 */
int tss_index_search(tss_index* ix, const float* queries, uint32_t nq, uint32_t k,
                     const tss_mask* mask, int mask_mode, uint32_t* out_rows,
                     float* out_scores, uint32_t* out_counts);

/* The same search split in two for a caller that pipelines: submit copies the
 * queries (1..TSS_PENDING_MAX_NQ of them, k <= TSS_MAX_FUSED_K), enqueues the
 * scan and returns a ticket without waiting; collect waits for that search
 * alone and unpacks its result.  Up to TSS_MAX_PENDING searches may be in
 * flight per index (one more submit returns TSS_ERR_STATE); they execute in
 * submission order, and under programmatic dependent launch the next scan's
 * prologue overlaps the previous one's merges, so a caller that keeps two in
 * flight sees the device-resident rate through host pointers.  Always the
 * exact scan (K1), whatever tss_index_set_batch_policy says.  A ticket is
 * collected exactly once.  A mask handed to submit may be rewritten right
 * away: mask writes are stream-ordered behind the searches that read the mask.
 * The reference's HnswIndex::search
 * (src/vector.rs:195-202) is an async fn that never awaits; this is the form
 * an awaiting implementation would take. */
int tss_index_search_submit(tss_index* ix, const float* queries, uint32_t nq, uint32_t k,
                            const tss_mask* mask, int mask_mode, uint64_t* out_ticket);
int tss_index_search_collect(tss_index* ix, uint64_t ticket, uint32_t* out_rows,
                             float* out_scores, uint32_t* out_counts);

/* Same work with every buffer already in HBM and no host synchronisation:
 * d_queries nq x dim fp32, d_out_keys nq x k packed keys
 * ((orderable(score) << 32) | ~row, 0 = unused slot), enqueued on the index
 * stream.  This is what `value` in bench.py times.
 * The tensor-core path redoes a query whose survivor list overflowed with the
 * exact scan; here that happens ON THE DEVICE: a few guarded scan launches
 * are enqueued behind every batch and each serves one flagged query if there
 * is one (none on ordinary data).  Should a batch flag more queries than
 * launches were enqueued, tss_index_sync reports TSS_ERR_STATE (those queries'
 * slots are not final) and later batches get twice as many.  One exception:
 * k > TSS_MAX_FUSED_K, whose redo runs by rounds from the host, synchronises. */
int tss_index_search_device(tss_index* ix, const float* d_queries, uint32_t nq, uint32_t k,
                            const tss_mask* mask, int mask_mode, uint64_t* d_out_keys);
/* Which batches take the tensor-core path (K2), whose results are bit-identical to the scan's
 * (candidates from bf16 tensor-core scores within a rigorous error margin are re-scored with the
 * scan's arithmetic).  Default: batches of >= 16 queries; and >= 3 queries on a corpus of >= 2M
 * rows once the bf16 matrix K2 reads exists (a bf16 index, or an fp32 index that has built its
 * bf16 shadow, +50 % memory, on its first K2 batch).
 *   min_queries >= 1: every batch of at least that many queries leaves the plain scan (0
 *     restores the default).  With 1, an fp32 index that has its shadow answers calls of one or
 *     two queries (k <= 32) by scanning the SHADOW for the top-64/128, proving the fp32 top-k is
 *     among them, and re-scoring those from the fp32 rows -- 1.17 ms instead of 2.05 ms over
 *     10M x 384, same bits; a query the proof fails for is redone by the fp32 scan.
 *   build_shadow_now != 0: the index builds its shadow inside this call (TSS_ERR_OOM if it does
 *     not fit; large batches of an fp32 index then stay on the scan).  The shadow is a bf16 copy
 *     of the matrix with every row scaled to unit length: the tensor cores read it, the
 *     accumulator is the score and the GEMM epilogue needs no per-row weight.  A bf16 index
 *     (+100 % memory) builds one on its own at its first large batch only when that leaves plenty
 *     of memory free; otherwise its stored rows are read and weighted by 1/|row|.
 * New entry: the reference has no batched or two-stage search (src/vector.rs:195-202). */
int tss_index_set_batch_policy(tss_index* ix, uint32_t min_queries, int build_shadow_now);
/* unpack keys produced by tss_index_search_device (host side, pure function). */
void tss_unpack_keys(const uint64_t* keys, uint64_t n, uint32_t* out_rows, float* out_scores);

/* ---- sharding (BASELINE.json config 5; no reference analogue) ---------------
 * One process per GPU.  Rank r owns global rows [row_base, row_base + size).
 * The only exchange on the path is nq*k packed keys per rank and a k-way merge on
 * every rank: fused into the scan's last CTA over NVLink peer memory (batch-1 ..
 * 4-query scans), or ncclAllGather + a merge kernel (tensor-core batches, > 8
 * ranks, no IPC). */
int tss_comm_unique_id(uint8_t out_id[128]);
int tss_comm_create(tss_comm** out, const uint8_t id[128], int rank, int nranks, int device);
void tss_comm_destroy(tss_comm* c);
/* Collective over the comm the first time a comm is attached to an index (the ranks map each
 * other's exchange buffers with CUDA IPC so the scan kernel can push its top-k straight into
 * peer memory over NVLink and merge in its last CTA; TSS_FUSED_XCHG=0 or an IPC failure falls
 * back to ncclAllGather + a merge kernel).  comm == NULL detaches. */
int tss_index_set_shard(tss_index* ix, uint64_t row_base, tss_comm* comm /* nullable */);

/* ---- masks ------------------------------------------------------------------
 * Ordering: every call that writes a mask (create, clear, set/clear_rows, upload,
 * tss_filter_mask, tss_prefix_mask*) and every call that reads one (a search,
 * download, popcount) is stream-ordered against the others through events the
 * mask carries -- a search enqueued after tss_prefix_mask sees the finished
 * mask, a clear enqueued after a search waits for that search -- whatever
 * streams the handles involved run on, and without host synchronisation
 * (download / popcount / upload block because they hand data to the host).
 * One thread mutates a given mask at a time; concurrent searches may read it. */
/* nbits rows of ONE shard; bit i <-> local row i (global row row_base + i). */
int tss_mask_create(tss_mask** out, uint64_t nbits, int device);
int tss_mask_clear(tss_mask* m);
/* set the bits of the listed GLOBAL rows that fall inside [row_base,row_base+nbits);
 * this is how the shim turns seen_cases (src/search.rs:187-206) into an
 * EXCLUDE mask. */
int tss_mask_set_rows(tss_mask* m, const uint32_t* rows, uint64_t n, uint64_t row_base);
int tss_mask_clear_rows(tss_mask* m, const uint32_t* rows, uint64_t n, uint64_t row_base);
int tss_mask_upload(tss_mask* m, const uint32_t* words /* ceil(nbits/32) */);
int tss_mask_download(const tss_mask* m, uint32_t* words /* ceil(nbits/32) */);
int tss_mask_popcount(const tss_mask* m, uint64_t* out);
uint64_t tss_mask_nbits(const tss_mask* m);
void tss_mask_destroy(tss_mask* m);

/* ---- metadata pre-filter --------------------------------------------
 * SearchEngine::apply_filters (src/search.rs:255-274) drops results AFTER the top-50 + sort, so
 * a selective court / date filter can return fewer than max_results.  With the two columns the
 * filter reads resident on the device (u16 court id assigned by the host, i32 decision date in
 * days), tss_filter_mask writes the rows that pass -- date in [date_lo, date_hi] AND (no court
 * list OR court in the list), the same conjunction as apply_filters -- into a mask that the
 * search takes as TSS_MASK_INCLUDE.  combine_and != 0 intersects with the mask's current
 * content (e.g. a prefix mask) instead of overwriting it.  6 bytes per row of traffic. */
int tss_columns_create(tss_columns** out, const uint16_t* court_ids, const int32_t* dates,
                       uint64_t nrows, int device);
void tss_columns_destroy(tss_columns* c);
int tss_filter_mask(tss_columns* c, const uint16_t* allowed_courts, uint32_t n_allowed,
                    int32_t date_lo, int32_t date_hi, tss_mask* mask, int combine_and);

/* ---- flattened trie ---------------------------------------------------------
 * tss_terms_create <- the populated TrieNode tree (src/trie.rs:51-57,211-221),
 *   exported as T unique terms (tokens joined by one ' '), byte-sorted,
 *   pool + offsets[T+1]; postings CSR post_off[T+1] / post_rows (global rows,
 *   insertion order, duplicates kept as in document_refs.push, src/trie.rs:219).
 * tss_prefix_mask  <- TrieNode::search's walk (src/trie.rs:223-238) + the subtree
 *   that collect_completions enumerates (src/trie.rs:257-278), without its
 *   limit-10 and including the matched node itself: sets bit (row - row_base)
 *   for every posting of every matching term.  The mask is NOT cleared first.
 */
typedef struct tss_prefix_stats {
  uint64_t exact_lo, exact_hi; /* [lo,hi) term range equal to the prefix (0 or 1 term) */
  uint64_t sub_lo, sub_hi;     /* [lo,hi) term range strictly below it */
  uint64_t npostings;          /* postings visited (before the shard filter) */
} tss_prefix_stats;

int tss_terms_create(tss_terms** out, const char* pool, const uint64_t* term_off,
                     const uint64_t* post_off, const uint32_t* post_rows, uint64_t nterms,
                     int device);
/* N2: build the same structure on the device from tokenised postings instead
 * of TrieNode::insert's one-HashMap-hop-per-token (src/trie.rs:211-221).  The vocabulary is V
 * unique tokens in byte order (pool + offsets[V+1]; no byte <= 0x20, so tuple order equals the
 * byte order of the ' '-joined strings); posting i is the tuple token_ids[i*max_tokens ..]
 * (id = vocabulary index + 1, 0 = padding at the end; an all-zero tuple is the root) with row
 * rows[i].  Postings of a term keep their input order (document_refs.push, src/trie.rs:219). */
int tss_terms_build(tss_terms** out, const char* vocab_pool, const uint64_t* vocab_off,
                    uint32_t vocab_size, const uint32_t* token_ids, uint32_t max_tokens,
                    const uint32_t* rows, uint64_t n_postings, int device);
/* N2 from raw strings: what feeds TrieIndex::insert_case_name / insert_content / insert_citation
 * (src/trie.rs:97-110) is text; the wrappers tokenise with split_whitespace() and lower-case
 * case names and content (src/trie.rs:147,158,171,177) but not citations (:190,196).  Here the
 * n phrases of one trie arrive as one byte buffer (text + phrase_off[n + 1]), phrase i posting
 * rows[i]; the DEVICE splits them at ASCII whitespace (0x09-0x0D, 0x20), lower-cases A-Z when
 * lowercase != 0, dictionary-encodes the tokens (string sort by radix passes over 8-byte chunks,
 * unique -> the byte-sorted vocabulary) and builds the flattened trie as tss_terms_build does.
 * A phrase without tokens posts on the root term "".  Limits (TSS_ERR_INVALID_ARG): tokens of at
 * most 128 bytes, at most max_tokens (<= 64) tokens per phrase, no control bytes below 0x20
 * other than whitespace.  Bytes >= 0x80 pass through unchanged: case folding and whitespace are
 * ASCII where the reference's are Unicode. */
int tss_terms_build_text(tss_terms** out, const char* text, const uint64_t* phrase_off,
                         const uint32_t* rows, uint64_t n_phrases, int lowercase, uint32_t max_tokens,
                         int device);
/* On-disk form of the flattened term array (second half): fills in
 * TrieIndex::save_to_disk / load_from_disk (src/trie.rs:83-94: a no-op and a NotSupported stub;
 * TrieConfig.index_path, src/config.rs:190).  64-byte header + term_off + post_off + post_rows +
 * pool exactly as they sit in HBM; the loader checks the sizes against the file and validates
 * ordering on the device. */
int tss_terms_save(const tss_terms* t, const char* path);
int tss_terms_load(tss_terms** out, const char* path, int device);
int tss_terms_sizes(const tss_terms* t, uint64_t* nterms, uint64_t* pool_bytes, uint64_t* npostings);
/* copy the flattened arrays back (caller-allocated per tss_terms_sizes) */
int tss_terms_export(const tss_terms* t, char* pool, uint64_t* term_off, uint64_t* post_off,
                     uint32_t* post_rows);
uint64_t tss_terms_size(const tss_terms* t);
void tss_terms_destroy(tss_terms* t);
int tss_prefix_mask(tss_terms* t, const char* prefix, uint32_t len, int kind, tss_mask* out,
                    uint64_t row_base, tss_prefix_stats* stats /* nullable: no host sync */);
/* tss_mask_clear + tss_prefix_mask as ONE cooperative launch: CTAs 0-3 find the four bounds with
 * a 256-ary search while every CTA zeroes its share of the mask, a grid barrier, then all CTAs
 * scatter the posting ranges (and, for <= 131 072 postings, list the unique rows for the
 * list-driven scan).  No host synchronisation (stats == NULL); what the hybrid path issues per
 * query (SearchEngine::execute_hybrid_search, src/search.rs:185-206 builds its seen-set the same
 * way: from scratch per query). */
int tss_prefix_mask_fresh(tss_terms* t, const char* prefix, uint32_t len, int kind, tss_mask* out,
                          uint64_t row_base, tss_prefix_stats* stats /* nullable: no host sync */);
/* Enqueue this handle's prefix searches on the index's stream (ix == NULL: back on its own):
 * prefix -> mask -> masked search then run back to back on ONE stream with no cross-stream
 * event hop.  The index must outlive the binding. */
int tss_terms_bind_stream(tss_terms* t, tss_index* ix /* nullable */);
/* The hybrid query of BASELINE.json config 4 as ONE host call: tss_prefix_mask_fresh(t, prefix,
 * ..., scratch, row base of ix) followed by tss_index_search(ix, queries, ..., scratch,
 * TSS_MASK_INCLUDE, ...) -- prefix search + mask + masked top-k are enqueued together and the
 * host waits once.  scratch is overwritten (a mask of >= tss_index_size(ix) bits on the index's
 * device); afterwards it holds the prefix's row set.  Replaces the trie walk + vector search
 * pair of SearchEngine::execute_hybrid_search (src/search.rs:189-190, 210) when the prefix is
 * used as a filter. */
int tss_index_search_prefix(tss_index* ix, tss_terms* t, const char* prefix, uint32_t len, int kind,
                            tss_mask* scratch, const float* queries, uint32_t nq, uint32_t k,
                            uint32_t* out_rows, float* out_scores, uint32_t* out_counts);
/* ... and without waiting: the prefix search, the mask and the masked scan are enqueued and a
 * ticket for tss_index_search_collect comes back (1..TSS_PENDING_MAX_NQ queries, k <=
 * TSS_MAX_FUSED_K, as tss_index_search_submit).  Each hybrid query in flight needs a scratch mask
 * of its own until it has been collected; with two in flight the next query's prefix search runs
 * while this one's rows stream. */
int tss_index_search_prefix_submit(tss_index* ix, tss_terms* t, const char* prefix, uint32_t len,
                                   int kind, tss_mask* scratch, const float* queries, uint32_t nq,
                                   uint32_t k, uint64_t* out_ticket);

/* ---- plumbing for callers that time or pipeline the device path ------------ */
void* tss_index_stream(tss_index* ix); /* cudaStream_t the index enqueues on */
/* waits for the index stream; also reports (TSS_ERR_NCCL) a sharded tss_index_search_device
 * whose peers never delivered their top-k */
int tss_index_sync(tss_index* ix);
int tss_dev_alloc(int device, uint64_t bytes, void** out);
int tss_dev_free(int device, void* p);
int tss_dev_h2d(int device, void* dst, const void* src, uint64_t bytes);
int tss_dev_d2h(int device, void* dst, const void* src, uint64_t bytes);
int tss_event_create(int device, void** out);
int tss_event_record(tss_index* ix, void* ev);
int tss_event_elapsed_ms(void* ev_a, void* ev_b, float* out_ms); /* synchronises on ev_b */
int tss_event_destroy(void* ev);
/* diagnostics: when d_buf (device, 256*8 u64) is non-NULL every scan CTA writes %globaltimer
 * stamps of its phases into it (slot = cta*8 + phase); NULL switches it off. */
int tss_index_debug_phases(tss_index* ix, void* d_buf);
/* kernels launched by this library in this process so far (bench's gpu_launches). */
uint64_t tss_launch_count(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* TSS_H */
