//! tss.rs -- the Rust side of the boundary: `extern "C"` block for libtss.so plus the
//! replacement bodies of the four `HnswIndex` methods (reference src/vector.rs:184-208).
//!
//! SOURCE ONLY: this build environment has no cargo/rustc and the
//! reference crate does not compile as written (F3), so this file has never been compiled.
//! It is the binding a maintainer drops into `src/` next to `vector.rs`; the same logic is
//! exercised in C++ by `../host/tss_host.cpp` (tests/host_shim_test.cpp).
//!
//! build.rs:  println!("cargo:rustc-link-lib=dylib=tss");
//!            println!("cargo:rustc-link-search=native={}", env!("TSS_LIB_DIR"));

#![allow(non_camel_case_types, dead_code)]

use crate::errors::{Result, SearchError};
use crate::DocRef;
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct tss_index { _p: [u8; 0] }
#[repr(C)]
pub struct tss_mask { _p: [u8; 0] }
#[repr(C)]
pub struct tss_terms { _p: [u8; 0] }
#[repr(C)]
pub struct tss_comm { _p: [u8; 0] }
#[repr(C)]
pub struct tss_columns { _p: [u8; 0] }

#[repr(C)]
#[derive(Default, Debug, Clone, Copy)]
pub struct tss_prefix_stats {
    pub exact_lo: u64,
    pub exact_hi: u64,
    pub sub_lo: u64,
    pub sub_hi: u64,
    pub npostings: u64,
}

pub const TSS_OK: c_int = 0;
pub const TSS_F32: c_int = 0;
pub const TSS_BF16: c_int = 1;
pub const TSS_MASK_NONE: c_int = 0;
pub const TSS_MASK_INCLUDE: c_int = 1;
pub const TSS_MASK_EXCLUDE: c_int = 2;
pub const TSS_PREFIX_TOKEN: c_int = 0;
pub const TSS_PREFIX_CHAR: c_int = 1;

#[link(name = "tss")]
extern "C" {
    pub fn tss_abi_version() -> c_int;
    pub fn tss_last_error() -> *const c_char;
    pub fn tss_device_count() -> c_int;

    pub fn tss_index_create(out: *mut *mut tss_index, dim: u32, storage: c_int, device: c_int) -> c_int;
    pub fn tss_index_reserve(ix: *mut tss_index, nrows: u64) -> c_int;
    pub fn tss_index_add(ix: *mut tss_index, rows: *const f32, nrows: u64) -> c_int;
    pub fn tss_index_add_synthetic(ix: *mut tss_index, row_begin: u64, nrows: u64, seed: u64) -> c_int;
    pub fn tss_index_finalize(ix: *mut tss_index) -> c_int;
    pub fn tss_index_size(ix: *const tss_index) -> u64;
    pub fn tss_index_dim(ix: *const tss_index) -> u32;
    pub fn tss_index_destroy(ix: *mut tss_index);
    pub fn tss_index_get_rows(ix: *mut tss_index, row_begin: u64, nrows: u64, out: *mut f32) -> c_int;
    pub fn tss_index_search(
        ix: *mut tss_index, queries: *const f32, nq: u32, k: u32, mask: *const tss_mask,
        mask_mode: c_int, out_rows: *mut u32, out_scores: *mut f32, out_counts: *mut u32,
    ) -> c_int;
    pub fn tss_index_search_device(
        ix: *mut tss_index, d_queries: *const f32, nq: u32, k: u32, mask: *const tss_mask,
        mask_mode: c_int, d_out_keys: *mut u64,
    ) -> c_int;
    pub fn tss_index_search_submit(
        ix: *mut tss_index, queries: *const f32, nq: u32, k: u32, mask: *const tss_mask,
        mask_mode: c_int, out_ticket: *mut u64,
    ) -> c_int;
    pub fn tss_index_search_collect(
        ix: *mut tss_index, ticket: u64, out_rows: *mut u32, out_scores: *mut f32,
        out_counts: *mut u32,
    ) -> c_int;
    pub fn tss_index_search_prefix(
        ix: *mut tss_index, t: *mut tss_terms, prefix: *const c_char, len: u32, kind: c_int,
        scratch: *mut tss_mask, queries: *const f32, nq: u32, k: u32, out_rows: *mut u32,
        out_scores: *mut f32, out_counts: *mut u32,
    ) -> c_int;
    pub fn tss_index_search_prefix_submit(
        ix: *mut tss_index, t: *mut tss_terms, prefix: *const c_char, len: u32, kind: c_int,
        scratch: *mut tss_mask, queries: *const f32, nq: u32, k: u32, out_ticket: *mut u64,
    ) -> c_int;
    pub fn tss_unpack_keys(keys: *const u64, n: u64, out_rows: *mut u32, out_scores: *mut f32);

    pub fn tss_comm_unique_id(out_id: *mut u8) -> c_int;
    pub fn tss_comm_create(out: *mut *mut tss_comm, id: *const u8, rank: c_int, nranks: c_int, device: c_int) -> c_int;
    pub fn tss_comm_destroy(c: *mut tss_comm);
    pub fn tss_index_set_shard(ix: *mut tss_index, row_base: u64, comm: *mut tss_comm) -> c_int;
    pub fn tss_index_set_batch_policy(ix: *mut tss_index, min_queries: u32, build_shadow_now: c_int) -> c_int;

    pub fn tss_mask_create(out: *mut *mut tss_mask, nbits: u64, device: c_int) -> c_int;
    pub fn tss_mask_clear(m: *mut tss_mask) -> c_int;
    pub fn tss_mask_set_rows(m: *mut tss_mask, rows: *const u32, n: u64, row_base: u64) -> c_int;
    pub fn tss_mask_upload(m: *mut tss_mask, words: *const u32) -> c_int;
    pub fn tss_mask_download(m: *const tss_mask, words: *mut u32) -> c_int;
    pub fn tss_mask_popcount(m: *const tss_mask, out: *mut u64) -> c_int;
    pub fn tss_mask_nbits(m: *const tss_mask) -> u64;
    pub fn tss_mask_destroy(m: *mut tss_mask);

    pub fn tss_terms_create(
        out: *mut *mut tss_terms, pool: *const c_char, term_off: *const u64, post_off: *const u64,
        post_rows: *const u32, nterms: u64, device: c_int,
    ) -> c_int;
    pub fn tss_terms_size(t: *const tss_terms) -> u64;
    pub fn tss_terms_destroy(t: *mut tss_terms);
    pub fn tss_prefix_mask(
        t: *mut tss_terms, prefix: *const c_char, len: u32, kind: c_int, out: *mut tss_mask,
        row_base: u64, stats: *mut tss_prefix_stats,
    ) -> c_int;
    pub fn tss_prefix_mask_fresh(
        t: *mut tss_terms, prefix: *const c_char, len: u32, kind: c_int, out: *mut tss_mask,
        row_base: u64, stats: *mut tss_prefix_stats,
    ) -> c_int;
    pub fn tss_terms_bind_stream(t: *mut tss_terms, ix: *mut tss_index) -> c_int;
    pub fn tss_terms_build_text(
        out: *mut *mut tss_terms, text: *const c_char, phrase_off: *const u64, rows: *const u32,
        n_phrases: u64, lowercase: c_int, max_tokens: u32, device: c_int,
    ) -> c_int;
    pub fn tss_terms_save(t: *const tss_terms, path: *const c_char) -> c_int;
    pub fn tss_terms_load(out: *mut *mut tss_terms, path: *const c_char, device: c_int) -> c_int;

    pub fn tss_terms_build(
        out: *mut *mut tss_terms, vocab_pool: *const c_char, vocab_off: *const u64, vocab_size: u32,
        token_ids: *const u32, max_tokens: u32, rows: *const u32, n_postings: u64, device: c_int,
    ) -> c_int;
    pub fn tss_terms_sizes(t: *const tss_terms, nterms: *mut u64, pool_bytes: *mut u64, npostings: *mut u64) -> c_int;
    pub fn tss_terms_export(
        t: *const tss_terms, pool: *mut c_char, term_off: *mut u64, post_off: *mut u64, post_rows: *mut u32,
    ) -> c_int;
    pub fn tss_index_save(ix: *mut tss_index, path: *const c_char) -> c_int;
    pub fn tss_index_load(out: *mut *mut tss_index, path: *const c_char, device: c_int) -> c_int;
    pub fn tss_mask_clear_rows(m: *mut tss_mask, rows: *const u32, n: u64, row_base: u64) -> c_int;
    pub fn tss_columns_create(
        out: *mut *mut tss_columns, court_ids: *const u16, dates: *const i32, nrows: u64, device: c_int,
    ) -> c_int;
    pub fn tss_columns_destroy(c: *mut tss_columns);
    pub fn tss_filter_mask(
        c: *mut tss_columns, allowed_courts: *const u16, n_allowed: u32, date_lo: i32, date_hi: i32,
        mask: *mut tss_mask, combine_and: c_int,
    ) -> c_int;

    pub fn tss_index_stream(ix: *mut tss_index) -> *mut c_void;
    pub fn tss_index_sync(ix: *mut tss_index) -> c_int;
    // measurement / diagnostics helpers (bench harnesses; not needed by vector.rs)
    pub fn tss_dev_alloc(device: c_int, bytes: u64, out: *mut *mut c_void) -> c_int;
    pub fn tss_dev_free(device: c_int, p: *mut c_void) -> c_int;
    pub fn tss_dev_h2d(device: c_int, dst: *mut c_void, src: *const c_void, bytes: u64) -> c_int;
    pub fn tss_dev_d2h(device: c_int, dst: *mut c_void, src: *const c_void, bytes: u64) -> c_int;
    pub fn tss_event_create(device: c_int, out: *mut *mut c_void) -> c_int;
    pub fn tss_event_record(ix: *mut tss_index, ev: *mut c_void) -> c_int;
    pub fn tss_event_elapsed_ms(ev_a: *mut c_void, ev_b: *mut c_void, out_ms: *mut f32) -> c_int;
    pub fn tss_event_destroy(ev: *mut c_void) -> c_int;
    pub fn tss_index_debug_phases(ix: *mut tss_index, d_buf: *mut c_void) -> c_int;
    pub fn tss_launch_count() -> u64;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(tss_last_error()).to_string_lossy().into_owned() }
}

/// Drop-in replacement for the stub `HnswIndex` of src/vector.rs:40-44.
pub struct HnswIndex {
    config: crate::config::HnswConfig,
    ix: *mut tss_index,
    dim: usize,
    /// row id -> DocRef; the GPU only ever sees dense u32 row ids
    doc_refs: Vec<DocRef>,
    /// rows staged on the host until the next search
    pending: Vec<f32>,
    dirty: bool,
}

// the handle is externally synchronised by the tokio RwLock around VectorIndex
// (src/search.rs:33-36,249-252)
unsafe impl Send for HnswIndex {}
unsafe impl Sync for HnswIndex {}

impl HnswIndex {
    /// src/vector.rs:185-188
    pub async fn new(config: crate::config::HnswConfig, dimension: usize) -> Result<Self> {
        let mut ix: *mut tss_index = std::ptr::null_mut();
        let rc = unsafe { tss_index_create(&mut ix, dimension as u32, TSS_F32, 0) };
        if rc != TSS_OK {
            return Err(SearchError::VectorIndexFailed { reason: last_error() });
        }
        unsafe { tss_index_reserve(ix, config.max_elements.min(1 << 20) as u64) };
        Ok(Self { config, ix, dim: dimension, doc_refs: Vec::new(), pending: Vec::new(), dirty: true })
    }

    /// src/vector.rs:190-193
    pub async fn add_vector(&mut self, doc_ref: DocRef, embedding: Vec<f32>) -> Result<()> {
        if embedding.len() != self.dim {
            return Err(SearchError::VectorIndexFailed {
                reason: format!("embedding has {} dims, index has {}", embedding.len(), self.dim),
            });
        }
        self.doc_refs.push(doc_ref);
        self.pending.extend_from_slice(&embedding);
        self.dirty = true;
        if self.pending.len() >= self.dim * 16384 {
            self.flush()?;
        }
        Ok(())
    }

    fn flush(&mut self) -> Result<()> {
        if !self.pending.is_empty() {
            let n = (self.pending.len() / self.dim) as u64;
            let rc = unsafe { tss_index_add(self.ix, self.pending.as_ptr(), n) };
            if rc != TSS_OK {
                return Err(SearchError::VectorIndexFailed { reason: last_error() });
            }
            self.pending.clear();
        }
        Ok(())
    }

    /// src/vector.rs:195-202 -- `&self` in the reference; the lazy flush needs `&mut`, which the
    /// only caller already has (VectorIndex::search takes `&mut self`, src/vector.rs:128-132).
    pub async fn search(&mut self, query_embedding: &[f32], top_k: usize) -> Result<Vec<(DocRef, f32)>> {
        if query_embedding.len() != self.dim {
            return Err(SearchError::HnswSearchError {
                details: format!("query has {} dims, index has {}", query_embedding.len(), self.dim),
            });
        }
        if top_k == 0 {
            return Ok(Vec::new());
        }
        if self.dirty {
            self.flush()?;
            if unsafe { tss_index_finalize(self.ix) } != TSS_OK {
                return Err(SearchError::VectorIndexFailed { reason: last_error() });
            }
            self.dirty = false;
        }
        let mut rows = vec![0u32; top_k];
        let mut scores = vec![0f32; top_k];
        let mut count = 0u32;
        let rc = unsafe {
            tss_index_search(
                self.ix, query_embedding.as_ptr(), 1, top_k as u32, std::ptr::null(), TSS_MASK_NONE,
                rows.as_mut_ptr(), scores.as_mut_ptr(), &mut count,
            )
        };
        if rc != TSS_OK {
            return Err(SearchError::HnswSearchError { details: last_error() });
        }
        // the ABI returns similarity; VectorIndex::search computes `1.0 - distance`
        // (src/vector.rs:144), so hand it a distance
        Ok((0..count as usize)
            .map(|i| (self.doc_refs[rows[i] as usize].clone(), 1.0 - scores[i]))
            .collect())
    }

    /// The same search split in two (no reference counterpart: `search` above never awaits).
    /// `submit` enqueues the scan and returns at once; up to `TSS_MAX_PENDING` (4) tickets may be
    /// outstanding per index, and `collect` waits for that ticket's scan alone -- so a task can
    /// keep two queries in flight and the device never idles between them.  The index must be
    /// flushed and finalized (call `search` once, or `flush` + `tss_index_finalize`).
    pub fn search_submit(&self, query_embedding: &[f32], top_k: usize) -> Result<(u64, usize)> {
        if query_embedding.len() != self.dim || top_k == 0 || top_k > 128 {
            return Err(SearchError::HnswSearchError {
                details: format!("submit: {} dims (index has {}), top_k {}", query_embedding.len(), self.dim, top_k),
            });
        }
        let mut ticket = 0u64;
        let rc = unsafe {
            tss_index_search_submit(
                self.ix, query_embedding.as_ptr(), 1, top_k as u32, std::ptr::null(), TSS_MASK_NONE, &mut ticket,
            )
        };
        if rc != TSS_OK {
            return Err(SearchError::HnswSearchError { details: last_error() });
        }
        Ok((ticket, top_k))
    }

    pub fn search_collect(&self, ticket: (u64, usize)) -> Result<Vec<(DocRef, f32)>> {
        let (ticket, top_k) = ticket;
        let mut rows = vec![0u32; top_k];
        let mut scores = vec![0f32; top_k];
        let mut count = 0u32;
        let rc = unsafe {
            tss_index_search_collect(self.ix, ticket, rows.as_mut_ptr(), scores.as_mut_ptr(), &mut count)
        };
        if rc != TSS_OK {
            return Err(SearchError::HnswSearchError { details: last_error() });
        }
        Ok((0..count as usize)
            .map(|i| (self.doc_refs[rows[i] as usize].clone(), 1.0 - scores[i]))
            .collect())
    }

    /// src/vector.rs:204-207
    pub fn size(&self) -> usize {
        self.doc_refs.len()
    }
}

impl Drop for HnswIndex {
    fn drop(&mut self) {
        unsafe { tss_index_destroy(self.ix) }
    }
}
