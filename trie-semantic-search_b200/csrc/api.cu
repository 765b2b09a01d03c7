// api.cu -- the extern "C" surface declared in include/tss.h.
//
// Host-side plumbing only: handles, HBM allocation, staging copies, launch
// dispatch, NCCL (loaded lazily with dlopen so a single-GPU user needs no
// NCCL at all).  No CPU fallback anywhere: every data-path call ends in a
// kernel launch or an error.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <atomic>
#include <condition_variable>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tss.h"
#include "aux_kernels.cuh"
#include "gemm_topk.cuh"
#include "terms_build.cuh"
#include "scan.cuh"
#include "scan_launch.h"

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  int code = (e == cudaErrorMemoryAllocation) ? TSS_ERR_OOM : TSS_ERR_CUDA;
  return fail(code, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}
#define CU(call)                                        \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

// ---- NCCL, resolved at first use -----------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;
int load_nccl() {
  if (g_nccl.ok) return TSS_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) return fail(TSS_ERR_NCCL, "dlopen(libnccl.so.2) failed: %s", dlerror());
#define SYM(field, name)                                                          \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.lib, name)); \
  if (!g_nccl.field) return fail(TSS_ERR_NCCL, "NCCL symbol %s missing", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllGather, "ncclAllGather")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.ok = true;
  return TSS_OK;
}
constexpr int kNcclUint64 = 5;

constexpr uint32_t kMaxBq = 4;           // queries per scan launch
constexpr uint32_t kScanSlots = 4;       // scan workspace slots (launches in flight under PDL)
constexpr uint32_t kWsQueries = 1024;    // device query/result workspace, in queries
constexpr uint32_t kPending = TSS_MAX_PENDING;  // searches in flight per handle (submit / collect)
constexpr uint32_t kPendingNq = TSS_PENDING_MAX_NQ;  // queries per pending search
constexpr size_t kStageBytes = 32u << 20;  // one of the two upload staging buffers
constexpr uint32_t kGemmCandCap = 32768;   // K2 survivor keys per query of a FULL workspace batch:
                                           // the pool (kWsQueries x this) is shared out per batch
constexpr uint32_t kGemmMaxSample = 8192;  // tiles sampled by the K2 threshold pass
constexpr uint32_t kMaskListCap = 131072;  // rows a prefix scatter may list for the list-driven scan
constexpr uint32_t kPrefilterMaxNq = 2;    // queries per call the shadow prefilter takes (K2 beyond)

typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                 const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmapEncodeFn g_tmap_encode = nullptr;
int load_tmap_encode() {
  if (g_tmap_encode) return TSS_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
    return fail(TSS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  g_tmap_encode = reinterpret_cast<TmapEncodeFn>(fn);
  return TSS_OK;
}
// 2-D bf16 tensor map over a row-major [rows][kpad] matrix, box = 64 elements x box_rows,
// 128-byte swizzle (what the UMMA K-major descriptors in gemm_topk.cu expect)
int make_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint32_t kpad, uint32_t box_rows) {
  cuuint64_t dims[2] = {kpad, rows ? rows : 1};
  cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_tmap_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                             strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TSS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return TSS_OK;
}

}  // namespace

struct tss_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1, device = 0;
};

// A mask is written and read by kernels on different streams (its own, a terms handle's, a
// columns handle's, an index's), all cudaStreamNonBlocking: nothing orders them implicitly.
// Every enqueue that touches d_words is therefore bracketed by mask_begin_write/mask_end_write or
// mask_begin_read/mask_end_read (below), which chain the streams with events: a write waits for
// the previous write and for every read enqueued since; a read waits for the last write.  No
// call relies on the legacy stream or on a host synchronisation for ordering.
struct tss_mask {
  int device = 0;
  uint64_t nbits = 0, nwords = 0;
  uint32_t* d_words = nullptr;
  unsigned long long* d_scratch = nullptr;
  cudaStream_t stream = nullptr;   // own stream: clear / upload / set_rows / download / popcount
  std::mutex mu;                   // guards the ordering state (concurrent readers are legal)
  bool has_write = false;
  cudaStream_t wstream = nullptr;  // stream the last write was enqueued on (identity only)
  cudaEvent_t wev = nullptr;       // recorded behind that write
  struct Reader {
    cudaStream_t s;
    cudaEvent_t ev;
  };
  std::vector<Reader> readers;     // reads enqueued since that write, one entry per stream
  std::vector<cudaEvent_t> spare;  // recycled reader events
  // Row list of the last tss_prefix_mask_fresh (see ScanParams::row_list): valid on the host
  // side while nothing else has touched the mask since; the device-side count says whether the
  // scatter actually produced one (few enough postings).
  uint32_t* d_list = nullptr;        // [kMaskListCap] unique local rows
  uint32_t* d_list_count = nullptr;
  bool list_valid = false;
};

struct tss_columns {
  int device = 0;
  uint64_t nrows = 0;
  uint16_t* d_court = nullptr;
  int32_t* d_date = nullptr;
  uint32_t* d_allow = nullptr;  // 65 536-bit allow set of the current call
  cudaStream_t stream = nullptr;
};

struct tss_terms {
  int device = 0;
  uint64_t nterms = 0, pool_bytes = 0, nposts = 0;
  char* d_pool = nullptr;
  uint64_t* d_term_off = nullptr;
  uint64_t* d_post_off = nullptr;
  uint32_t* d_post_rows = nullptr;
  char* d_keys = nullptr;     // probe key bytes
  char* h_keys = nullptr;     // pinned
  uint64_t* d_bounds = nullptr;  // 4 bounds + npostings, then two u32 grid-barrier words at [6]
  int num_sms = 0;
  uint32_t key_cap = 0;
  cudaStream_t stream = nullptr;      // the stream prefix searches are enqueued on
  cudaStream_t own_stream = nullptr;  // (stream == own_stream unless bound to an index's)
};

struct tss_index {
  // every entry point that touches the handle's stream or workspaces holds this: two host threads
  // may call into one handle (searches included); the calls are serialised inside, not racing
  std::mutex mu;
  int device = 0;
  uint32_t dim = 0;
  int storage = TSS_F32;
  int ns = 0;                 // storage stripes of 128 elements
  uint32_t stride_elems = 0;  // ns * 128
  size_t row_bytes = 0;
  uint8_t* d_rows = nullptr;
  uint64_t capacity = 0, n_rows = 0;
  bool finalized = false;
  cudaStream_t stream = nullptr;
  int num_sms = 0;
  uint64_t row_base = 0;
  tss_comm* comm = nullptr;
  // workspaces
  // upload pipeline (lazy): two pinned host buffers + two device staging buffers, so the host's
  // copy of chunk i+1 out of the caller's (pageable) memory overlaps the DMA + packing of chunk i
  float* d_stage = nullptr;        // 2 x kStageBytes
  float* h_stage = nullptr;        // 2 x kStageBytes, pinned
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  int* d_flag = nullptr;
  float* d_queries = nullptr;      // kWsQueries x dim
  uint64_t* d_keys = nullptr;      // kWsQueries x TSS_MAX_FUSED_K local results
  uint64_t* d_gather = nullptr;    // nranks x kWsQueries x k (lazy)
  uint64_t* d_merged = nullptr;    // kWsQueries x k (lazy)
  // Scan workspaces come in kScanSlots slots used round-robin, because consecutive scans
  // overlap under programmatic dependent launch: d_partials[slot][kMaxBq][num_sms][128] and
  // d_counter[slot*4 + {0 done ticket, 1 tile claims, 2 generation}]; d_counter[16] is the
  // exchange status word.
  uint64_t* d_partials = nullptr;
  unsigned int* d_counter = nullptr;
  unsigned int* d_walk_ctr = nullptr;  // [kScanSlots][kWalkCounters] claim counters of the masked
                                       // scan's walk, one per 128-byte line
  uint32_t launch_no = 0;
  bool no_host_sync = false;  // inside tss_index_search_device: fix-ups must stay on the device
  uint64_t shard_min_rows = 0;  // rows of the smallest shard of the group (tss_index_set_shard)
  bool pdl = true;
  // tile schedule of the unmasked scan (see scan.cuh): share of tiles walked statically,
  // tiles per dynamic claim, and how many warp-rounds at the very end are claimed one
  // tile at a time (tail balancing)
  float static_frac = 0.0f;
  uint32_t dyn_chunk = 8;
  uint32_t walk_run_log2 = 5;  // masked scan: a warp walks runs of 32 consecutive tiles (scan.cuh)
  float walk_static_frac = 0.5f;  // ... this share of them assigned statically, the rest claimed
  float fine_rounds = 2.0f;
  unsigned long long* d_dbg = nullptr;  // diagnostics (tss_index_debug_phases)
  float* h_queries = nullptr;  // pinned
  uint64_t* h_keys = nullptr;  // pinned + mapped: the scan writes its result straight into it
  unsigned int* h_status = nullptr;  // pinned + mapped exchange status word
  // Searches in flight (tss_index_search_submit / _collect, and every blocking search of <= 4
  // queries): each owns a pinned query slot, a device query slot, a pinned + mapped result slot
  // the scan's last CTA writes straight into, and an event.  The handle's lock is held while a
  // search is ENQUEUED, not while it is waited for, so the scans of several host threads (or of
  // one pipelining thread) queue up back to back on the device.
  struct Pending {
    cudaEvent_t ev = nullptr;
    uint32_t nq = 0, k = 0;
    uint32_t gen = 0;   // bumped by every submit: a stale or repeated ticket is caught
    int state = 0;      // 0 free, 1 in flight, 2 being collected
  } pend[kPending];
  uint32_t pend_next = 0;
  float* h_pq = nullptr;     // [kPending][kPendingNq][dim] pinned
  float* d_pq = nullptr;     // [kPending][kPendingNq][dim]
  uint64_t* h_pk = nullptr;  // [kPending][kPendingNq][TSS_MAX_FUSED_K] pinned + mapped
  std::condition_variable pend_cv;
  // K2 (tensor-core) workspace, allocated on first large-batch search of a bf16 index
  struct Gemm {
    bool ready = false;
    uint64_t norm_rows = 0;        // rows covered by d_inv_norm
    const void* norm_base = nullptr;
    float* d_inv_norm = nullptr;   // [capacity]
    uint64_t inv_norm_cap = 0;
    uint16_t* d_qbf16 = nullptr;   // [kWsQueries][kpad]
    float* d_inv_q = nullptr;      // [kWsQueries]
    float* d_margin = nullptr;     // [kWsQueries] rescoring margins
    float* d_thr = nullptr;        // [kWsQueries]
    float* d_tile_max = nullptr;   // [kGemmMaxSample][kWsQueries]
    uint64_t* d_cand = nullptr;    // [kWsQueries][kGemmCandCap]
    uint32_t* d_cand_count = nullptr;  // [kWsQueries][nslices]
    uint32_t* d_overflow = nullptr;    // [kWsQueries]
    uint64_t* d_pref_keys = nullptr;   // [kPrefilterMaxNq][128] shadow-scan candidates
    uint32_t* h_cand_count = nullptr;  // pinned copy of d_overflow
    // device-side fix-up of a batch (queries whose survivor list overflowed, or whose shadow
    // proof failed): [0] = count, [1..] = query indices; the mapped pinned mirror is read by
    // the host only where it synchronises anyway
    uint32_t* d_redo = nullptr;
    uint32_t* h_redo = nullptr;
    uint32_t fixups = 2;               // guarded fix-up scans enqueued behind every batch
                                       // (≈ 5 µs each when idle; doubled on demand up to 64)
    CUtensorMap tmap_q, tmap_e, tmap_e_half, tmap_e_quarter;  // corpus boxes of 256 / 128 (CTA
                                                              // pairs) / 64 rows (quads)
    uint64_t tmap_rows = 0;
    const void* tmap_base = nullptr;
    // The shadow: a bf16 copy of the matrix with every row scaled to UNIT length, which is what
    // the tensor cores read when it exists -- the accumulator then is the score and the epilogue
    // needs no per-row weight.  An fp32 index always gets one (+50 % memory; the survivors are
    // re-scored from the fp32 rows, so results stay bit-identical to the fp32 scan); a bf16 index
    // gets one (+100 %) when memory is plentiful or tss_index_set_batch_policy asks for it, and
    // otherwise the tensor cores read the stored rows and the epilogue weights by 1/|row|.
    uint8_t* d_shadow = nullptr;
    unsigned int* d_shadow_err = nullptr;  // float bits: max over rows of |e_n - e/|e|| (margins)
    int unit_policy = -1;          // bf16 index: -1 decide at the first batch, 0 stored rows, 1 shadow
    bool reading_shadow = false;   // what tmap_e / d_inv_norm currently describe
    uint64_t shadow_cap = 0, shadow_rows = 0;
    const void* shadow_base = nullptr;
    bool shadow_failed = false;    // no memory for it: large batches stay on the scan
    // Self-tuning of the survivor estimate: doubled (up to 64x) whenever more than 5 % of a
    // batch's queries overflowed their lists and had to be redone by the scan -- a corpus whose
    // scores crowd near the top (near-duplicates, tight clusters) then gets a larger tile
    // sample and smaller batches instead of one fallback scan per query.
    uint32_t spread_boost = 1;
  } gemm;
  // Batches at least this large use K2.  On a big corpus (where its five launches and one host
  // sync are noise) K2 also takes batches from gemm_small_nq queries when the bf16 matrix it
  // reads exists already: 3..15 queries cost one 1.3 ms pass over 10M bf16 rows instead of
  // ceil(nq/4) scans of 1.9 ms (bf16) / 2.3 ms (fp32).  Results are bit-identical either way.
  uint32_t gemm_min_nq = 16;
  uint32_t gemm_small_nq = 3;
  uint64_t gemm_small_rows = 2'000'000;
  uint32_t* d_round_mask = nullptr;  // scratch mask of the k > 128 scan rounds
  uint64_t round_mask_words = 0;
  // fused sharded merge: exchange buffers of all ranks mapped with CUDA IPC (<= 8 ranks)
  struct Xchg {
    bool ready = false;
    tss_comm* comm = nullptr;
    uint8_t* local = nullptr;
    uint8_t* peer[8] = {};
    uint32_t seq = 0;
    bool suppress = false;  // a rank-local redo (K2 overflow fallback) must not exchange
  } xchg;
};

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// ---- mask ordering (see struct tss_mask) ----------------------------------------------------
cudaError_t mask_begin_write(tss_mask* m, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(m->mu);
  m->list_valid = false;  // (tss_prefix_mask_fresh sets it again after its enqueue)
  cudaError_t e = cudaSuccess;
  if (m->has_write && m->wstream != s) e = cudaStreamWaitEvent(s, m->wev, 0);
  for (auto& r : m->readers) {
    if (e == cudaSuccess && r.s != s) e = cudaStreamWaitEvent(s, r.ev, 0);
    m->spare.push_back(r.ev);
  }
  m->readers.clear();
  return e;
}
cudaError_t mask_end_write(tss_mask* m, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(m->mu);
  m->has_write = true;
  m->wstream = s;
  return cudaEventRecord(m->wev, s);
}
cudaError_t mask_begin_read(tss_mask* m, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(m->mu);
  if (m->has_write && m->wstream != s) return cudaStreamWaitEvent(s, m->wev, 0);
  return cudaSuccess;
}
cudaError_t mask_end_read(tss_mask* m, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(m->mu);
  for (auto& r : m->readers)
    if (r.s == s) return cudaEventRecord(r.ev, s);
  cudaEvent_t ev = nullptr;
  if (!m->spare.empty()) {
    ev = m->spare.back();
    m->spare.pop_back();
  } else {
    cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  m->readers.push_back({s, ev});
  return cudaEventRecord(ev, s);
}
// brackets the enqueues of one search that reads `mask` on stream s
struct MaskReadScope {
  tss_mask* m;
  cudaStream_t s;
  cudaError_t err = cudaSuccess;
  MaskReadScope(const tss_mask* mask, int mode, cudaStream_t st)
      : m(mode != TSS_MASK_NONE ? const_cast<tss_mask*>(mask) : nullptr), s(st) {
    if (m) err = mask_begin_read(m, s);
  }
  ~MaskReadScope() {
    if (m) mask_end_read(m, s);
  }
};

// Host copy into a pinned staging buffer, split over a few threads: one core moves ~11 GB/s out
// of pageable memory, the link takes ~50.  (TSS_UPLOAD_THREADS overrides; 1 = plain memcpy.)
void staged_copy(void* dst, const void* src, size_t bytes) {
  static const unsigned nthreads = [] {
    unsigned n = std::thread::hardware_concurrency();
    n = n >= 16 ? 8 : n >= 8 ? 4 : n >= 4 ? 2 : 1;
    if (const char* env = getenv("TSS_UPLOAD_THREADS")) n = (unsigned)atoi(env);
    return n < 1 ? 1u : n > 8 ? 8u : n;
  }();
  if (nthreads == 1 || bytes < (8u << 20)) {
    memcpy(dst, src, bytes);
    return;
  }
  const size_t part = ((bytes / nthreads) + 4095) & ~(size_t)4095;
  std::thread th[8];
  unsigned started = 0;
  size_t off = part;  // this thread copies [0, part) itself
  for (; started + 1 < nthreads && off < bytes; ++started, off += part) {
    const size_t n = bytes - off < part ? bytes - off : part;
    char* d = static_cast<char*>(dst) + off;
    const char* sp = static_cast<const char*>(src) + off;
    try {
      th[started] = std::thread([d, sp, n] { memcpy(d, sp, n); });
    } catch (...) {  // no thread to be had: this one does the rest (nothing unwinds across the ABI)
      break;
    }
  }
  memcpy(dst, src, part < bytes ? part : bytes);
  if (off < bytes) memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, bytes - off);
  for (unsigned i = 0; i < started; ++i) th[i].join();
}

int ensure_capacity(tss_index* ix, uint64_t need) {
  if (need <= ix->capacity) return TSS_OK;
  uint64_t cap = ix->capacity ? ix->capacity + ix->capacity / 2 : 0;
  if (cap < need) cap = need;
  // one R-row tile of slack so the last bulk copy never leaves the allocation
  size_t bytes = (size_t)(cap + 16) * ix->row_bytes;
  uint8_t* nd = nullptr;
  cudaError_t e = cudaMalloc(&nd, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(TSS_ERR_OOM, "cudaMalloc(%zu bytes for %llu rows) failed: %s", bytes,
                (unsigned long long)cap, cudaGetErrorString(e));
  }
  if (ix->n_rows) {
    CU(cudaMemcpyAsync(nd, ix->d_rows, (size_t)ix->n_rows * ix->row_bytes, cudaMemcpyDeviceToDevice,
                       ix->stream));
    CU(cudaStreamSynchronize(ix->stream));
  }
  if (ix->d_rows) cudaFree(ix->d_rows);
  ix->d_rows = nd;
  ix->capacity = cap;
  return TSS_OK;
}

int check_mask(const tss_index* ix, const tss_mask* mask, int mode) {
  if (mode != TSS_MASK_NONE && mode != TSS_MASK_INCLUDE && mode != TSS_MASK_EXCLUDE)
    return fail(TSS_ERR_INVALID_ARG, "mask_mode %d is not a TSS_MASK_* value", mode);
  if (mode != TSS_MASK_NONE) {
    if (!mask) return fail(TSS_ERR_INVALID_ARG, "mask_mode set but mask is NULL");
    if (mask->device != ix->device) return fail(TSS_ERR_INVALID_ARG, "mask lives on another device");
    if (mask->nbits < ix->n_rows)
      return fail(TSS_ERR_INVALID_ARG, "mask has %llu bits, shard has %llu rows",
                  (unsigned long long)mask->nbits, (unsigned long long)ix->n_rows);
  }
  return TSS_OK;
}

// enqueue the scan for nq device-resident queries -> d_out (nq x k local keys).  A single
// query of <= 384 dims may instead be handed over from host memory inside the kernel parameters.
// bf16_rows: scan this bf16 matrix of the same geometry (the shadow of an fp32 index) instead
// of the stored rows.
// guard != null: ONE guarded launch (see ScanParams::guard_list): d_queries / d_out are the
// bases of the batch, the query it serves (if any) is picked on the device.
struct ScanGuard {
  const uint32_t* list;
  const uint32_t* count;
  uint32_t index;
};
int enqueue_scan(tss_index* ix, const float* d_queries, uint32_t nq, uint32_t k,
                 const tss_mask* mask, int mode, uint64_t* d_out, const float* h_inline_query,
                 const uint8_t* bf16_rows, const ScanGuard* guard) {
  const bool as_bf16 = bf16_rows || ix->storage == TSS_BF16;
  const size_t row_bytes = bf16_rows ? (size_t)ix->stride_elems * 2 : ix->row_bytes;
  const uint32_t kp = tss::kp_for_k(k), cap = tss::cap_for_k(k);
  const uint32_t bq_max = (uint32_t)tss::max_bq_for_k(k);
  for (uint32_t q0 = 0; q0 < nq;) {
    uint32_t left = nq - q0;
    uint32_t take = left < bq_max ? left : bq_max;
    int bq = take >= 3 ? 4 : (int)take;  // kernel instances: 1, 2, 4
    tss::ScanParams p{};
    p.rows = bf16_rows ? bf16_rows : ix->d_rows;
    p.n_rows = ix->n_rows;
    p.row_base = (uint32_t)ix->row_base;
    p.dim = ix->dim;
    p.queries = d_queries + (size_t)q0 * ix->dim;
    if (h_inline_query && nq == 1 && ix->dim <= 384) {
      memcpy(p.q_inline, h_inline_query, ix->dim * sizeof(float));
      p.use_inline = 1;
    }
    p.nq_valid = take;
    p.k = k;
    p.kp = kp;
    p.cap = cap;
    p.mask = mode != TSS_MASK_NONE ? mask->d_words : nullptr;
    p.mask_mode = mode;
    if (guard) {
      p.guard_list = guard->list;
      p.guard_count = guard->count;
      p.guard_index = guard->index;
    }
    if (mode == TSS_MASK_INCLUDE && mask->list_valid && mask->d_list) {
      p.row_list = mask->d_list;
      p.row_list_count = mask->d_list_count;
      p.row_list_cap = kMaskListCap;
    }
    const uint32_t no = ++ix->launch_no;  // 1, 2, ...
    const uint32_t slot = no % kScanSlots;
    p.partials = ix->d_partials + (size_t)slot * kMaxBq * ix->num_sms * 128;
    p.done_counter = ix->d_counter + slot * 4;
    p.tile_counter = ix->d_counter + slot * 4 + 1;
    p.slot_gen = ix->d_counter + slot * 4 + 2;
    p.walk_counters = ix->d_walk_ctr + (size_t)slot * tss::kWalkCounters * 32;
    p.launch_no = no;
    p.expect_gen = no > kScanSlots ? no - kScanSlots : 0;
    p.pdl = ix->pdl ? 1 : 0;
    {
      // tiles per warp-round = num_sms * 16 warps; rows per tile from the storage geometry
      const uint64_t tile_rows = 12288 / row_bytes >= 16  ? 16
                                 : 12288 / row_bytes >= 8 ? 8
                                 : 12288 / row_bytes >= 4 ? 4
                                 : 12288 / row_bytes >= 2 ? 2
                                                          : 1;
      const uint64_t tiles = (ix->n_rows + tile_rows - 1) / tile_rows;
      const uint64_t gw = (uint64_t)ix->num_sms * 16;
      p.static_rounds = (uint32_t)((double)(tiles / gw) * ix->static_frac);
      p.dyn_chunk = ix->dyn_chunk;
      p.walk_run_log2 = ix->walk_run_log2;
      {
        const uint32_t wl = ix->walk_run_log2 & 7u;
        const uint64_t runs = (tiles + ((1ull << wl) - 1)) >> wl;
        p.walk_static_rounds = (uint32_t)((double)(runs / gw) * ix->walk_static_frac);
      }
      const uint64_t fine_tiles = (uint64_t)((double)gw * ix->fine_rounds);
      p.fine_start = tiles > fine_tiles ? tiles - fine_tiles : 0;
    }
    p.out_keys = d_out + (size_t)q0 * k;
    p.dbg = ix->d_dbg;
    if (ix->comm && ix->xchg.ready && !ix->xchg.suppress) {
      for (int r = 0; r < ix->comm->nranks; ++r) p.xchg_peer[r] = ix->xchg.peer[r];
      p.xchg_nranks = (uint32_t)ix->comm->nranks;
      p.xchg_rank = (uint32_t)ix->comm->rank;
      p.xchg_seq = ++ix->xchg.seq;
      p.xchg_status = ix->h_status;  // mapped pinned: the host reads it without a copy
      p.xchg_turn = ix->d_counter + 17;
    }
    cudaError_t e = tss::launch_scan(ix->ns, p, bq, as_bf16, mode != TSS_MASK_NONE,
                                     ix->num_sms, ix->device, ix->stream);
    if (e != cudaSuccess) {
      // nothing will ever publish slot_gen / xchg_turn for this number: give it back, or the
      // launch that reuses the slot (the next exchange) would wait for it forever
      --ix->launch_no;
      if (p.xchg_nranks) --ix->xchg.seq;
      return cuda_fail(e, "scan_topk_kernel launch");
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    q0 += take;
  }
  return TSS_OK;
}

// sharded tail: all-gather local keys, merge.  d_local: nq x k.  result in d_merged.
int enqueue_gather_merge(tss_index* ix, const uint64_t* d_local, uint32_t nq, uint32_t k,
                         uint64_t* d_merged_out) {
  tss_comm* c = ix->comm;
  size_t per_rank = (size_t)nq * k;
  int rc = g_nccl.AllGather(d_local, ix->d_gather, per_rank, kNcclUint64, c->comm, ix->stream);
  if (rc != 0) return fail(TSS_ERR_NCCL, "ncclAllGather: %s", g_nccl.GetErrorString(rc));
  cudaError_t e = tss::launch_merge_gathered(ix->d_gather, d_merged_out, (uint32_t)c->nranks, nq, k,
                                             ix->stream);
  if (e != cudaSuccess) return cuda_fail(e, "merge_gathered_kernel launch");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return TSS_OK;
}

int ensure_gather_ws(tss_index* ix) {
  if (!ix->comm || ix->d_gather) return TSS_OK;
  CU(cudaMalloc(&ix->d_gather, (size_t)ix->comm->nranks * kWsQueries * TSS_MAX_K * sizeof(uint64_t)));
  CU(cudaMalloc(&ix->d_merged, (size_t)kWsQueries * TSS_MAX_K * sizeof(uint64_t)));
  return TSS_OK;
}

bool gemm_eligible(const tss_index* ix, uint32_t nq, uint32_t k, int mode) {
  (void)mode;  // masks ride along: a masked row's 1/|row| is NaN in the epilogue
  if (ix->storage == TSS_F32 && ix->gemm.shadow_failed && !ix->comm) return false;
  // (every rank of a sharded index must take the same path, or the exchange deadlocks: the
  // rank-local conditions -- shard size, shadow built -- only count on an unsharded index)
  const bool have_bf16 = ix->storage == TSS_BF16 || ix->gemm.shadow_rows == ix->n_rows;
  const bool small_ok = !ix->comm && have_bf16 && nq >= ix->gemm_small_nq &&
                        ix->n_rows >= ix->gemm_small_rows;
  // (sharded: the smallest shard of the group decides, so every rank routes alike)
  const uint64_t rows = ix->comm ? ix->shard_min_rows : ix->n_rows;
  return (nq >= ix->gemm_min_nq || small_ok) && rows >= 4ull * 256 * k && k <= TSS_MAX_K;
}

// survivors a query is expected to leave in the K2 lists when `sample` tiles set its threshold
uint64_t gemm_expected_survivors(const tss_index* ix, uint32_t k, uint32_t num_tiles, uint32_t sample) {
  // e^(z * margin / sigma) with the margins of prep_queries_kernel, z ~ 4.3 at k = 100 of 10M
  const uint64_t spread = ix->storage == TSS_F32 || ix->gemm.reading_shadow ? 5 : 3;
  return spread * ix->gemm.spread_boost * k * (uint64_t)num_tiles / (sample ? sample : 1);
}
// largest batch (a multiple of 256 queries, <= kWsQueries) whose queries' expected survivors fit
// half their share of the pool with the largest tile sample the threshold pass can take
uint32_t gemm_batch_limit(const tss_index* ix, uint32_t k) {
  const uint32_t num_tiles = (uint32_t)((ix->n_rows + 255) / 256);
  uint32_t max_sample = kGemmMaxSample / (uint32_t)tss::gemm_col_split();
  if (max_sample > num_tiles) max_sample = num_tiles;
  const uint64_t need = gemm_expected_survivors(ix, k, num_tiles, max_sample) * 2;
  uint64_t lim = (uint64_t)kWsQueries * kGemmCandCap / (need ? need : 1);
  lim = lim / 256 * 256;
  if (lim < 256) lim = 256;
  return lim > kWsQueries ? kWsQueries : (uint32_t)lim;
}

// want_shadow: (bf16 index) read the unit-row shadow rather than the stored rows.  An fp32 index
// always reads its shadow.
int ensure_gemm_ws(tss_index* ix, bool want_shadow) {
  tss_index::Gemm& g = ix->gemm;
  int rc = load_tmap_encode();
  if (rc) return rc;
  const uint32_t kpad = ix->stride_elems;
  if (!g.ready) {
    CU(cudaMalloc(&g.d_qbf16, (size_t)kWsQueries * kpad * 2));
    CU(cudaMalloc(&g.d_inv_q, kWsQueries * sizeof(float)));
    CU(cudaMalloc(&g.d_margin, kWsQueries * sizeof(float)));
    CU(cudaMalloc(&g.d_thr, kWsQueries * sizeof(float)));
    CU(cudaMalloc(&g.d_tile_max, (size_t)kGemmMaxSample * kWsQueries * sizeof(float)));
    CU(cudaMalloc(&g.d_cand, (size_t)kWsQueries * kGemmCandCap * sizeof(uint64_t)));
    CU(cudaMalloc(&g.d_cand_count, (size_t)kWsQueries * 1024 * sizeof(uint32_t)));
    CU(cudaMalloc(&g.d_overflow, kWsQueries * sizeof(uint32_t)));
    CU(cudaMalloc(&g.d_pref_keys, (size_t)kPrefilterMaxNq * 128 * sizeof(uint64_t)));
    CU(cudaMallocHost(&g.h_cand_count, kWsQueries * sizeof(uint32_t)));
    CU(cudaMalloc(&g.d_redo, (kWsQueries + 1) * sizeof(uint32_t)));
    CU(cudaMallocHost(&g.h_redo, (kWsQueries + 1) * sizeof(uint32_t)));
    memset(g.h_redo, 0, (kWsQueries + 1) * sizeof(uint32_t));
    if ((rc = make_tmap(&g.tmap_q, g.d_qbf16, kWsQueries, kpad, 128))) return rc;
    g.ready = true;
  }
  // the bf16 matrix the tensor cores read: the unit-row shadow, or a bf16 index's own rows
  const uint8_t* e_rows = ix->d_rows;
  if (ix->storage == TSS_F32 || want_shadow) {
    if (g.shadow_cap < ix->n_rows) {
      cudaFree(g.d_shadow);
      g.d_shadow = nullptr;
      g.shadow_cap = g.shadow_rows = 0;
      const size_t bytes = (size_t)(ix->capacity + 16) * kpad * 2;
      if (cudaMalloc(&g.d_shadow, bytes) != cudaSuccess) {
        cudaGetLastError();
        g.shadow_failed = true;
        return fail(TSS_ERR_OOM, "no memory for the %zu-byte bf16 shadow of the index", bytes);
      }
      g.shadow_cap = ix->capacity;
    }
    if (g.shadow_rows != ix->n_rows || g.shadow_base != ix->d_rows) {
      // (the stored rows are already padded to the stride)
      if (!g.d_shadow_err) CU(cudaMalloc(&g.d_shadow_err, sizeof(unsigned int)));
      cudaError_t e = tss::launch_normalize_rows(ix->d_rows, ix->storage == TSS_BF16, g.d_shadow,
                                                 ix->n_rows, kpad, g.d_shadow_err, ix->stream);
      if (e != cudaSuccess) return cuda_fail(e, "unit-row shadow launch");
      g_launches.fetch_add(1, std::memory_order_relaxed);
      g.shadow_rows = ix->n_rows;
      g.shadow_base = ix->d_rows;
    }
    e_rows = g.d_shadow;
  }
  g.reading_shadow = e_rows == g.d_shadow;
  if (g.inv_norm_cap < ix->n_rows) {
    cudaFree(g.d_inv_norm);
    g.d_inv_norm = nullptr;
    CU(cudaMalloc(&g.d_inv_norm, (size_t)ix->capacity * sizeof(float)));
    g.inv_norm_cap = ix->capacity;
    g.norm_rows = 0;
  }
  if (g.norm_rows != ix->n_rows || g.norm_base != e_rows) {
    cudaError_t e =
        tss::launch_row_inv_norm(e_rows, ix->n_rows, ix->stride_elems, g.d_inv_norm, ix->stream);
    if (e != cudaSuccess) return cuda_fail(e, "row_inv_norm launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    g.norm_rows = ix->n_rows;
    g.norm_base = e_rows;
  }
  if (g.tmap_rows != ix->n_rows || g.tmap_base != e_rows) {
    if ((rc = make_tmap(&g.tmap_e, e_rows, ix->n_rows, kpad, 256))) return rc;
    if ((rc = make_tmap(&g.tmap_e_half, e_rows, ix->n_rows, kpad, 128))) return rc;
    if ((rc = make_tmap(&g.tmap_e_quarter, e_rows, ix->n_rows, kpad, 64))) return rc;
    g.tmap_rows = ix->n_rows;
    g.tmap_base = e_rows;
  }
  return TSS_OK;
}

// bf16 index: should the tensor cores read a unit-row shadow?  Decided once, at the first batch:
// yes when the copy leaves at least as much memory free again (TSS_GEMM_UNIT_SHADOW=0/1 forces).
bool unit_shadow_wanted(tss_index* ix) {
  tss_index::Gemm& g = ix->gemm;
  if (ix->storage != TSS_BF16) return true;
  if (g.unit_policy < 0) {
    g.unit_policy = 0;
    if (const char* env = getenv("TSS_GEMM_UNIT_SHADOW")) {
      g.unit_policy = atoi(env) != 0;
    } else {
      size_t free_b = 0, total_b = 0;
      const size_t bytes = (size_t)(ix->capacity + 16) * ix->stride_elems * 2;
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b > 2 * bytes + (4ull << 30))
        g.unit_policy = 1;
    }
  }
  if (g.unit_policy == 1 && g.shadow_failed) return false;
  return g.unit_policy == 1;
}

// gemm_eligible, and (fp32 index) the bf16 shadow could be built.  Without room for the shadow the
// batch stays on the scan: exact all the same, one launch per 4 queries.
bool gemm_route(tss_index* ix, uint32_t nq, uint32_t k, int mode) {
  if (!gemm_eligible(ix, nq, k, mode)) return false;
  // a sharded index reports the failure instead (enqueue_gemm returns it): quietly scanning on
  // one rank while the others run K2 would desynchronise the exchange
  if (ix->storage == TSS_F32 && !ix->comm && ensure_gemm_ws(ix, true) == TSS_ERR_OOM &&
      ix->gemm.shadow_failed)
    return false;
  return true;
}

struct ScanGuard;
int enqueue_scan(tss_index* ix, const float* d_queries, uint32_t nq, uint32_t k,
                 const tss_mask* mask, int mode, uint64_t* d_out,
                 const float* h_inline_query = nullptr, const uint8_t* bf16_rows = nullptr,
                 const ScanGuard* guard = nullptr);
int enqueue_fixups(tss_index* ix, const float* d_queries, uint32_t nq, uint32_t k,
                   const tss_mask* mask, int mode, uint64_t* d_out);
int enqueue_scan_rounds(tss_index* ix, const float* d_queries, uint32_t nq, uint32_t k,
                        const tss_mask* mask, int mode, uint64_t* d_out);

// K2: nq <= kWsQueries device-resident fp32 queries -> d_out (nq x k local keys).  Queries whose
// survivor list overflowed (adversarially clustered scores) are redone exactly by the K1 scan:
// see enqueue_fixups for who does that and when (the host only where it synchronises anyway).
int enqueue_gemm(tss_index* ix, const float* d_queries, uint32_t nq, uint32_t k,
                 const tss_mask* mask, int mode, uint64_t* d_out) {
  // survivors are re-scored with the scan's arithmetic unless TSS_GEMM_RESCORE=0 (then the
  // result is the top-k of the bf16 x bf16 tensor-core scores over the STORED bf16 rows)
  bool rescore = true;
  if (const char* rsc = getenv("TSS_GEMM_RESCORE")) rescore = atoi(rsc) != 0;
  if (ix->storage == TSS_F32) rescore = true;  // the raw scores would be those of the shadow
  tss_index::Gemm& g = ix->gemm;
  int rc = ensure_gemm_ws(ix, rescore && unit_shadow_wanted(ix));
  if (rc == TSS_ERR_OOM && ix->storage == TSS_BF16 && g.shadow_failed)
    rc = ensure_gemm_ws(ix, false);  // no room for the unit copy: weight the stored rows instead
  if (rc) return rc;
  const uint32_t kpad = ix->stride_elems;
  // CTA pairs (one cta_group::2 MMA over two query blocks): the batch is padded to an even
  // number of 128-query blocks (a padding query is all zeros and its threshold is +inf)
  // Up to 128 queries every SM streams its own tiles and independent CTAs are faster.
  int cluster = nq > 128 ? tss::TSS_GEMM_PAIR : tss::TSS_GEMM_SINGLE;
  // Quads (two pairs sharing corpus tiles by multicast) halve the L2 -> SM feed, the resource
  // the pair kernel runs out of; they need a multiple of four query blocks (no extra padding
  // beyond the pairs') and enough resident clusters for at least as many SMs as ... see below
  const int kb = (int)(ix->stride_elems / 64);
  int max_quads = 0;
  if (nq > 256 && ((nq + 255) / 256) % 2 == 0) {
    max_quads = tss::gemm_max_quads(kb);
    const int groups4 = (int)((nq + 511) / 512);
    if (max_quads >= groups4) cluster = tss::TSS_GEMM_QUAD;
  }
  if (const char* cl = getenv("TSS_GEMM_CLUSTER")) {  // diagnostics: 1 independent CTAs, 2 pairs, 4 quads
    const int want = atoi(cl);
    if (want == tss::TSS_GEMM_SINGLE) cluster = tss::TSS_GEMM_SINGLE;
    else if (want == tss::TSS_GEMM_PAIR || cluster != tss::TSS_GEMM_QUAD) cluster = tss::TSS_GEMM_PAIR;
  }
  const uint32_t qblock = cluster == tss::TSS_GEMM_QUAD ? 512 : cluster == tss::TSS_GEMM_PAIR ? 256 : 128;
  const uint32_t nq_pad = (nq + qblock - 1) / qblock * qblock, mb = nq_pad / 128;
  const uint32_t num_tiles = (uint32_t)((ix->n_rows + 255) / 256);
  // the threshold pass yields `split` maxima per sampled tile (one per column part)
  const uint32_t split = (uint32_t)tss::gemm_col_split();
  uint32_t sample = k * 8 / split > 1024 ? k * 8 / split : 1024;
  // The survivor pool is shared out among the queries of this batch.  About
  // spread * k * num_tiles / sample keys of a query pass its threshold (spread: the rescoring
  // margin lets in more); sample enough tiles for that to stay under half a query's share.
  const uint64_t share = (uint64_t)kWsQueries * kGemmCandCap / nq_pad;
  const uint64_t want = gemm_expected_survivors(ix, k, num_tiles, 1) * 2 / share + 1;
  if (want > sample) sample = (uint32_t)(want < num_tiles ? want : num_tiles);
  if (const char* sm = getenv("TSS_GEMM_SAMPLE")) sample = (uint32_t)atoi(sm);
  if (sample < (k + split - 1) / split) sample = (k + split - 1) / split;
  if (sample > kGemmMaxSample / split) sample = kGemmMaxSample / split;
  if (sample > num_tiles) sample = num_tiles;
  // every cluster the device can hold is launched; the kernel spreads the (query group, tile)
  // work over all of them (gemm_topk.cu "work assignment"), so no SM idles because the SM count
  // is not a multiple of the query blocks.  Lists per query: one per (slice of its group, part).
  const int cl = cluster == tss::TSS_GEMM_QUAD ? 4 : cluster == tss::TSS_GEMM_PAIR ? 2 : 1;
  int nclusters = cluster == tss::TSS_GEMM_QUAD ? max_quads : ix->num_sms / cl;
  if (const char* nc = getenv("TSS_GEMM_CLUSTERS"))  // diagnostics: fewer clusters
    if (atoi(nc) > 0 && atoi(nc) < nclusters) nclusters = atoi(nc);
  const int ngroups = (int)mb / cl;
  if (nclusters < ngroups) return fail(TSS_ERR_INVALID_ARG, "batch of %u queries exceeds one K2 launch", nq);
  int nslices = nclusters / ngroups + nclusters % ngroups;
  while (nslices * (int)tss::gemm_col_split() > 1024) {  // (select_kernel's list limit)
    --nclusters;
    nslices = nclusters / ngroups + nclusters % ngroups;
  }
  const int grid = nclusters * cl;
  const CUtensorMap& tmap_e = cluster == tss::TSS_GEMM_SINGLE ? g.tmap_e
                              : cluster == tss::TSS_GEMM_PAIR ? g.tmap_e_half
                                                              : g.tmap_e_quarter;
  cudaError_t e;
  e = tss::launch_prep_queries(d_queries, nq, ix->dim, kpad, nq_pad, g.d_qbf16, g.d_inv_q,
                               g.d_margin, g.reading_shadow ? g.d_shadow_err : nullptr, ix->stream);
  if (e != cudaSuccess) return cuda_fail(e, "prep_queries launch");
  const uint32_t nsub = (uint32_t)nslices * split;
  const uint64_t cap64 = share / nsub;
  const uint32_t cap_s = cap64 > 4096 ? 4096u : (uint32_t)cap64;
  tss::GemmParams p{};
  p.n_rows = ix->n_rows;
  p.row_base = (uint32_t)ix->row_base;
  // unit rows and no mask: the accumulator is the score, the epilogue applies no weight
  p.inv_norm = g.reading_shadow && mode == TSS_MASK_NONE ? nullptr : g.d_inv_norm;
  if (const char* wt = getenv("TSS_GEMM_WEIGHTS"))  // diagnostics: 1 = always weight
    if (atoi(wt) != 0) p.inv_norm = g.d_inv_norm;
  p.mask = mode != TSS_MASK_NONE ? mask->d_words : nullptr;
  p.mask_mode = mode;
  p.mb = mb;
  p.num_tiles = num_tiles;
  p.sample_stride = num_tiles / sample;
  p.sample_count = sample;
  p.tile_max = g.d_tile_max;
  p.thr = g.d_thr;
  p.cand = g.d_cand;
  p.cand_count = g.d_cand_count;
  p.cand_cap = cap_s;
  if (const char* rs = getenv("TSS_GEMM_STAGES")) p.ring_stages = (uint32_t)atoi(rs);
  if (const char* dbg = getenv("TSS_GEMM_DEBUG")) p.debug = (uint32_t)atoi(dbg);
  p.mode = 0;
  if ((e = tss::launch_gemm_topk(kb, cluster, g.tmap_q, tmap_e, p, grid, ix->stream)) != cudaSuccess)
    return cuda_fail(e, "gemm_topk_kernel (threshold pass) launch");
  if ((e = tss::launch_threshold(g.d_tile_max, sample * split, nq_pad, nq, k,
                                  rescore ? g.d_margin : nullptr, g.d_thr, ix->stream)) != cudaSuccess)
    return cuda_fail(e, "threshold_kernel launch");
  p.mode = 1;
  if ((e = tss::launch_gemm_topk(kb, cluster, g.tmap_q, tmap_e, p, grid, ix->stream)) != cudaSuccess)
    return cuda_fail(e, "gemm_topk_kernel (collect pass) launch");
  if ((e = tss::launch_select(g.d_cand, g.d_cand_count, nsub, cap_s, g.d_inv_q, nq, k,
                              rescore ? d_queries : nullptr, g.d_margin, ix->d_rows,
                              ix->storage == TSS_F32, ix->dim, kpad, (uint32_t)ix->row_base,
                              ix->n_rows, d_out, g.d_overflow, ix->stream)) != cudaSuccess)
    return cuda_fail(e, "select_kernel launch");
  g_launches.fetch_add(5, std::memory_order_relaxed);
  return enqueue_fixups(ix, d_queries, nq, k, mask, mode, d_out);
}

// Queries flagged in g.d_overflow (a K2 survivor list overflowed, or the shadow prefilter could
// not prove completeness) are redone exactly by the scan.  On the device: the flags are compacted
// into a list and `fixups` GUARDED scans are enqueued behind the batch -- launch j serves the
// j-th flagged query if there is one, else it is a few microseconds of nothing -- so
// tss_index_search_device needs no host synchronisation; more flagged queries than launches
// leave a sticky status for tss_index_sync and the number of fix-up launches doubles.
// tss_index_search, which synchronises anyway, reads the list and redoes them from the host.
// k > 128 (scan by rounds) is redone from the host in either entry.
int enqueue_fixups(tss_index* ix, const float* d_queries, uint32_t nq, uint32_t k,
                   const tss_mask* mask, int mode, uint64_t* d_out) {
  tss_index::Gemm& g = ix->gemm;
  const bool on_device = k <= TSS_MAX_FUSED_K;
  if (on_device && ix->no_host_sync) {
    // nobody waits for this batch's flags: tune on what an earlier batch needed (its mirror has
    // landed in mapped host memory by now, or will be seen by a later call)
    const uint32_t seen = g.h_redo[0];
    g.h_redo[0] = 0;
    if (seen > g.fixups && g.fixups < 64) g.fixups *= 2;
    if (seen * 20 > nq && g.spread_boost < 64) g.spread_boost *= 2;
  }
  // (tss_index_search synchronises anyway and redoes flagged queries from the host: it does not
  // pay for guarded launches that almost never have work)
  const uint32_t fixups = on_device && ix->no_host_sync ? (g.fixups < nq ? g.fixups : nq) : 0;
  cudaError_t e = tss::launch_redo_compact(g.d_overflow, nq, g.d_redo + 1, g.d_redo, g.h_redo + 1,
                                           g.h_redo, on_device && ix->no_host_sync ? fixups : nq,
                                           ix->h_status + 1, ix->stream);
  if (e != cudaSuccess) return cuda_fail(e, "redo_compact launch");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  int rc = TSS_OK;
  const bool saved = ix->xchg.suppress;
  ix->xchg.suppress = true;  // a redo is rank-local; the caller merges across ranks
  for (uint32_t j = 0; j < fixups && !rc; ++j) {
    ScanGuard guard{g.d_redo + 1, g.d_redo, j};
    rc = enqueue_scan(ix, d_queries, 1, k, mask, mode, d_out, nullptr, nullptr, &guard);
  }
  ix->xchg.suppress = saved;
  if (rc) return rc;
  if (on_device && ix->no_host_sync) return TSS_OK;
  CU(cudaStreamSynchronize(ix->stream));
  const uint32_t count = g.h_redo[0];
  if (count * 20 > nq && g.spread_boost < 64) g.spread_boost *= 2;
  ix->xchg.suppress = true;
  for (uint32_t j = fixups; j < count && !rc; ++j) {
    const uint32_t qi = g.h_redo[1 + j];
    if (k > TSS_MAX_FUSED_K)
      rc = enqueue_scan_rounds(ix, d_queries + (size_t)qi * ix->dim, 1, k, mask, mode,
                               d_out + (size_t)qi * k);
    else
      rc = enqueue_scan(ix, d_queries + (size_t)qi * ix->dim, 1, k, mask, mode,
                        d_out + (size_t)qi * k);
  }
  ix->xchg.suppress = saved;
  return rc;
}

// k > TSS_MAX_FUSED_K on the scan path: rounds of 128, each excluding what the previous rounds
// found (exact; ceil(k/128) scans per query, one query at a time because the exclusions differ).
int enqueue_scan_rounds(tss_index* ix, const float* d_queries, uint32_t nq, uint32_t k,
                        const tss_mask* mask, int mode, uint64_t* d_out) {
  const uint64_t words = (ix->n_rows + 31) / 32 + 1;
  if (ix->round_mask_words < words) {
    cudaFree(ix->d_round_mask);
    ix->d_round_mask = nullptr;
    CU(cudaMalloc(&ix->d_round_mask, words * sizeof(uint32_t)));
    ix->round_mask_words = words;
  }
  tss_mask scratch;
  scratch.device = ix->device;
  scratch.nbits = ix->n_rows;
  scratch.nwords = words - 1;
  scratch.d_words = ix->d_round_mask;
  // NONE -> exclude what was found; EXCLUDE -> same on a copy; INCLUDE -> clear found bits of a copy
  const int round_mode = mode == TSS_MASK_INCLUDE ? TSS_MASK_INCLUDE : TSS_MASK_EXCLUDE;
  const bool saved = ix->xchg.suppress;
  ix->xchg.suppress = true;  // rounds are rank-local; the caller merges across ranks
  int rc = TSS_OK;
  for (uint32_t qi = 0; qi < nq && !rc; ++qi) {
    if (mode == TSS_MASK_NONE)
      CU(cudaMemsetAsync(ix->d_round_mask, 0, words * sizeof(uint32_t), ix->stream));
    else
      CU(cudaMemcpyAsync(ix->d_round_mask, mask->d_words, (words - 1) * sizeof(uint32_t),
                         cudaMemcpyDeviceToDevice, ix->stream));
    for (uint32_t done = 0; done < k && !rc; done += TSS_MAX_FUSED_K) {
      const uint32_t kr = k - done < TSS_MAX_FUSED_K ? k - done : TSS_MAX_FUSED_K;
      uint64_t* dst = d_out + (size_t)qi * k + done;
      rc = enqueue_scan(ix, d_queries + (size_t)qi * ix->dim, 1, kr, &scratch, round_mode, dst);
      if (rc) break;
      cudaError_t e = tss::launch_mask_update_from_keys(ix->d_round_mask, ix->n_rows, dst, kr,
                                                        ix->row_base, round_mode == TSS_MASK_EXCLUDE,
                                                        ix->stream);
      if (e != cudaSuccess) rc = cuda_fail(e, "mask_update_from_keys launch");
      g_launches.fetch_add(1, std::memory_order_relaxed);
    }
  }
  ix->xchg.suppress = saved;
  scratch.d_words = nullptr;  // borrowed
  return rc;
}

// Shadow prefilter: one or two queries on an fp32 index whose bf16 shadow exists.  K1 streams the
// shadow (2 bytes per element) for the top-kc, kc = 64 or 128 > 4k; refine_kernel proves the
// fp32 top-k is among them, re-scores them from the fp32 rows and writes it.  A query it cannot
// prove (the kc-th shadow score within the error margin of the k-th) is redone by the fp32 scan.
int enqueue_prefilter(tss_index* ix, const float* d_queries, uint32_t nq, uint32_t k,
                      const tss_mask* mask, int mode, uint64_t* d_out) {
  int rc = ensure_gemm_ws(ix, true);
  if (rc) return rc;
  tss_index::Gemm& g = ix->gemm;
  const uint32_t kc = 4 * k <= 64 ? 64 : 128;
  const bool saved = ix->xchg.suppress;
  ix->xchg.suppress = true;  // rank-local candidates; the caller merges across ranks
  for (uint32_t qi = 0; qi < nq && !rc; ++qi)
    rc = enqueue_scan(ix, d_queries + (size_t)qi * ix->dim, 1, kc, mask, mode,
                      g.d_pref_keys + (size_t)qi * kc, nullptr, g.d_shadow);
  ix->xchg.suppress = saved;
  if (rc) return rc;
  cudaError_t e = tss::launch_refine(g.d_pref_keys, kc, d_queries, ix->d_rows, ix->dim,
                                     ix->stride_elems, (uint32_t)ix->row_base, nq, ix->n_rows,
                                     g.d_shadow_err, k, d_out, g.d_overflow, ix->stream);
  if (e != cudaSuccess) return cuda_fail(e, "refine_kernel launch");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  // a query the proof failed for is redone by the fp32 scan: guarded launches (nq <= 2 of
  // them always suffice), no host synchronisation
  return enqueue_fixups(ix, d_queries, nq, k, mask, mode, d_out);
}

// K2 over nq queries in batches the survivor pool can take (large k at large N: smaller batches)
int enqueue_gemm_batches(tss_index* ix, const float* d_queries, uint32_t nq, uint32_t k,
                         const tss_mask* mask, int mode, uint64_t* d_out) {
  if (ix->storage == TSS_F32 && nq <= kPrefilterMaxNq && k <= 32 &&
      ix->gemm.shadow_rows == ix->n_rows && ix->gemm.shadow_base == ix->d_rows)
    return enqueue_prefilter(ix, d_queries, nq, k, mask, mode, d_out);
  const uint32_t step = gemm_batch_limit(ix, k);
  for (uint32_t q0 = 0; q0 < nq; q0 += step) {
    uint32_t n = nq - q0 < step ? nq - q0 : step;
    int rc = enqueue_gemm(ix, d_queries + (size_t)q0 * ix->dim, n, k, mask, mode,
                          d_out + (size_t)q0 * k);
    if (rc) return rc;
  }
  return TSS_OK;
}

// local (per-shard) search of nq device-resident queries: picks K2 or K1.  *merged is set
// when d_out already holds the GLOBAL result (the K1 scan of a sharded index exchanges and
// merges inside its last CTA; K2 leaves that to NCCL + merge_gathered_kernel).
// gemm: the route (gemm_route) decided ONCE for the whole call -- every chunk of a call, and every
// rank of a shard group, must take the same one.
int enqueue_local(tss_index* ix, const float* d_queries, uint32_t nq, uint32_t k,
                  const tss_mask* mask, int mode, uint64_t* d_out, bool gemm, bool* merged) {
  *merged = false;
  if (!gemm) {
    if (k > TSS_MAX_FUSED_K) return enqueue_scan_rounds(ix, d_queries, nq, k, mask, mode, d_out);
    *merged = ix->comm && ix->xchg.ready;
    return enqueue_scan(ix, d_queries, nq, k, mask, mode, d_out);
  }
  return enqueue_gemm_batches(ix, d_queries, nq, k, mask, mode, d_out);
}

int validate_search(const tss_index* ix, const void* queries, uint32_t nq, uint32_t k) {
  if (!ix) return fail(TSS_ERR_INVALID_ARG, "index is NULL");
  if (!queries && nq) return fail(TSS_ERR_INVALID_ARG, "queries is NULL");
  if (k == 0 || k > TSS_MAX_K) return fail(TSS_ERR_INVALID_ARG, "k=%u outside [1,%u]", k, TSS_MAX_K);
  if (!ix->finalized) return fail(TSS_ERR_STATE, "search before tss_index_finalize");
  return TSS_OK;
}
int check_k_path(const tss_index* ix, uint32_t nq, uint32_t k, int mode) {
  (void)nq;
  (void)mode;
  if (k > TSS_MAX_FUSED_K && ix->comm && (uint64_t)ix->comm->nranks * k * 8 > 48 * 1024)
    return fail(TSS_ERR_INVALID_ARG, "k=%u over %d ranks exceeds the merge kernel (ranks*k <= 6144)",
                k, ix->comm->nranks);
  return TSS_OK;
}

// ---- searches in flight ------------------------------------------------------------------
// ix->mu held.  Enqueues one scan-path search of nq <= kPendingNq queries, k <= TSS_MAX_FUSED_K,
// whose result lands in a pending slot; *slot_out = -1 (and TSS_OK) when every slot is taken.
int pending_submit_locked(tss_index* ix, const float* queries, uint32_t nq, uint32_t k,
                          const tss_mask* mask, int mask_mode, int* slot_out) {
  *slot_out = -1;
  int s = -1;
  for (uint32_t i = 0; i < kPending; ++i) {
    uint32_t c = (ix->pend_next + i) % kPending;
    if (ix->pend[c].state == 0) {
      s = (int)c;
      break;
    }
  }
  if (s < 0) return TSS_OK;
  int rc;
  if ((rc = ensure_gather_ws(ix))) return rc;
  MaskReadScope mrs(mask, mask_mode, ix->stream);
  if (mrs.err != cudaSuccess) return cuda_fail(mrs.err, "mask ordering");
  const size_t qelems = (size_t)kPendingNq * ix->dim;
  float* hq = ix->h_pq + (size_t)s * qelems;
  float* dq = ix->d_pq + (size_t)s * qelems;
  uint64_t* hk = ix->h_pk + (size_t)s * kPendingNq * TSS_MAX_FUSED_K;
  // a lone query travels in the kernel parameters; the last CTA writes the result straight into
  // mapped pinned host memory -- no H2D / D2H copy operations at all
  const bool inline_q = nq == 1 && ix->dim <= 384;
  if (!inline_q) {
    memcpy(hq, queries, (size_t)nq * ix->dim * sizeof(float));
    CU(cudaMemcpyAsync(dq, hq, (size_t)nq * ix->dim * sizeof(float), cudaMemcpyHostToDevice, ix->stream));
  }
  const bool merged = ix->comm && ix->xchg.ready;
  const bool direct = !ix->comm || merged;
  rc = enqueue_scan(ix, dq, nq, k, mask, mask_mode, direct ? hk : ix->d_keys,
                    inline_q ? queries : nullptr);
  if (rc) return rc;
  if (!direct) {
    if ((rc = enqueue_gather_merge(ix, ix->d_keys, nq, k, ix->d_merged))) return rc;
    CU(cudaMemcpyAsync(hk, ix->d_merged, (size_t)nq * k * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                       ix->stream));
  }
  tss_index::Pending& pd = ix->pend[s];
  CU(cudaEventRecord(pd.ev, ix->stream));
  pd.nq = nq;
  pd.k = k;
  pd.gen = pd.gen + 1 ? pd.gen + 1 : 1;
  pd.state = 1;
  ix->pend_next = ((uint32_t)s + 1) % kPending;
  *slot_out = s;
  return TSS_OK;
}

// slot s is in state 2 (being collected by this thread); ix->mu NOT held while waiting
int pending_collect(tss_index* ix, int s, uint32_t* out_rows, float* out_scores, uint32_t* out_counts) {
  tss_index::Pending& pd = ix->pend[s];
  int rc = TSS_OK;
  {
    DeviceGuard g(ix->device);
    cudaError_t e = cudaEventSynchronize(pd.ev);
    if (e != cudaSuccess) rc = cuda_fail(e, "waiting for a pending search");
  }
  if (!rc && ix->comm && ix->xchg.ready && *ix->h_status) {
    *ix->h_status = 0;
    rc = fail(TSS_ERR_NCCL, "a rank of the shard group did not deliver its top-k within 20 s");
  }
  if (!rc) {
    const uint64_t* hk = ix->h_pk + (size_t)s * kPendingNq * TSS_MAX_FUSED_K;
    tss_unpack_keys(hk, (uint64_t)pd.nq * pd.k, out_rows, out_scores);
    for (uint32_t qi = 0; qi < pd.nq; ++qi) {
      uint32_t c = 0;
      while (c < pd.k && hk[(size_t)qi * pd.k + c] != 0) ++c;
      out_counts[qi] = c;
    }
  }
  {
    std::lock_guard<std::mutex> lock(ix->mu);
    pd.state = 0;
  }
  ix->pend_cv.notify_all();
  return rc;
}

int validate_host_queries(const tss_index* ix, const float* queries, uint32_t nq) {
  for (uint64_t i = 0; i < (uint64_t)nq * ix->dim; ++i)
    if (!std::isfinite(queries[i]))
      return fail(TSS_ERR_INVALID_ARG, "query %llu contains NaN or Inf",
                  (unsigned long long)(i / ix->dim));
  return TSS_OK;
}

}  // namespace

extern "C" {

int tss_abi_version(void) { return TSS_ABI_VERSION; }
const char* tss_last_error(void) { return g_err; }
uint64_t tss_launch_count(void) { return g_launches.load(); }

int tss_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// ---- index ---------------------------------------------------------------------------
int tss_index_create(tss_index** out, uint32_t dim, int storage, int device) {
  if (!out) return fail(TSS_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (dim == 0) return fail(TSS_ERR_INVALID_ARG, "dim must be > 0");
  int ns = tss::storage_stripes_for_dim(dim);
  if (!ns) return fail(TSS_ERR_INVALID_ARG, "dim=%u > 1024 is not supported", dim);
  if (storage != TSS_F32 && storage != TSS_BF16)
    return fail(TSS_ERR_INVALID_ARG, "storage %d is not TSS_F32/TSS_BF16", storage);
  int ndev = tss_device_count();
  if (ndev == 0) return fail(TSS_ERR_CUDA, "no CUDA device visible (libtss has no CPU path)");
  if (device < 0 || device >= ndev) return fail(TSS_ERR_INVALID_ARG, "device %d of %d", device, ndev);
  DeviceGuard g(device);
  if (!g.ok) return fail(TSS_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(TSS_ERR_CUDA, "device %d is sm_%d%d; libtss is built for sm_100a only", device,
                prop.major, prop.minor);
  tss_index* ix = new (std::nothrow) tss_index();
  if (!ix) return fail(TSS_ERR_OOM, "host allocation failed");
  ix->device = device;
  ix->dim = dim;
  ix->storage = storage;
  ix->ns = ns;
  ix->stride_elems = (uint32_t)ns * 128u;
  ix->row_bytes = (size_t)ix->stride_elems * (storage == TSS_BF16 ? 2 : 4);
  ix->num_sms = prop.multiProcessorCount;
  cudaError_t e;
#define ALLOC(expr)                    \
  if ((e = (expr)) != cudaSuccess) {   \
    tss_index_destroy(ix);             \
    return cuda_fail(e, #expr);        \
  }
  ALLOC(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking))
  ALLOC(cudaMalloc(&ix->d_flag, sizeof(int)))
  ALLOC(cudaMalloc(&ix->d_queries, (size_t)kWsQueries * dim * sizeof(float)))
  ALLOC(cudaMalloc(&ix->d_keys, (size_t)kWsQueries * TSS_MAX_K * sizeof(uint64_t)))
  ALLOC(cudaMalloc(&ix->d_partials, (size_t)kScanSlots * kMaxBq * ix->num_sms * 128 * sizeof(uint64_t)))
  ALLOC(cudaMalloc(&ix->d_counter, 32 * sizeof(unsigned int)))
  ALLOC(cudaMemset(ix->d_counter, 0, 32 * sizeof(unsigned int)))
  ALLOC(cudaMalloc(&ix->d_walk_ctr, (size_t)kScanSlots * tss::kWalkCounters * 32 * sizeof(unsigned int)))
  ALLOC(cudaMemset(ix->d_walk_ctr, 0, (size_t)kScanSlots * tss::kWalkCounters * 32 * sizeof(unsigned int)))
  if (const char* sf = getenv("TSS_PDL")) ix->pdl = atoi(sf) != 0;
  if (const char* sf = getenv("TSS_STATIC_FRAC")) ix->static_frac = (float)atof(sf);
  if (const char* sf = getenv("TSS_FINE_ROUNDS")) ix->fine_rounds = (float)atof(sf);
  if (const char* sf = getenv("TSS_DYN_CHUNK")) ix->dyn_chunk = (uint32_t)atoi(sf);
  if (ix->dyn_chunk < 1) ix->dyn_chunk = 1;
  if (const char* sf = getenv("TSS_WALK_STATIC")) ix->walk_static_frac = (float)atof(sf);
  if (const char* sf = getenv("TSS_WALK_RUN")) {  // diagnostics: log2 run (0..5), +8 = consecutive runs to different CTAs
    const uint32_t v = (uint32_t)atoi(sf);
    ix->walk_run_log2 = ((v & 7u) > 5 ? 5u : (v & 7u)) | (v & 24u);  // +16: stay on the home counter
  }
  ALLOC(cudaMallocHost(&ix->h_queries, (size_t)kWsQueries * dim * sizeof(float)))
  ALLOC(cudaMallocHost(&ix->h_keys, (size_t)kWsQueries * TSS_MAX_K * sizeof(uint64_t)))
  ALLOC(cudaMallocHost(&ix->h_status, 64))
  memset(ix->h_status, 0, 64);
  ALLOC(cudaMallocHost(&ix->h_pq, (size_t)kPending * kPendingNq * dim * sizeof(float)))
  ALLOC(cudaMalloc(&ix->d_pq, (size_t)kPending * kPendingNq * dim * sizeof(float)))
  ALLOC(cudaMallocHost(&ix->h_pk, (size_t)kPending * kPendingNq * TSS_MAX_FUSED_K * sizeof(uint64_t)))
  for (auto& pd : ix->pend) ALLOC(cudaEventCreateWithFlags(&pd.ev, cudaEventDisableTiming))
  if (const char* sf = getenv("TSS_GEMM_MIN_NQ")) {  // diagnostics: one plain threshold
    ix->gemm_min_nq = (uint32_t)atoi(sf);
    ix->gemm_small_nq = 0xFFFFFFFFu;
  }
#undef ALLOC
  *out = ix;
  return TSS_OK;
}

void tss_index_destroy(tss_index* ix) {
  if (!ix) return;
  DeviceGuard g(ix->device);
  if (ix->stream) cudaStreamSynchronize(ix->stream);
  cudaFree(ix->d_rows);
  cudaFree(ix->d_stage);
  if (ix->h_stage) cudaFreeHost(ix->h_stage);
  for (cudaEvent_t ev : ix->stage_ev)
    if (ev) cudaEventDestroy(ev);
  cudaFree(ix->d_flag);
  cudaFree(ix->d_queries);
  cudaFree(ix->d_keys);
  cudaFree(ix->d_gather);
  cudaFree(ix->d_merged);
  cudaFree(ix->d_partials);
  cudaFree(ix->d_counter);
  cudaFree(ix->d_walk_ctr);
  if (ix->h_queries) cudaFreeHost(ix->h_queries);
  if (ix->h_keys) cudaFreeHost(ix->h_keys);
  if (ix->h_status) cudaFreeHost(ix->h_status);
  if (ix->h_pq) cudaFreeHost(ix->h_pq);
  cudaFree(ix->d_pq);
  if (ix->h_pk) cudaFreeHost(ix->h_pk);
  for (auto& pd : ix->pend)
    if (pd.ev) cudaEventDestroy(pd.ev);
  for (int r = 0; r < 8; ++r)
    if (ix->xchg.peer[r] && ix->xchg.peer[r] != ix->xchg.local) cudaIpcCloseMemHandle(ix->xchg.peer[r]);
  cudaFree(ix->xchg.local);
  cudaFree(ix->d_round_mask);
  cudaFree(ix->gemm.d_inv_norm);
  cudaFree(ix->gemm.d_shadow);
  cudaFree(ix->gemm.d_shadow_err);
  cudaFree(ix->gemm.d_qbf16);
  cudaFree(ix->gemm.d_inv_q);
  cudaFree(ix->gemm.d_margin);
  cudaFree(ix->gemm.d_thr);
  cudaFree(ix->gemm.d_tile_max);
  cudaFree(ix->gemm.d_cand);
  cudaFree(ix->gemm.d_cand_count);
  cudaFree(ix->gemm.d_overflow);
  cudaFree(ix->gemm.d_pref_keys);
  if (ix->gemm.h_cand_count) cudaFreeHost(ix->gemm.h_cand_count);
  cudaFree(ix->gemm.d_redo);
  if (ix->gemm.h_redo) cudaFreeHost(ix->gemm.h_redo);
  if (ix->stream) cudaStreamDestroy(ix->stream);
  delete ix;
}

int tss_index_reserve(tss_index* ix, uint64_t nrows) {
  if (!ix) return fail(TSS_ERR_INVALID_ARG, "index is NULL");
  std::lock_guard<std::mutex> lock(ix->mu);
  if (nrows >= 0xFFFFFFFFull) return fail(TSS_ERR_INVALID_ARG, "row ids are 32-bit");
  DeviceGuard g(ix->device);
  return ensure_capacity(ix, nrows);
}

int tss_index_add(tss_index* ix, const float* rows, uint64_t nrows) {
  if (!ix) return fail(TSS_ERR_INVALID_ARG, "index is NULL");
  std::lock_guard<std::mutex> lock(ix->mu);
  if (!nrows) return TSS_OK;
  if (!rows) return fail(TSS_ERR_INVALID_ARG, "rows is NULL");
  if (ix->row_base + ix->n_rows + nrows >= 0xFFFFFFFFull)
    return fail(TSS_ERR_INVALID_ARG, "row ids are 32-bit");
  DeviceGuard g(ix->device);
  int rc = ensure_capacity(ix, ix->n_rows + nrows);
  if (rc) return rc;
  if (!ix->d_stage) {
    CU(cudaMalloc(&ix->d_stage, 2 * kStageBytes));
    CU(cudaMallocHost(&ix->h_stage, 2 * kStageBytes));
    for (cudaEvent_t& ev : ix->stage_ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  }
  CU(cudaMemsetAsync(ix->d_flag, 0, sizeof(int), ix->stream));
  const uint64_t chunk_rows = kStageBytes / ((size_t)ix->dim * sizeof(float));
  if (!chunk_rows) return fail(TSS_ERR_INVALID_ARG, "dim too large for the staging buffer");
  uint32_t ci = 0;
  for (uint64_t r0 = 0; r0 < nrows; r0 += chunk_rows, ++ci) {
    const uint64_t n = nrows - r0 < chunk_rows ? nrows - r0 : chunk_rows;
    const size_t bytes = (size_t)n * ix->dim * sizeof(float);
    const uint32_t b = ci & 1u;
    float* hs = ix->h_stage + b * (kStageBytes / sizeof(float));
    float* ds = ix->d_stage + b * (kStageBytes / sizeof(float));
    // buffer b was last used by chunk ci - 2: its DMA and packing must be done
    if (ci >= 2) CU(cudaEventSynchronize(ix->stage_ev[b]));
    staged_copy(hs, rows + (size_t)r0 * ix->dim, bytes);  // overlaps chunk ci - 1 on the device
    CU(cudaMemcpyAsync(ds, hs, bytes, cudaMemcpyHostToDevice, ix->stream));
    cudaError_t e = tss::launch_check_finite(ds, n * ix->dim, ix->d_flag, ix->stream);
    if (e != cudaSuccess) return cuda_fail(e, "check_finite launch");
    e = tss::launch_pack_rows(ds, ix->d_rows + (size_t)(ix->n_rows + r0) * ix->row_bytes, n,
                              ix->dim, ix->stride_elems, ix->storage == TSS_BF16, ix->stream);
    if (e != cudaSuccess) return cuda_fail(e, "pack_rows launch");
    g_launches.fetch_add(2, std::memory_order_relaxed);
    CU(cudaEventRecord(ix->stage_ev[b], ix->stream));
  }
  int flag = 0;
  CU(cudaMemcpyAsync(&flag, ix->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  if (flag)
    return fail(TSS_ERR_INVALID_ARG, "rows contain NaN or Inf; index unchanged (%llu rows)",
                (unsigned long long)ix->n_rows);
  ix->n_rows += nrows;
  ix->finalized = false;
  return TSS_OK;
}

int tss_index_add_synthetic(tss_index* ix, uint64_t row_begin, uint64_t nrows, uint64_t seed) {
  if (!ix) return fail(TSS_ERR_INVALID_ARG, "index is NULL");
  std::lock_guard<std::mutex> lock(ix->mu);
  if (!nrows) return TSS_OK;
  if (ix->row_base + ix->n_rows + nrows >= 0xFFFFFFFFull)
    return fail(TSS_ERR_INVALID_ARG, "row ids are 32-bit");
  DeviceGuard g(ix->device);
  int rc = ensure_capacity(ix, ix->n_rows + nrows);
  if (rc) return rc;
  cudaError_t e = tss::launch_synth_fill(ix->d_rows + (size_t)ix->n_rows * ix->row_bytes, row_begin,
                                         nrows, ix->dim, ix->stride_elems, ix->storage == TSS_BF16,
                                         seed, ix->stream);
  if (e != cudaSuccess) return cuda_fail(e, "synth_fill launch");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  ix->n_rows += nrows;
  ix->finalized = false;
  return TSS_OK;
}

int tss_index_finalize(tss_index* ix) {
  if (!ix) return fail(TSS_ERR_INVALID_ARG, "index is NULL");
  std::lock_guard<std::mutex> lock(ix->mu);
  DeviceGuard g(ix->device);
  if (!ix->d_rows) {
    int rc = ensure_capacity(ix, 1);  // an empty index still scans (and finds nothing)
    if (rc) return rc;
  }
  CU(cudaStreamSynchronize(ix->stream));
  if (ix->d_stage) {  // the upload pipeline is only needed while rows arrive
    cudaFree(ix->d_stage);
    ix->d_stage = nullptr;
    cudaFreeHost(ix->h_stage);
    ix->h_stage = nullptr;
    for (cudaEvent_t& ev : ix->stage_ev) {
      cudaEventDestroy(ev);
      ev = nullptr;
    }
  }
  ix->finalized = true;
  return TSS_OK;
}

uint64_t tss_index_size(const tss_index* ix) { return ix ? ix->n_rows : 0; }
uint32_t tss_index_dim(const tss_index* ix) { return ix ? ix->dim : 0; }

int tss_index_get_rows(tss_index* ix, uint64_t row_begin, uint64_t nrows, float* out) {
  if (!ix || !out) return fail(TSS_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> lock(ix->mu);
  if (row_begin + nrows > ix->n_rows) return fail(TSS_ERR_INVALID_ARG, "row range out of bounds");
  if (!nrows) return TSS_OK;
  DeviceGuard g(ix->device);
  float* d_tmp = nullptr;
  CU(cudaMalloc(&d_tmp, (size_t)nrows * ix->dim * sizeof(float)));
  cudaError_t e = tss::launch_unpack_rows(ix->d_rows + (size_t)row_begin * ix->row_bytes, d_tmp,
                                          nrows, ix->dim, ix->stride_elems,
                                          ix->storage == TSS_BF16, ix->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(out, d_tmp, (size_t)nrows * ix->dim * sizeof(float), cudaMemcpyDeviceToHost,
                        ix->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
  cudaFree(d_tmp);
  if (e != cudaSuccess) return cuda_fail(e, "get_rows");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return TSS_OK;
}

// ---- on-disk format (VectorIndex::save_to_disk / load_from_disk,
// reference src/vector.rs:83-95 are TODO stubs) ---------------------------------------------
// 64-byte header + the padded rows exactly as they sit in HBM, so loading is a straight
// file -> pinned buffer -> HBM copy with no repacking.
namespace {
struct IndexFileHeader {
  char magic[8];  // "TSSIDX01"
  uint32_t dim, storage, stride_elems, reserved;
  uint64_t n_rows, row_bytes;
  uint8_t pad[24];
};
static_assert(sizeof(IndexFileHeader) == 64, "header is 64 bytes");
constexpr size_t kIoChunk = 32u << 20;
}  // namespace

int tss_index_save(tss_index* ix, const char* path) {
  if (!ix || !path) return fail(TSS_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> lock(ix->mu);
  DeviceGuard g(ix->device);
  CU(cudaStreamSynchronize(ix->stream));
  FILE* f = fopen(path, "wb");
  if (!f) return fail(TSS_ERR_INVALID_ARG, "cannot open %s for writing", path);
  IndexFileHeader h{};
  memcpy(h.magic, "TSSIDX01", 8);
  h.dim = ix->dim;
  h.storage = (uint32_t)ix->storage;
  h.stride_elems = ix->stride_elems;
  h.n_rows = ix->n_rows;
  h.row_bytes = ix->row_bytes;
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  void* hbuf = nullptr;
  if (ok && cudaMallocHost(&hbuf, kIoChunk) != cudaSuccess) ok = false;
  const size_t total = (size_t)ix->n_rows * ix->row_bytes;
  for (size_t off = 0; ok && off < total; off += kIoChunk) {
    size_t n = total - off < kIoChunk ? total - off : kIoChunk;
    ok = cudaMemcpy(hbuf, ix->d_rows + off, n, cudaMemcpyDeviceToHost) == cudaSuccess &&
         fwrite(hbuf, 1, n, f) == n;
  }
  if (hbuf) cudaFreeHost(hbuf);
  ok = (fclose(f) == 0) && ok;
  if (!ok) return fail(TSS_ERR_STATE, "writing %s failed", path);
  return TSS_OK;
}

int tss_index_load(tss_index** out, const char* path, int device) {
  if (!out || !path) return fail(TSS_ERR_INVALID_ARG, "NULL argument");
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) return fail(TSS_ERR_INVALID_ARG, "cannot open %s", path);
  IndexFileHeader h{};
  if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "TSSIDX01", 8) != 0) {
    fclose(f);
    return fail(TSS_ERR_INVALID_ARG, "%s is not a TSSIDX01 file", path);
  }
  const int ns = h.dim ? tss::storage_stripes_for_dim(h.dim) : 0;
  const uint64_t elem = h.storage == TSS_BF16 ? 2 : 4;
  if (!ns || (h.storage != TSS_F32 && h.storage != TSS_BF16) ||
      h.stride_elems != (uint32_t)ns * 128u || h.row_bytes != h.stride_elems * elem ||
      h.n_rows >= 0xFFFFFFFFull) {
    fclose(f);
    return fail(TSS_ERR_INVALID_ARG, "%s: inconsistent header", path);
  }
  tss_index* ix = nullptr;
  int rc = tss_index_create(&ix, h.dim, (int)h.storage, device);
  if (rc) {
    fclose(f);
    return rc;
  }
  DeviceGuard g(device);
  rc = ensure_capacity(ix, h.n_rows ? h.n_rows : 1);
  // two pinned buffers: the read of chunk i+1 from the file overlaps the DMA of chunk i
  uint8_t* hbuf = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  if (!rc && (cudaMallocHost(&hbuf, 2 * kIoChunk) != cudaSuccess ||
              cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) != cudaSuccess ||
              cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) != cudaSuccess))
    rc = fail(TSS_ERR_OOM, "pinned buffer");
  const size_t total = (size_t)h.n_rows * h.row_bytes;
  uint32_t ci = 0;
  for (size_t off = 0; !rc && off < total; off += kIoChunk, ++ci) {
    size_t n = total - off < kIoChunk ? total - off : kIoChunk;
    uint8_t* hb = hbuf + (ci & 1u) * kIoChunk;
    if (ci >= 2 && cudaEventSynchronize(ev[ci & 1u]) != cudaSuccess)
      rc = fail(TSS_ERR_CUDA, "upload of %s failed", path);
    else if (fread(hb, 1, n, f) != n)
      rc = fail(TSS_ERR_INVALID_ARG, "%s is truncated", path);
    else if (cudaMemcpyAsync(ix->d_rows + off, hb, n, cudaMemcpyHostToDevice, ix->stream) != cudaSuccess ||
             cudaEventRecord(ev[ci & 1u], ix->stream) != cudaSuccess)
      rc = fail(TSS_ERR_CUDA, "upload of %s failed", path);
  }
  if (cudaStreamSynchronize(ix->stream) != cudaSuccess && !rc)
    rc = fail(TSS_ERR_CUDA, "upload of %s failed", path);
  if (hbuf) cudaFreeHost(hbuf);
  for (cudaEvent_t e2 : ev)
    if (e2) cudaEventDestroy(e2);
  fclose(f);
  if (rc) {
    tss_index_destroy(ix);
    return rc;
  }
  ix->n_rows = h.n_rows;
  if ((rc = tss_index_finalize(ix))) {
    tss_index_destroy(ix);
    return rc;
  }
  *out = ix;
  return TSS_OK;
}

// ---- search ---------------------------------------------------------------------------
int tss_index_search_device(tss_index* ix, const float* d_queries, uint32_t nq, uint32_t k,
                            const tss_mask* mask, int mask_mode, uint64_t* d_out_keys) {
  if (!ix) return fail(TSS_ERR_INVALID_ARG, "index is NULL");
  std::lock_guard<std::mutex> lock(ix->mu);
  int rc = validate_search(ix, d_queries, nq, k);
  if (rc) return rc;
  if (!d_out_keys) return fail(TSS_ERR_INVALID_ARG, "d_out_keys is NULL");
  if ((rc = check_mask(ix, mask, mask_mode))) return rc;
  if ((rc = check_k_path(ix, nq, k, mask_mode))) return rc;
  if (!nq) return TSS_OK;
  DeviceGuard g(ix->device);
  MaskReadScope mrs(mask, mask_mode, ix->stream);
  if (mrs.err != cudaSuccess) return cuda_fail(mrs.err, "mask ordering");
  bool merged = false;
  const bool gemm = gemm_route(ix, nq, k, mask_mode);  // the whole call takes one route
  struct NoSync {  // K2 fix-ups stay on the device for the duration of this call
    tss_index* ix;
    explicit NoSync(tss_index* i) : ix(i) { ix->no_host_sync = true; }
    ~NoSync() { ix->no_host_sync = false; }
  } no_sync(ix);
  if (!ix->comm) return enqueue_local(ix, d_queries, nq, k, mask, mask_mode, d_out_keys, gemm, &merged);
  if ((rc = ensure_gather_ws(ix))) return rc;
  const bool fused = ix->xchg.ready && !gemm && k <= TSS_MAX_FUSED_K;
  for (uint32_t q0 = 0; q0 < nq; q0 += kWsQueries) {
    uint32_t n = nq - q0 < kWsQueries ? nq - q0 : kWsQueries;
    uint64_t* d_local = fused ? d_out_keys + (size_t)q0 * k : ix->d_keys;
    if ((rc = enqueue_local(ix, d_queries + (size_t)q0 * ix->dim, n, k, mask, mask_mode, d_local,
                            gemm, &merged)))
      return rc;
    if (merged != fused) return fail(TSS_ERR_STATE, "sharded search: route and exchange disagree");
    if (!merged && (rc = enqueue_gather_merge(ix, ix->d_keys, n, k, d_out_keys + (size_t)q0 * k)))
      return rc;
  }
  return TSS_OK;
}

void tss_unpack_keys(const uint64_t* keys, uint64_t n, uint32_t* out_rows, float* out_scores) {
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t key = keys[i];
    if (key == 0) {
      if (out_rows) out_rows[i] = TSS_ROW_NONE;
      if (out_scores) out_scores[i] = 0.0f;
      continue;
    }
    uint32_t o = (uint32_t)(key >> 32);
    uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    float s;
    memcpy(&s, &u, 4);
    if (out_rows) out_rows[i] = 0xFFFFFFFFu - (uint32_t)key;
    if (out_scores) out_scores[i] = s;
  }
}

int tss_index_search(tss_index* ix, const float* queries, uint32_t nq, uint32_t k,
                     const tss_mask* mask, int mask_mode, uint32_t* out_rows, float* out_scores,
                     uint32_t* out_counts) {
  if (!ix) return fail(TSS_ERR_INVALID_ARG, "index is NULL");
  std::unique_lock<std::mutex> lock(ix->mu);
  int rc = validate_search(ix, queries, nq, k);
  if (rc) return rc;
  if (!out_rows || !out_scores || !out_counts)
    return fail(TSS_ERR_INVALID_ARG, "output pointer is NULL");
  if ((rc = check_mask(ix, mask, mask_mode))) return rc;
  if ((rc = check_k_path(ix, nq, k, mask_mode))) return rc;
  if (!nq) return TSS_OK;
  if ((rc = validate_host_queries(ix, queries, nq))) return rc;
  DeviceGuard g(ix->device);
  if (nq <= kPendingNq && k <= TSS_MAX_FUSED_K && !gemm_route(ix, nq, k, mask_mode)) {
    // one scan launch: enqueue under the lock, wait outside it, so the searches of several host
    // threads queue up back to back on the device (the next one's prologue overlaps this one's
    // merges) instead of each waiting out the other's launch + wake-up latency
    int s = -1;
    for (;;) {
      if ((rc = pending_submit_locked(ix, queries, nq, k, mask, mask_mode, &s))) return rc;
      if (s >= 0) break;
      // every slot is taken.  If some thread is collecting one, it is about to be handed back:
      // wait for it.  If all of them are submitted-but-uncollected tickets, nobody may ever
      // collect while this thread waits (they could all be this thread's own): take the
      // serialised path below instead, which needs no slot.
      bool collecting = false;
      for (const auto& pd : ix->pend) collecting |= pd.state == 2;
      if (!collecting) break;
      ix->pend_cv.wait(lock);
    }
    if (s >= 0) {
      ix->pend[s].state = 2;
      lock.unlock();
      return pending_collect(ix, s, out_rows, out_scores, out_counts);
    }
  }
  if ((rc = ensure_gather_ws(ix))) return rc;
  MaskReadScope mrs(mask, mask_mode, ix->stream);
  if (mrs.err != cudaSuccess) return cuda_fail(mrs.err, "mask ordering");
  for (uint32_t q0 = 0; q0 < nq; q0 += kWsQueries) {
    uint32_t n = nq - q0 < kWsQueries ? nq - q0 : kWsQueries;
    size_t qbytes = (size_t)n * ix->dim * sizeof(float);
    const float* hq = queries + (size_t)q0 * ix->dim;
    // a batch takes the tensor-core path as a whole (nq, not n, decides)
    const bool gemm = gemm_route(ix, nq, k, mask_mode);
    const bool rounds = !gemm && k > TSS_MAX_FUSED_K;
    // scan path: a lone query travels in the kernel parameters, and the last CTA writes the
    // result straight into mapped pinned host memory -- no H2D / D2H copy operations at all
    const bool inline_q = !gemm && !rounds && n == 1 && ix->dim <= 384;
    if (!inline_q) {
      memcpy(ix->h_queries, hq, qbytes);
      CU(cudaMemcpyAsync(ix->d_queries, ix->h_queries, qbytes, cudaMemcpyHostToDevice, ix->stream));
    }
    bool merged = false;
    bool direct = false;  // result already lands in h_keys
    if (gemm) {
      rc = enqueue_gemm_batches(ix, ix->d_queries, n, k, mask, mask_mode, ix->d_keys);
    } else if (rounds) {
      rc = enqueue_scan_rounds(ix, ix->d_queries, n, k, mask, mask_mode, ix->d_keys);
    } else {
      merged = ix->comm && ix->xchg.ready;
      direct = !ix->comm || merged;
      rc = enqueue_scan(ix, ix->d_queries, n, k, mask, mask_mode, direct ? ix->h_keys : ix->d_keys,
                        inline_q ? hq : nullptr);
    }
    if (rc) return rc;
    if (!direct) {
      const uint64_t* d_res = ix->d_keys;
      if (ix->comm && !merged) {
        if ((rc = enqueue_gather_merge(ix, ix->d_keys, n, k, ix->d_merged))) return rc;
        d_res = ix->d_merged;
      }
      CU(cudaMemcpyAsync(ix->h_keys, d_res, (size_t)n * k * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                         ix->stream));
    }
    CU(cudaStreamSynchronize(ix->stream));
    if (merged && *ix->h_status) {
      *ix->h_status = 0;
      return fail(TSS_ERR_NCCL, "a rank of the shard group did not deliver its top-k within 20 s");
    }
    tss_unpack_keys(ix->h_keys, (uint64_t)n * k, out_rows + (size_t)q0 * k,
                    out_scores + (size_t)q0 * k);
    for (uint32_t qi = 0; qi < n; ++qi) {
      uint32_t c = 0;
      while (c < k && ix->h_keys[(size_t)qi * k + c] != 0) ++c;
      out_counts[q0 + qi] = c;
    }
  }
  return TSS_OK;
}

int tss_index_search_submit(tss_index* ix, const float* queries, uint32_t nq, uint32_t k,
                            const tss_mask* mask, int mask_mode, uint64_t* out_ticket) {
  if (!ix) return fail(TSS_ERR_INVALID_ARG, "index is NULL");
  if (!out_ticket) return fail(TSS_ERR_INVALID_ARG, "out_ticket is NULL");
  *out_ticket = 0;
  std::lock_guard<std::mutex> lock(ix->mu);
  int rc = validate_search(ix, queries, nq, k);
  if (rc) return rc;
  if (nq == 0 || nq > kPendingNq || k > TSS_MAX_FUSED_K)
    return fail(TSS_ERR_INVALID_ARG, "a pending search takes 1..%u queries and k <= %u (got nq=%u k=%u)",
                kPendingNq, TSS_MAX_FUSED_K, nq, k);
  if ((rc = check_mask(ix, mask, mask_mode))) return rc;
  if ((rc = validate_host_queries(ix, queries, nq))) return rc;
  DeviceGuard g(ix->device);
  int s = -1;
  if ((rc = pending_submit_locked(ix, queries, nq, k, mask, mask_mode, &s))) return rc;
  if (s < 0)
    return fail(TSS_ERR_STATE, "%u searches are already pending on this index: collect one first", kPending);
  *out_ticket = ((uint64_t)ix->pend[s].gen << 8) | (uint64_t)(s + 1);
  return TSS_OK;
}

int tss_index_search_collect(tss_index* ix, uint64_t ticket, uint32_t* out_rows, float* out_scores,
                             uint32_t* out_counts) {
  if (!ix) return fail(TSS_ERR_INVALID_ARG, "index is NULL");
  if (!out_rows || !out_scores || !out_counts)
    return fail(TSS_ERR_INVALID_ARG, "output pointer is NULL");
  const uint32_t slot1 = (uint32_t)(ticket & 0xFF), gen = (uint32_t)(ticket >> 8);
  if (slot1 == 0 || slot1 > kPending) return fail(TSS_ERR_INVALID_ARG, "not a ticket of tss_index_search_submit");
  const int s = (int)slot1 - 1;
  {
    std::lock_guard<std::mutex> lock(ix->mu);
    if (ix->pend[s].state != 1 || ix->pend[s].gen != gen)
      return fail(TSS_ERR_STATE, "ticket was already collected (or never issued by this index)");
    ix->pend[s].state = 2;
  }
  return pending_collect(ix, s, out_rows, out_scores, out_counts);
}

int tss_index_search_prefix_submit(tss_index* ix, tss_terms* t, const char* prefix, uint32_t len,
                                   int kind, tss_mask* scratch, const float* queries, uint32_t nq,
                                   uint32_t k, uint64_t* out_ticket) {
  if (!ix) return fail(TSS_ERR_INVALID_ARG, "index is NULL");
  if (!out_ticket) return fail(TSS_ERR_INVALID_ARG, "out_ticket is NULL");
  *out_ticket = 0;
  int rc;
  uint64_t row_base;
  {  // everything the search would reject is rejected before the scratch mask is touched
    std::lock_guard<std::mutex> lock(ix->mu);
    if ((rc = validate_search(ix, queries, nq, k))) return rc;
    if (nq == 0 || nq > kPendingNq || k > TSS_MAX_FUSED_K)
      return fail(TSS_ERR_INVALID_ARG, "a pending search takes 1..%u queries and k <= %u (got nq=%u k=%u)",
                  kPendingNq, TSS_MAX_FUSED_K, nq, k);
    if ((rc = check_mask(ix, scratch, TSS_MASK_INCLUDE))) return rc;
    if ((rc = validate_host_queries(ix, queries, nq))) return rc;
    bool free_slot = false;
    for (const auto& pd : ix->pend) free_slot |= pd.state == 0;
    if (!free_slot)
      return fail(TSS_ERR_STATE, "%u searches are already pending on this index: collect one first", kPending);
    row_base = ix->row_base;
  }
  if ((rc = tss_prefix_mask_fresh(t, prefix, len, kind, scratch, row_base, nullptr))) return rc;
  return tss_index_search_submit(ix, queries, nq, k, scratch, TSS_MASK_INCLUDE, out_ticket);
}

int tss_index_search_prefix(tss_index* ix, tss_terms* t, const char* prefix, uint32_t len, int kind,
                            tss_mask* scratch, const float* queries, uint32_t nq, uint32_t k,
                            uint32_t* out_rows, float* out_scores, uint32_t* out_counts) {
  if (!ix) return fail(TSS_ERR_INVALID_ARG, "index is NULL");
  int rc;
  uint64_t row_base;
  {  // everything the search would reject is rejected before the scratch mask is touched
    std::lock_guard<std::mutex> lock(ix->mu);
    if ((rc = validate_search(ix, queries, nq, k))) return rc;
    if (!out_rows || !out_scores || !out_counts)
      return fail(TSS_ERR_INVALID_ARG, "output pointer is NULL");
    if ((rc = check_mask(ix, scratch, TSS_MASK_INCLUDE))) return rc;
    if ((rc = check_k_path(ix, nq, k, TSS_MASK_INCLUDE))) return rc;
    if ((rc = validate_host_queries(ix, queries, nq))) return rc;
    row_base = ix->row_base;
  }
  if ((rc = tss_prefix_mask_fresh(t, prefix, len, kind, scratch, row_base, nullptr))) return rc;
  return tss_index_search(ix, queries, nq, k, scratch, TSS_MASK_INCLUDE, out_rows, out_scores, out_counts);
}

// ---- sharding ---------------------------------------------------------------------------
int tss_comm_unique_id(uint8_t out_id[128]) {
  if (!out_id) return fail(TSS_ERR_INVALID_ARG, "out_id is NULL");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  int r = g_nccl.GetUniqueId(&id);
  if (r != 0) return fail(TSS_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
  memcpy(out_id, id.internal, 128);
  return TSS_OK;
}

int tss_comm_create(tss_comm** out, const uint8_t id[128], int rank, int nranks, int device) {
  if (!out || !id) return fail(TSS_ERR_INVALID_ARG, "NULL argument");
  *out = nullptr;
  if (nranks < 1 || rank < 0 || rank >= nranks)
    return fail(TSS_ERR_INVALID_ARG, "rank %d of %d", rank, nranks);
  int rc = load_nccl();
  if (rc) return rc;
  DeviceGuard g(device);
  if (!g.ok) return fail(TSS_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  ncclUniqueId uid;
  memcpy(uid.internal, id, 128);
  ncclComm_t comm = nullptr;
  int r = g_nccl.CommInitRank(&comm, nranks, uid, rank);
  if (r != 0) return fail(TSS_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
  tss_comm* c = new (std::nothrow) tss_comm();
  if (!c) return fail(TSS_ERR_OOM, "host allocation failed");
  c->comm = comm;
  c->rank = rank;
  c->nranks = nranks;
  c->device = device;
  *out = c;
  return TSS_OK;
}

void tss_comm_destroy(tss_comm* c) {
  if (!c) return;
  DeviceGuard g(c->device);
  if (c->comm && g_nccl.ok) g_nccl.CommDestroy(c->comm);
  delete c;
}

namespace {
// all-gather `bytes` bytes per rank over the comm through a scratch device buffer
int comm_allgather_host(tss_index* ix, tss_comm* comm, const void* mine, void* all, size_t bytes) {
  uint8_t* d = nullptr;
  CU(cudaMalloc(&d, bytes * (size_t)(comm->nranks + 1)));
  cudaError_t e = cudaMemcpyAsync(d, mine, bytes, cudaMemcpyHostToDevice, ix->stream);
  int nrc = 0;
  if (e == cudaSuccess)
    nrc = g_nccl.AllGather(d, d + bytes, bytes, /*ncclUint8*/ 1, comm->comm, ix->stream);
  if (e == cudaSuccess && nrc == 0)
    e = cudaMemcpyAsync(all, d + bytes, bytes * comm->nranks, cudaMemcpyDeviceToHost, ix->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
  cudaFree(d);
  if (nrc != 0) return fail(TSS_ERR_NCCL, "ncclAllGather (shard setup): %s", g_nccl.GetErrorString(nrc));
  if (e != cudaSuccess) return cuda_fail(e, "shard setup exchange");
  return TSS_OK;
}

// Collective over the comm.  (1) Every rank learns every shard's row count: the K1-vs-K2 route
// of a sharded call must not depend on rank-local state, so it uses the SMALLEST shard.
// (2) Every rank's exchange buffer is mapped into this process (CUDA IPC) for the fused merge;
// the fused path is used only if EVERY rank mapped EVERY buffer (agreed by a second all-gather),
// otherwise all ranks use ncclAllGather + the merge kernel.  A rank with a local failure still
// takes part in both collectives, so its peers never hang in NCCL.
struct ShardHello {
  uint64_t n_rows;
  uint32_t have_handle;
  uint32_t pad;
  cudaIpcMemHandle_t handle;
};
int setup_shard_group(tss_index* ix, tss_comm* comm) {
  tss_index::Xchg& x = ix->xchg;
  x.ready = false;
  const char* env = getenv("TSS_FUSED_XCHG");
  const bool want_fused = !(env && atoi(env) == 0) && comm->nranks >= 2 &&
                          comm->nranks <= (int)tss::kXchgMaxRanks;
  if (x.comm != comm) {
    // another group: sequence numbers restart, so the arrival flags the old group left in this
    // buffer must go (every rank does this before the first collective below completes)
    x.seq = 0;
    for (int r = 0; r < 8; ++r) {
      if (x.peer[r] && x.peer[r] != x.local) cudaIpcCloseMemHandle(x.peer[r]);
      x.peer[r] = nullptr;
    }
    if (x.local) CU(cudaMemset(x.local, 0, tss::kXchgBytes));
    CU(cudaMemset(ix->d_counter + 17, 0, sizeof(unsigned int)));
    CU(cudaDeviceSynchronize());
    x.comm = nullptr;
  }
  ShardHello mine{};
  mine.n_rows = ix->n_rows;
  if (want_fused) {
    bool ok = true;
    if (!x.local) {
      ok = cudaMalloc(&x.local, tss::kXchgBytes) == cudaSuccess &&
           cudaMemset(x.local, 0, tss::kXchgBytes) == cudaSuccess &&
           cudaDeviceSynchronize() == cudaSuccess;
    }
    if (ok) ok = cudaIpcGetMemHandle(&mine.handle, x.local) == cudaSuccess;
    if (!ok) cudaGetLastError();
    mine.have_handle = ok ? 1u : 0u;
  }
  std::vector<ShardHello> all(comm->nranks);
  int rc = comm_allgather_host(ix, comm, &mine, all.data(), sizeof(ShardHello));
  if (rc) return rc;
  uint64_t min_rows = ~0ull;
  bool all_have = true;
  for (int r = 0; r < comm->nranks; ++r) {
    if (all[r].n_rows < min_rows) min_rows = all[r].n_rows;
    all_have = all_have && all[r].have_handle;
  }
  ix->shard_min_rows = min_rows;
  uint32_t mapped = 0;
  if (want_fused && all_have) {
    mapped = 1;
    for (int r = 0; r < comm->nranks; ++r) {
      if (r == comm->rank) {
        x.peer[r] = x.local;
        continue;
      }
      if (x.peer[r]) continue;  // still mapped from the previous attach of this comm
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        mapped = 0;
        break;
      }
      x.peer[r] = static_cast<uint8_t*>(ptr);
    }
  }
  std::vector<uint32_t> oks(comm->nranks);
  if ((rc = comm_allgather_host(ix, comm, &mapped, oks.data(), sizeof(uint32_t)))) return rc;
  bool ready = want_fused;
  for (int r = 0; r < comm->nranks; ++r) ready = ready && oks[r];
  x.comm = comm;
  x.ready = ready;
  return TSS_OK;
}
}  // namespace

int tss_index_set_batch_policy(tss_index* ix, uint32_t min_queries, int build_shadow_now) {
  if (!ix) return fail(TSS_ERR_INVALID_ARG, "index is NULL");
  std::lock_guard<std::mutex> lock(ix->mu);
  if (min_queries) {
    ix->gemm_min_nq = min_queries;
    ix->gemm_small_nq = 0xFFFFFFFFu;
  } else {
    ix->gemm_min_nq = 16;
    ix->gemm_small_nq = 3;
  }
  if (build_shadow_now) {  // fp32: the shadow it needs anyway; bf16: opt in to the unit-row copy
    if (!ix->finalized) return fail(TSS_ERR_STATE, "build the shadow after tss_index_finalize");
    if (!ix->n_rows) return TSS_OK;
    DeviceGuard g(ix->device);
    ix->gemm.shadow_failed = false;
    ix->gemm.unit_policy = 1;
    int rc = ensure_gemm_ws(ix, true);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ix->stream));
  }
  return TSS_OK;
}

int tss_index_set_shard(tss_index* ix, uint64_t row_base, tss_comm* comm) {
  if (!ix) return fail(TSS_ERR_INVALID_ARG, "index is NULL");
  std::lock_guard<std::mutex> lock(ix->mu);
  if (row_base + ix->n_rows >= 0xFFFFFFFFull)
    return fail(TSS_ERR_INVALID_ARG, "row ids are 32-bit");
  if (comm && comm->device != ix->device)
    return fail(TSS_ERR_INVALID_ARG, "comm and index live on different devices");
  if (comm && (uint64_t)comm->nranks * TSS_MAX_FUSED_K * 8 > 48 * 1024)
    return fail(TSS_ERR_INVALID_ARG, "at most %d ranks", 48 * 1024 / (TSS_MAX_FUSED_K * 8));
  ix->row_base = row_base;
  ix->comm = comm;
  ix->shard_min_rows = ix->n_rows;
  if (comm) {  // collective over the comm: every rank calls it with its own shard
    DeviceGuard g(ix->device);
    int rc = load_nccl();
    if (!rc) rc = setup_shard_group(ix, comm);
    if (rc) return rc;
  }
  return TSS_OK;
}

// ---- masks ---------------------------------------------------------------------------------
int tss_mask_create(tss_mask** out, uint64_t nbits, int device) {
  if (!out) return fail(TSS_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  int ndev = tss_device_count();
  if (ndev == 0) return fail(TSS_ERR_CUDA, "no CUDA device visible (libtss has no CPU path)");
  if (device < 0 || device >= ndev) return fail(TSS_ERR_INVALID_ARG, "device %d of %d", device, ndev);
  DeviceGuard g(device);
  tss_mask* m = new (std::nothrow) tss_mask();
  if (!m) return fail(TSS_ERR_OOM, "host allocation failed");
  m->device = device;
  m->nbits = nbits;
  m->nwords = (nbits + 31) / 32;
  // the scan reads whole words for the last (partial) tile: pad by one word
  cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->wev, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(&m->d_words, (size_t)(m->nwords + 1) * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&m->d_scratch, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&m->d_list, (size_t)kMaskListCap * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&m->d_list_count, sizeof(uint32_t));
  if (e == cudaSuccess)
    e = cudaMemsetAsync(m->d_words, 0, (size_t)(m->nwords + 1) * sizeof(uint32_t), m->stream);
  if (e == cudaSuccess) e = mask_end_write(m, m->stream);
  if (e != cudaSuccess) {
    tss_mask_destroy(m);
    return cuda_fail(e, "mask allocation");
  }
  *out = m;
  return TSS_OK;
}

void tss_mask_destroy(tss_mask* m) {
  if (!m) return;
  DeviceGuard g(m->device);
  // everything that was enqueued against the words must have run before they are freed
  if (m->has_write) cudaEventSynchronize(m->wev);
  for (auto& r : m->readers) {
    cudaEventSynchronize(r.ev);
    cudaEventDestroy(r.ev);
  }
  for (cudaEvent_t ev : m->spare) cudaEventDestroy(ev);
  if (m->stream) cudaStreamSynchronize(m->stream);
  cudaFree(m->d_words);
  cudaFree(m->d_scratch);
  cudaFree(m->d_list);
  cudaFree(m->d_list_count);
  if (m->wev) cudaEventDestroy(m->wev);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

uint64_t tss_mask_nbits(const tss_mask* m) { return m ? m->nbits : 0; }

int tss_mask_clear(tss_mask* m) {
  if (!m) return fail(TSS_ERR_INVALID_ARG, "mask is NULL");
  DeviceGuard g(m->device);
  CU(mask_begin_write(m, m->stream));
  CU(cudaMemsetAsync(m->d_words, 0, (size_t)(m->nwords + 1) * sizeof(uint32_t), m->stream));
  CU(mask_end_write(m, m->stream));
  return TSS_OK;
}

namespace {
int mask_update_rows(tss_mask* m, const uint32_t* rows, uint64_t n, uint64_t row_base, bool set) {
  if (!m) return fail(TSS_ERR_INVALID_ARG, "mask is NULL");
  if (!n) return TSS_OK;
  if (!rows) return fail(TSS_ERR_INVALID_ARG, "rows is NULL");
  DeviceGuard g(m->device);
  CU(mask_begin_write(m, m->stream));
  cudaError_t e = cudaSuccess;
  if (n <= tss::kInlineRows) {
    // a short list (the seen cases of one query) rides in the kernel parameters: no staging
    // allocation, no copy, no host synchronisation
    tss::InlineRows ir;
    ir.n = (uint32_t)n;
    memcpy(ir.rows, rows, (size_t)n * sizeof(uint32_t));
    e = tss::launch_mask_set_rows_inline(m->d_words, m->nbits, ir, row_base, set, m->stream);
  } else {
    uint32_t* d_rows = nullptr;
    CU(cudaMalloc(&d_rows, (size_t)n * sizeof(uint32_t)));
    e = cudaMemcpyAsync(d_rows, rows, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, m->stream);
    if (e == cudaSuccess)
      e = tss::launch_mask_set_rows(m->d_words, m->nbits, d_rows, n, row_base, set, m->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);  // d_rows is freed below
    cudaFree(d_rows);
  }
  if (e != cudaSuccess) return cuda_fail(e, "mask_set_rows");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CU(mask_end_write(m, m->stream));
  return TSS_OK;
}
}  // namespace

int tss_mask_set_rows(tss_mask* m, const uint32_t* rows, uint64_t n, uint64_t row_base) {
  return mask_update_rows(m, rows, n, row_base, true);
}
int tss_mask_clear_rows(tss_mask* m, const uint32_t* rows, uint64_t n, uint64_t row_base) {
  return mask_update_rows(m, rows, n, row_base, false);
}

// ---- N3: columnar metadata pre-filter ----------------------------------------------------------
int tss_columns_create(tss_columns** out, const uint16_t* court_ids, const int32_t* dates,
                       uint64_t nrows, int device) {
  if (!out) return fail(TSS_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (nrows && (!court_ids || !dates)) return fail(TSS_ERR_INVALID_ARG, "column pointer is NULL");
  int ndev = tss_device_count();
  if (ndev == 0) return fail(TSS_ERR_CUDA, "no CUDA device visible (libtss has no CPU path)");
  if (device < 0 || device >= ndev) return fail(TSS_ERR_INVALID_ARG, "device %d of %d", device, ndev);
  DeviceGuard g(device);
  tss_columns* c = new (std::nothrow) tss_columns();
  if (!c) return fail(TSS_ERR_OOM, "host allocation failed");
  c->device = device;
  c->nrows = nrows;
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_court, (nrows + 1) * sizeof(uint16_t));
  if (e == cudaSuccess) e = cudaMalloc(&c->d_date, (nrows + 1) * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&c->d_allow, 2048 * sizeof(uint32_t));
  if (e == cudaSuccess && nrows)
    e = cudaMemcpy(c->d_court, court_ids, nrows * sizeof(uint16_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && nrows)
    e = cudaMemcpy(c->d_date, dates, nrows * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();  // pageable H2D: see tss_terms_create
  if (e != cudaSuccess) {
    tss_columns_destroy(c);
    return cuda_fail(e, "columns upload");
  }
  *out = c;
  return TSS_OK;
}

void tss_columns_destroy(tss_columns* c) {
  if (!c) return;
  DeviceGuard g(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaFree(c->d_court);
  cudaFree(c->d_date);
  cudaFree(c->d_allow);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int tss_filter_mask(tss_columns* c, const uint16_t* allowed_courts, uint32_t n_allowed,
                    int32_t date_lo, int32_t date_hi, tss_mask* mask, int combine_and) {
  if (!c || !mask) return fail(TSS_ERR_INVALID_ARG, "NULL argument");
  if (n_allowed && !allowed_courts) return fail(TSS_ERR_INVALID_ARG, "allowed_courts is NULL");
  if (c->device != mask->device) return fail(TSS_ERR_INVALID_ARG, "columns and mask on different devices");
  if (mask->nbits < c->nrows)
    return fail(TSS_ERR_INVALID_ARG, "mask has %llu bits, columns have %llu rows",
                (unsigned long long)mask->nbits, (unsigned long long)c->nrows);
  DeviceGuard g(c->device);
  if (n_allowed) {
    std::vector<uint32_t> bits(2048, 0);
    for (uint32_t i = 0; i < n_allowed; ++i) bits[allowed_courts[i] >> 5] |= 1u << (allowed_courts[i] & 31);
    // (pageable source: the copy is staged before the call returns; in order on c->stream with
    // the previous call's kernel, which read d_allow)
    CU(cudaMemcpyAsync(c->d_allow, bits.data(), 2048 * sizeof(uint32_t), cudaMemcpyHostToDevice,
                       c->stream));
  }
  CU(mask_begin_write(mask, c->stream));
  cudaError_t e = tss::launch_filter_mask(c->d_court, c->d_date, c->nrows, c->d_allow, n_allowed == 0,
                                          date_lo, date_hi, mask->d_words, combine_and != 0, c->stream);
  if (e != cudaSuccess) return cuda_fail(e, "filter_mask");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CU(mask_end_write(mask, c->stream));
  return TSS_OK;
}

int tss_mask_upload(tss_mask* m, const uint32_t* words) {
  if (!m || !words) return fail(TSS_ERR_INVALID_ARG, "NULL argument");
  DeviceGuard g(m->device);
  CU(mask_begin_write(m, m->stream));
  CU(cudaMemcpyAsync(m->d_words, words, (size_t)m->nwords * sizeof(uint32_t), cudaMemcpyHostToDevice,
                     m->stream));
  CU(mask_end_write(m, m->stream));
  // the caller may reuse `words` right away
  CU(cudaStreamSynchronize(m->stream));
  return TSS_OK;
}

int tss_mask_download(const tss_mask* m, uint32_t* words) {
  if (!m || !words) return fail(TSS_ERR_INVALID_ARG, "NULL argument");
  DeviceGuard g(m->device);
  tss_mask* mm = const_cast<tss_mask*>(m);
  CU(mask_begin_read(mm, mm->stream));
  CU(cudaMemcpyAsync(words, m->d_words, (size_t)m->nwords * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                     mm->stream));
  CU(mask_end_read(mm, mm->stream));
  CU(cudaStreamSynchronize(mm->stream));
  return TSS_OK;
}

int tss_mask_popcount(const tss_mask* m, uint64_t* out) {
  if (!m || !out) return fail(TSS_ERR_INVALID_ARG, "NULL argument");
  DeviceGuard g(m->device);
  tss_mask* mm = const_cast<tss_mask*>(m);
  CU(mask_begin_read(mm, mm->stream));
  cudaError_t e = tss::launch_mask_popcount(m->d_words, m->nwords, m->d_scratch, mm->stream);
  if (e != cudaSuccess) return cuda_fail(e, "mask_popcount launch");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  unsigned long long v = 0;
  CU(cudaMemcpyAsync(&v, m->d_scratch, sizeof(v), cudaMemcpyDeviceToHost, mm->stream));
  CU(mask_end_read(mm, mm->stream));
  CU(cudaStreamSynchronize(mm->stream));
  *out = v;
  return TSS_OK;
}

// ---- flattened trie ---------------------------------------------------------------------------
int tss_terms_create(tss_terms** out, const char* pool, const uint64_t* term_off,
                     const uint64_t* post_off, const uint32_t* post_rows, uint64_t nterms,
                     int device) {
  if (!out) return fail(TSS_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (!term_off || !post_off) return fail(TSS_ERR_INVALID_ARG, "offset arrays are NULL");
  if (term_off[0] != 0 || post_off[0] != 0)
    return fail(TSS_ERR_INVALID_ARG, "offset arrays must start at 0");
  const uint64_t pool_bytes = term_off[nterms], nposts = post_off[nterms];
  if ((pool_bytes && !pool) || (nposts && !post_rows))
    return fail(TSS_ERR_INVALID_ARG, "pool / post_rows is NULL");
  // byte-sorted, strictly increasing (unique) terms: the kernel's binary search relies on it
  for (uint64_t i = 0; i < nterms; ++i) {
    if (term_off[i + 1] < term_off[i] || post_off[i + 1] < post_off[i])
      return fail(TSS_ERR_INVALID_ARG, "offsets not monotone at term %llu", (unsigned long long)i);
    if (i) {
      uint64_t la = term_off[i] - term_off[i - 1], lb = term_off[i + 1] - term_off[i];
      int c = memcmp(pool + term_off[i - 1], pool + term_off[i], la < lb ? la : lb);
      if (c > 0 || (c == 0 && la >= lb))
        return fail(TSS_ERR_INVALID_ARG, "terms not strictly byte-sorted at term %llu",
                    (unsigned long long)i);
    }
  }
  int ndev = tss_device_count();
  if (ndev == 0) return fail(TSS_ERR_CUDA, "no CUDA device visible (libtss has no CPU path)");
  if (device < 0 || device >= ndev) return fail(TSS_ERR_INVALID_ARG, "device %d of %d", device, ndev);
  DeviceGuard g(device);
  tss_terms* t = new (std::nothrow) tss_terms();
  if (!t) return fail(TSS_ERR_OOM, "host allocation failed");
  t->device = device;
  t->nterms = nterms;
  t->pool_bytes = pool_bytes;
  t->nposts = nposts;
  t->key_cap = 4096;
  cudaError_t e = cudaSuccess;
#define TRY(expr) \
  if (e == cudaSuccess) e = (expr);
  TRY(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking))
  TRY(cudaMalloc(&t->d_pool, pool_bytes + 16))
  TRY(cudaMalloc(&t->d_term_off, (nterms + 1) * sizeof(uint64_t)))
  TRY(cudaMalloc(&t->d_post_off, (nterms + 1) * sizeof(uint64_t)))
  TRY(cudaMalloc(&t->d_post_rows, (nposts + 1) * sizeof(uint32_t)))
  TRY(cudaMalloc(&t->d_keys, t->key_cap))
  TRY(cudaMallocHost(&t->h_keys, t->key_cap))
  TRY(cudaMalloc(&t->d_bounds, 8 * sizeof(uint64_t)))
  if (pool_bytes) TRY(cudaMemcpy(t->d_pool, pool, pool_bytes, cudaMemcpyHostToDevice))
  TRY(cudaMemcpy(t->d_term_off, term_off, (nterms + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice))
  TRY(cudaMemcpy(t->d_post_off, post_off, (nterms + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice))
  if (nposts) TRY(cudaMemcpy(t->d_post_rows, post_rows, nposts * sizeof(uint32_t), cudaMemcpyHostToDevice))
  // (a pageable H2D cudaMemcpy may return before its last DMA lands, and the non-blocking
  // streams that will read these arrays do not order against the legacy stream)
  TRY(cudaDeviceSynchronize())
#undef TRY
  t->own_stream = t->stream;
  if (e != cudaSuccess) {
    tss_terms_destroy(t);
    return cuda_fail(e, "terms upload");
  }
  *out = t;
  return TSS_OK;
}

// ---- N2: build the flattened trie on the device -------------------------------------------------
int tss_terms_build(tss_terms** out, const char* vocab_pool, const uint64_t* vocab_off,
                    uint32_t vocab_size, const uint32_t* token_ids, uint32_t max_tokens,
                    const uint32_t* rows, uint64_t n_postings, int device) {
  if (!out) return fail(TSS_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (!vocab_off || (vocab_size && !vocab_pool))
    return fail(TSS_ERR_INVALID_ARG, "vocabulary is NULL");
  if (n_postings && (!token_ids || !rows || !max_tokens))
    return fail(TSS_ERR_INVALID_ARG, "postings are NULL");
  if (n_postings >= 0x7FFFFFFFull) return fail(TSS_ERR_INVALID_ARG, "at most 2^31-1 postings per build");
  if (vocab_off[0] != 0) return fail(TSS_ERR_INVALID_ARG, "vocab_off must start at 0");
  // token ids follow byte order only if the vocabulary is byte-sorted, and tuple order equals
  // the byte order of the ' '-joined strings only if no token holds a byte <= ' '
  for (uint32_t i = 0; i < vocab_size; ++i) {
    if (vocab_off[i + 1] <= vocab_off[i])
      return fail(TSS_ERR_INVALID_ARG, "vocabulary token %u is empty or offsets decrease", i);
    for (uint64_t b = vocab_off[i]; b < vocab_off[i + 1]; ++b)
      if ((unsigned char)vocab_pool[b] <= 0x20)
        return fail(TSS_ERR_INVALID_ARG, "vocabulary token %u holds a byte <= 0x20", i);
    if (i) {
      uint64_t la = vocab_off[i] - vocab_off[i - 1], lb = vocab_off[i + 1] - vocab_off[i];
      int c = memcmp(vocab_pool + vocab_off[i - 1], vocab_pool + vocab_off[i], la < lb ? la : lb);
      if (c > 0 || (c == 0 && la >= lb))
        return fail(TSS_ERR_INVALID_ARG, "vocabulary not strictly byte-sorted at token %u", i);
    }
  }
  for (uint64_t i = 0; i < n_postings; ++i) {
    bool ended = false;
    for (uint32_t j = 0; j < max_tokens; ++j) {
      uint32_t id = token_ids[i * max_tokens + j];
      if (id > vocab_size) return fail(TSS_ERR_INVALID_ARG, "posting %llu: token id %u > vocabulary", (unsigned long long)i, id);
      if (id && ended) return fail(TSS_ERR_INVALID_ARG, "posting %llu: token after padding", (unsigned long long)i);
      if (!id) ended = true;
    }
  }
  int ndev = tss_device_count();
  if (ndev == 0) return fail(TSS_ERR_CUDA, "no CUDA device visible (libtss has no CPU path)");
  if (device < 0 || device >= ndev) return fail(TSS_ERR_INVALID_ARG, "device %d of %d", device, ndev);
  DeviceGuard g(device);
  tss_terms* t = new (std::nothrow) tss_terms();
  if (!t) return fail(TSS_ERR_OOM, "host allocation failed");
  t->device = device;
  t->key_cap = 4096;
  cudaError_t e = cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking);
  t->own_stream = t->stream;
  if (e == cudaSuccess) e = cudaMalloc(&t->d_keys, t->key_cap);
  if (e == cudaSuccess) e = cudaMallocHost(&t->h_keys, t->key_cap);
  if (e == cudaSuccess) e = cudaMalloc(&t->d_bounds, 8 * sizeof(uint64_t));
  tss::BuiltTerms b;
  if (e == cudaSuccess)
    e = tss::build_terms_device(vocab_pool, vocab_off, vocab_size, token_ids, max_tokens, rows,
                                n_postings, t->stream, &b);
  if (e != cudaSuccess) {
    tss_terms_destroy(t);
    return cuda_fail(e, "terms build");
  }
  g_launches.fetch_add(2 * max_tokens + 6, std::memory_order_relaxed);
  t->d_pool = b.d_pool;
  t->d_term_off = b.d_term_off;
  t->d_post_off = b.d_post_off;
  t->d_post_rows = b.d_post_rows;
  t->nterms = b.nterms;
  t->pool_bytes = b.pool_bytes;
  t->nposts = b.nposts;
  *out = t;
  return TSS_OK;
}

int tss_terms_build_text(tss_terms** out, const char* text, const uint64_t* phrase_off,
                         const uint32_t* rows, uint64_t n_phrases, int lowercase, uint32_t max_tokens,
                         int device) {
  if (!out) return fail(TSS_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (!phrase_off) return fail(TSS_ERR_INVALID_ARG, "phrase_off is NULL");
  if (phrase_off[0] != 0) return fail(TSS_ERR_INVALID_ARG, "phrase_off must start at 0");
  if (n_phrases && !rows) return fail(TSS_ERR_INVALID_ARG, "rows is NULL");
  if (!max_tokens || max_tokens > 64) return fail(TSS_ERR_INVALID_ARG, "max_tokens must be in [1,64]");
  for (uint64_t i = 0; i < n_phrases; ++i)
    if (phrase_off[i + 1] < phrase_off[i])
      return fail(TSS_ERR_INVALID_ARG, "phrase_off decreases at phrase %llu", (unsigned long long)i);
  const uint64_t total = phrase_off[n_phrases];
  if (total && !text) return fail(TSS_ERR_INVALID_ARG, "text is NULL");
  if (total >= 0x7FFFFFFFull || n_phrases >= 0x7FFFFFFFull || n_phrases * max_tokens >= 0xFFFFFFFFull)
    return fail(TSS_ERR_INVALID_ARG, "at most 2^31-1 text bytes / phrases (2^32-1 id slots) per build");
  int ndev = tss_device_count();
  if (ndev == 0) return fail(TSS_ERR_CUDA, "no CUDA device visible (libtss has no CPU path)");
  if (device < 0 || device >= ndev) return fail(TSS_ERR_INVALID_ARG, "device %d of %d", device, ndev);
  DeviceGuard g(device);
  tss_terms* t = new (std::nothrow) tss_terms();
  if (!t) return fail(TSS_ERR_OOM, "host allocation failed");
  t->device = device;
  t->key_cap = 4096;
  cudaError_t e = cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking);
  t->own_stream = t->stream;
  if (e == cudaSuccess) e = cudaMalloc(&t->d_keys, t->key_cap);
  if (e == cudaSuccess) e = cudaMallocHost(&t->h_keys, t->key_cap);
  if (e == cudaSuccess) e = cudaMalloc(&t->d_bounds, 8 * sizeof(uint64_t));
  tss::BuiltTerms b;
  uint32_t err_bits = 0, vocab = 0;
  uint64_t ntok = 0;
  if (e == cudaSuccess)
    e = tss::build_terms_from_text(text, phrase_off, rows, n_phrases, lowercase != 0, max_tokens,
                                   t->stream, &b, &err_bits, &ntok, &vocab);
  if (e != cudaSuccess) {
    tss_terms_destroy(t);
    return cuda_fail(e, "terms build from text");
  }
  if (err_bits) {
    tss_terms_destroy(t);
    return fail(TSS_ERR_INVALID_ARG, "text rejected:%s%s%s",
                (err_bits & 1) ? " a control byte below 0x20 that is not whitespace;" : "",
                (err_bits & 2) ? " a token longer than 128 bytes;" : "",
                (err_bits & 4) ? " a phrase with more than max_tokens tokens;" : "");
  }
  // tokenise + 2 scans, ceil(longest/8) chunk passes, dictionary, then the tuple build
  g_launches.fetch_add(12 + 2 * max_tokens + 6, std::memory_order_relaxed);
  t->d_pool = b.d_pool;
  t->d_term_off = b.d_term_off;
  t->d_post_off = b.d_post_off;
  t->d_post_rows = b.d_post_rows;
  t->nterms = b.nterms;
  t->pool_bytes = b.pool_bytes;
  t->nposts = b.nposts;
  *out = t;
  return TSS_OK;
}

// ---- N1, second half: the flattened term array on disk ------------------------------------------
// TrieIndex::save_to_disk / load_from_disk (reference src/trie.rs:83-94) are a no-op and a
// NotSupported stub; TrieConfig.index_path (src/config.rs:190) has nowhere to point.  The file is
// the four arrays exactly as they sit in HBM behind a 64-byte header, so loading is file ->
// pinned buffer -> HBM with no parsing: [header][term_off (T+1) u64][post_off (T+1) u64]
// [post_rows P u32][pool bytes], each section padded to 8 bytes.
namespace {
struct TermsFileHeader {
  char magic[8];  // "TSSTRM01"
  uint64_t nterms, pool_bytes, nposts;
  uint8_t pad[32];
};
static_assert(sizeof(TermsFileHeader) == 64, "header is 64 bytes");
inline uint64_t pad8(uint64_t n) { return (n + 7) & ~7ull; }

// device -> file / file -> device through one pinned bounce buffer
bool write_section(FILE* f, void* hbuf, const void* d_src, uint64_t bytes, uint64_t padded) {
  const uint8_t* src = static_cast<const uint8_t*>(d_src);
  for (uint64_t off = 0; off < bytes; off += kIoChunk) {
    size_t n = bytes - off < kIoChunk ? bytes - off : kIoChunk;
    if (cudaMemcpy(hbuf, src + off, n, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
    if (fwrite(hbuf, 1, n, f) != n) return false;
  }
  const uint8_t zeros[8] = {};
  return padded == bytes || fwrite(zeros, 1, padded - bytes, f) == padded - bytes;
}
bool read_section(FILE* f, void* hbuf, void* d_dst, uint64_t bytes, uint64_t padded) {
  uint8_t* dst = static_cast<uint8_t*>(d_dst);
  for (uint64_t off = 0; off < bytes; off += kIoChunk) {
    size_t n = bytes - off < kIoChunk ? bytes - off : kIoChunk;
    if (fread(hbuf, 1, n, f) != n) return false;
    if (cudaMemcpy(dst + off, hbuf, n, cudaMemcpyHostToDevice) != cudaSuccess) return false;
  }
  uint8_t skip[8];
  return padded == bytes || fread(skip, 1, padded - bytes, f) == padded - bytes;
}
}  // namespace

int tss_terms_save(const tss_terms* t, const char* path) {
  if (!t || !path) return fail(TSS_ERR_INVALID_ARG, "NULL argument");
  DeviceGuard g(t->device);
  CU(cudaStreamSynchronize(t->stream));
  FILE* f = fopen(path, "wb");
  if (!f) return fail(TSS_ERR_INVALID_ARG, "cannot open %s for writing", path);
  TermsFileHeader h{};
  memcpy(h.magic, "TSSTRM01", 8);
  h.nterms = t->nterms;
  h.pool_bytes = t->pool_bytes;
  h.nposts = t->nposts;
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  void* hbuf = nullptr;
  if (ok && cudaMallocHost(&hbuf, kIoChunk) != cudaSuccess) ok = false;
  const uint64_t ob = (t->nterms + 1) * 8;
  ok = ok && write_section(f, hbuf, t->d_term_off, ob, ob);
  ok = ok && write_section(f, hbuf, t->d_post_off, ob, ob);
  ok = ok && write_section(f, hbuf, t->d_post_rows, t->nposts * 4, pad8(t->nposts * 4));
  ok = ok && write_section(f, hbuf, t->d_pool, t->pool_bytes, pad8(t->pool_bytes));
  if (hbuf) cudaFreeHost(hbuf);
  ok = (fclose(f) == 0) && ok;
  if (!ok) return fail(TSS_ERR_STATE, "writing %s failed", path);
  return TSS_OK;
}

int tss_terms_load(tss_terms** out, const char* path, int device) {
  if (!out || !path) return fail(TSS_ERR_INVALID_ARG, "NULL argument");
  *out = nullptr;
  int ndev = tss_device_count();
  if (ndev == 0) return fail(TSS_ERR_CUDA, "no CUDA device visible (libtss has no CPU path)");
  if (device < 0 || device >= ndev) return fail(TSS_ERR_INVALID_ARG, "device %d of %d", device, ndev);
  FILE* f = fopen(path, "rb");
  if (!f) return fail(TSS_ERR_INVALID_ARG, "cannot open %s", path);
  TermsFileHeader h{};
  if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "TSSTRM01", 8) != 0) {
    fclose(f);
    return fail(TSS_ERR_INVALID_ARG, "%s is not a TSSTRM01 file", path);
  }
  // the sizes must add up to the file before anything is allocated from them
  const uint64_t ob = (h.nterms + 1) * 8;
  const uint64_t want = 64 + 2 * ob + pad8(h.nposts * 4) + pad8(h.pool_bytes);
  bool sane = h.nterms < (1ull << 40) && h.nposts < (1ull << 40) && h.pool_bytes < (1ull << 44) &&
              fseek(f, 0, SEEK_END) == 0 && (uint64_t)ftell(f) == want &&
              fseek(f, 64, SEEK_SET) == 0;
  if (!sane) {
    fclose(f);
    return fail(TSS_ERR_INVALID_ARG, "%s: header and file size disagree (truncated or corrupt)", path);
  }
  DeviceGuard g(device);
  tss_terms* t = new (std::nothrow) tss_terms();
  if (!t) {
    fclose(f);
    return fail(TSS_ERR_OOM, "host allocation failed");
  }
  t->device = device;
  t->nterms = h.nterms;
  t->pool_bytes = h.pool_bytes;
  t->nposts = h.nposts;
  t->key_cap = 4096;
  cudaError_t e = cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking);
  t->own_stream = t->stream;
  void* hbuf = nullptr;
  if (e == cudaSuccess) e = cudaMalloc(&t->d_pool, h.pool_bytes + 16);
  if (e == cudaSuccess) e = cudaMalloc(&t->d_term_off, ob);
  if (e == cudaSuccess) e = cudaMalloc(&t->d_post_off, ob);
  if (e == cudaSuccess) e = cudaMalloc(&t->d_post_rows, (h.nposts + 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&t->d_keys, t->key_cap);
  if (e == cudaSuccess) e = cudaMallocHost(&t->h_keys, t->key_cap);
  if (e == cudaSuccess) e = cudaMalloc(&t->d_bounds, 8 * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMallocHost(&hbuf, kIoChunk);
  if (e != cudaSuccess) {
    fclose(f);
    if (hbuf) cudaFreeHost(hbuf);
    tss_terms_destroy(t);
    return cuda_fail(e, "terms load allocation");
  }
  bool ok = read_section(f, hbuf, t->d_term_off, ob, ob) && read_section(f, hbuf, t->d_post_off, ob, ob) &&
            read_section(f, hbuf, t->d_post_rows, h.nposts * 4, pad8(h.nposts * 4)) &&
            read_section(f, hbuf, t->d_pool, h.pool_bytes, pad8(h.pool_bytes));
  cudaFreeHost(hbuf);
  fclose(f);
  // what tss_terms_create validates on the host is validated here on the device: offsets
  // monotone and closing on the header's totals, terms strictly byte-sorted
  unsigned int bad = 1;
  if (ok) {
    unsigned int* d_bad = reinterpret_cast<unsigned int*>(t->d_bounds);
    ok = cudaMemsetAsync(d_bad, 0, sizeof(unsigned int), t->stream) == cudaSuccess;
    tss::TermsDev td{t->d_pool, t->d_term_off, t->d_post_off, t->d_post_rows, t->nterms};
    ok = ok && tss::launch_terms_validate(td, h.pool_bytes, h.nposts, d_bad, t->stream) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, t->stream) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(t->stream) == cudaSuccess;
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  if (!ok || bad) {
    tss_terms_destroy(t);
    return fail(TSS_ERR_INVALID_ARG, "%s: %s", path,
                ok ? "offsets not monotone or terms not strictly byte-sorted" : "read failed");
  }
  *out = t;
  return TSS_OK;
}

int tss_terms_sizes(const tss_terms* t, uint64_t* nterms, uint64_t* pool_bytes, uint64_t* npostings) {
  if (!t) return fail(TSS_ERR_INVALID_ARG, "terms is NULL");
  if (nterms) *nterms = t->nterms;
  if (pool_bytes) *pool_bytes = t->pool_bytes;
  if (npostings) *npostings = t->nposts;
  return TSS_OK;
}

int tss_terms_export(const tss_terms* t, char* pool, uint64_t* term_off, uint64_t* post_off,
                     uint32_t* post_rows) {
  if (!t || !term_off || !post_off) return fail(TSS_ERR_INVALID_ARG, "NULL argument");
  if ((t->pool_bytes && !pool) || (t->nposts && !post_rows))
    return fail(TSS_ERR_INVALID_ARG, "NULL argument");
  DeviceGuard g(t->device);
  CU(cudaStreamSynchronize(t->stream));
  if (t->pool_bytes) CU(cudaMemcpy(pool, t->d_pool, t->pool_bytes, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(term_off, t->d_term_off, (t->nterms + 1) * 8, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(post_off, t->d_post_off, (t->nterms + 1) * 8, cudaMemcpyDeviceToHost));
  if (t->nposts) CU(cudaMemcpy(post_rows, t->d_post_rows, t->nposts * 4, cudaMemcpyDeviceToHost));
  return TSS_OK;
}

uint64_t tss_terms_size(const tss_terms* t) { return t ? t->nterms : 0; }

void tss_terms_destroy(tss_terms* t) {
  if (!t) return;
  DeviceGuard g(t->device);
  if (t->stream) cudaStreamSynchronize(t->stream);
  if (t->own_stream && t->own_stream != t->stream) cudaStreamSynchronize(t->own_stream);
  cudaFree(t->d_pool);
  cudaFree(t->d_term_off);
  cudaFree(t->d_post_off);
  cudaFree(t->d_post_rows);
  cudaFree(t->d_keys);
  cudaFree(t->d_bounds);
  if (t->h_keys) cudaFreeHost(t->h_keys);
  if (t->own_stream) cudaStreamDestroy(t->own_stream);
  delete t;
}

namespace {
// clear_first: the output mask is zeroed by the search kernel's spare CTAs (one launch less than
// tss_mask_clear + tss_prefix_mask, and no cross-stream hop)
int prefix_mask_impl(tss_terms* t, const char* prefix, uint32_t len, int kind, tss_mask* out,
                     uint64_t row_base, tss_prefix_stats* stats, bool clear_first) {
  if (!t || !out) return fail(TSS_ERR_INVALID_ARG, "NULL argument");
  if (len && !prefix) return fail(TSS_ERR_INVALID_ARG, "prefix is NULL");
  if (kind != TSS_PREFIX_TOKEN && kind != TSS_PREFIX_CHAR)
    return fail(TSS_ERR_INVALID_ARG, "kind %d is not a TSS_PREFIX_* value", kind);
  if (t->device != out->device) return fail(TSS_ERR_INVALID_ARG, "terms and mask on different devices");
  if (4 * (len + 1) > t->key_cap) return fail(TSS_ERR_INVALID_ARG, "prefix longer than %u bytes", t->key_cap / 4 - 1);
  DeviceGuard g(t->device);
  // probes: bounds[0..1) = exact range, bounds[2..3) = subtree range
  tss::PrefixKeys keys{};
  std::string kb;
  auto push = [&](int i, const std::string& s) {
    keys.off[i] = (uint32_t)kb.size();
    kb += s;
    keys.off[i + 1] = (uint32_t)kb.size();
    keys.fixed[i] = -1;
  };
  auto fixed = [&](int i, int32_t v) {
    keys.off[i] = (uint32_t)kb.size();
    keys.off[i + 1] = (uint32_t)kb.size();
    keys.fixed[i] = v;
  };
  const std::string p(prefix ? prefix : "", len);
  if (kind == TSS_PREFIX_TOKEN) {
    if (len == 0) {
      // zero tokens: node = root.  The root is terminal only if an empty token list was
      // inserted (term ""), which sorts first; everything else is its subtree.
      fixed(0, 0);
      push(1, std::string(1, '\0'));
      push(2, std::string(1, '\0'));
      fixed(3, -2);
    } else {
      push(0, p);                      // first term >= P
      push(1, p + std::string(1, '\0'));  // first term > P
      push(2, p + " ");                // first term with a further token
      push(3, p + "!");                // ' ' + 1
    }
  } else {
    fixed(0, 0);
    fixed(1, 0);
    if (len == 0) {
      fixed(2, 0);
      fixed(3, -2);
    } else {
      push(2, p);
      std::string succ = p;  // smallest string greater than every string prefixed by P
      while (!succ.empty() && (unsigned char)succ.back() == 0xFF) succ.pop_back();
      if (succ.empty()) {
        fixed(3, -2);
      } else {
        succ.back() = (char)((unsigned char)succ.back() + 1);
        push(3, succ);
      }
    }
  }
  cudaStream_t st = t->stream;
  const char* d_keybytes = nullptr;
  if (kb.size() <= tss::kPrefixInlineBytes) {
    memcpy(keys.bytes, kb.data(), kb.size());  // the keys travel in the kernel parameters
  } else {
    CU(cudaStreamSynchronize(st));  // long prefix: the pinned staging buffer is reused
    memcpy(t->h_keys, kb.data(), kb.size());
    CU(cudaMemcpyAsync(t->d_keys, t->h_keys, kb.size(), cudaMemcpyHostToDevice, st));
    d_keybytes = t->d_keys;
  }
  CU(mask_begin_write(out, st));
  if (!t->num_sms) {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, t->device));
    t->num_sms = prop.multiProcessorCount;
    CU(cudaMemsetAsync(t->d_bounds, 0, 8 * sizeof(uint64_t), st));  // incl. the barrier words
  }
  tss::PrefixMaskArgs pa{};
  pa.t = tss::TermsDev{t->d_pool, t->d_term_off, t->d_post_off, t->d_post_rows, t->nterms};
  pa.d_keybytes = d_keybytes;
  pa.bounds = t->d_bounds;
  pa.words = out->d_words;
  pa.clear_nwords = clear_first ? out->nwords + 1 : 0;
  pa.nbits = out->nbits;
  pa.row_base = row_base;
  // only a scatter into a freshly cleared mask knows the mask's complete row set
  pa.list = clear_first ? out->d_list : nullptr;
  pa.list_count = clear_first ? out->d_list_count : nullptr;
  pa.list_cap = kMaskListCap;
  pa.sync = reinterpret_cast<unsigned int*>(t->d_bounds + 6);
  cudaError_t e = tss::launch_prefix_mask(pa, keys, t->num_sms, st);
  if (e != cudaSuccess) return cuda_fail(e, "prefix_mask_kernel launch");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  // consumers on other streams are ordered behind the scatter by the mask's write event:
  // no host synchronisation unless the caller wants the statistics
  CU(mask_end_write(out, st));
  out->list_valid = clear_first;
  if (stats) {
    uint64_t h[5];
    CU(cudaMemcpyAsync(h, t->d_bounds, sizeof(h), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    stats->exact_lo = h[0];
    stats->exact_hi = h[1] > h[0] ? h[1] : h[0];
    stats->sub_lo = h[2];
    stats->sub_hi = h[3] > h[2] ? h[3] : h[2];
    stats->npostings = h[4];
  }
  return TSS_OK;
}
}  // namespace

int tss_prefix_mask(tss_terms* t, const char* prefix, uint32_t len, int kind, tss_mask* out,
                    uint64_t row_base, tss_prefix_stats* stats) {
  return prefix_mask_impl(t, prefix, len, kind, out, row_base, stats, false);
}
int tss_prefix_mask_fresh(tss_terms* t, const char* prefix, uint32_t len, int kind, tss_mask* out,
                          uint64_t row_base, tss_prefix_stats* stats) {
  return prefix_mask_impl(t, prefix, len, kind, out, row_base, stats, true);
}

int tss_terms_bind_stream(tss_terms* t, tss_index* ix) {
  if (!t) return fail(TSS_ERR_INVALID_ARG, "terms is NULL");
  if (ix && ix->device != t->device) return fail(TSS_ERR_INVALID_ARG, "terms and index on different devices");
  DeviceGuard g(t->device);
  CU(cudaStreamSynchronize(t->stream));  // d_bounds / d_keys change streams: drain the old one
  t->stream = ix ? ix->stream : t->own_stream;
  return TSS_OK;
}

// ---- plumbing --------------------------------------------------------------------------------
int tss_index_debug_phases(tss_index* ix, void* d_buf) {
  if (!ix) return fail(TSS_ERR_INVALID_ARG, "index is NULL");
  ix->d_dbg = reinterpret_cast<unsigned long long*>(d_buf);
  return TSS_OK;
}
void* tss_index_stream(tss_index* ix) { return ix ? (void*)ix->stream : nullptr; }

int tss_index_sync(tss_index* ix) {
  if (!ix) return fail(TSS_ERR_INVALID_ARG, "index is NULL");
  std::lock_guard<std::mutex> lock(ix->mu);
  DeviceGuard g(ix->device);
  CU(cudaStreamSynchronize(ix->stream));
  if (ix->h_status && *ix->h_status) {  // a fused sharded search gave up waiting for a peer
    *ix->h_status = 0;
    return fail(TSS_ERR_NCCL, "a rank of the shard group did not deliver its top-k within 20 s");
  }
  if (ix->h_status && ix->h_status[1]) {
    // a device-resident tensor-core batch had more overflowing queries than fix-up launches
    ix->h_status[1] = 0;
    if (ix->gemm.fixups < 64) ix->gemm.fixups *= 2;
    return fail(TSS_ERR_STATE,
                "a tss_index_search_device batch flagged more queries for an exact redo than its "
                "device-side fix-up launches cover: repeat the call (the number of fix-ups has "
                "been raised to %u) or use tss_index_search",
                ix->gemm.fixups);
  }
  return TSS_OK;
}

int tss_dev_alloc(int device, uint64_t bytes, void** out) {
  if (!out) return fail(TSS_ERR_INVALID_ARG, "out is NULL");
  DeviceGuard g(device);
  if (!g.ok) return fail(TSS_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  CU(cudaMalloc(out, bytes ? bytes : 1));
  return TSS_OK;
}
int tss_dev_free(int device, void* p) {
  DeviceGuard g(device);
  CU(cudaFree(p));
  return TSS_OK;
}
int tss_dev_h2d(int device, void* dst, const void* src, uint64_t bytes) {
  DeviceGuard g(device);
  CU(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  CU(cudaDeviceSynchronize());  // pageable H2D may return before the last DMA lands
  return TSS_OK;
}
int tss_dev_d2h(int device, void* dst, const void* src, uint64_t bytes) {
  DeviceGuard g(device);
  CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return TSS_OK;
}
int tss_event_create(int device, void** out) {
  if (!out) return fail(TSS_ERR_INVALID_ARG, "out is NULL");
  DeviceGuard g(device);
  cudaEvent_t ev;
  CU(cudaEventCreate(&ev));
  *out = ev;
  return TSS_OK;
}
int tss_event_record(tss_index* ix, void* ev) {
  if (!ix || !ev) return fail(TSS_ERR_INVALID_ARG, "NULL argument");
  DeviceGuard g(ix->device);
  CU(cudaEventRecord((cudaEvent_t)ev, ix->stream));
  return TSS_OK;
}
int tss_event_elapsed_ms(void* a, void* b, float* out_ms) {
  if (!a || !b || !out_ms) return fail(TSS_ERR_INVALID_ARG, "NULL argument");
  CU(cudaEventSynchronize((cudaEvent_t)b));
  CU(cudaEventElapsedTime(out_ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return TSS_OK;
}
int tss_event_destroy(void* ev) {
  if (ev) CU(cudaEventDestroy((cudaEvent_t)ev));
  return TSS_OK;
}

}  // extern "C"
