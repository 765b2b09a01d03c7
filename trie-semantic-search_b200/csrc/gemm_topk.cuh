// gemm_topk.cuh -- host-visible interface of the K2 tensor-core path (see gemm_topk.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tss {

struct GemmParams {
  uint64_t n_rows;         // rows in this shard
  uint32_t row_base;       // global id of local row 0
  const float* inv_norm;   // [n_rows] 1/|row|; null: the rows the tensor cores read are unit
                           // vectors and there is no mask -- the epilogue applies no weight
  const uint32_t* mask;    // row mask words (bit r&31 of word r>>5 <-> local row r) or null
  int mask_mode;           // TSS_MASK_*: masked rows get 1/|row| = NaN, which fmax and >= ignore
  uint32_t mb;             // 128-query blocks in this launch (grid = nslices * mb)
  uint32_t num_tiles;      // ceil(n_rows / 256)
  int mode;                // 0 = per-tile maxima over the sample, 1 = collect survivors
  uint32_t sample_stride, sample_count;  // mode 0: tiles j * stride, j < count
  float* tile_max;         // mode 0 out: [mb*128][sample_count*split] (one maximum per tile part)
  const float* thr;        // mode 1 in:  [mb*128]
  uint64_t* cand;          // mode 1 out: [mb*128][nslices*split][cand_cap] keys (unscaled by 1/|q|)
  uint32_t* cand_count;    // [mb*128][nslices*split] survivors seen (may exceed cand_cap)
  uint32_t cand_cap;       // per (query, slice, column part) list
  uint32_t ring_stages;    // 0 = all the stages the kernel has; fewer for latency experiments
  uint32_t debug;          // diagnostics: 1 = no epilogue math, 2 = no MMA issue, 4 = no corpus TMA
};

// How the CTAs of a launch cooperate.  PAIR needs an even mb and a tmap_e with a 128-row box:
// two CTAs score different query blocks against the same corpus tiles as one cta_group::2 M256
// MMA, each staging half of every tile.  SINGLE: independent CTAs, 256-row box.
// QUAD: clusters of four = two pairs sharing corpus tiles by TMA multicast (mb a multiple of 4,
// tmap_e with a 64-row box); at most gemm_max_quads(kb) clusters are resident at once.
enum GemmCluster { TSS_GEMM_SINGLE = 1, TSS_GEMM_PAIR = 2, TSS_GEMM_QUAD = 4 };
size_t gemm_smem_bytes(int kb, bool pair);
int gemm_max_quads(int kb);
int gemm_col_split();  // survivor lists / threshold samples per (slice, tile)
cudaError_t launch_gemm_topk(int kb, int cluster, const CUtensorMap& tmap_q,
                             const CUtensorMap& tmap_e, const GemmParams& p, int grid,
                             cudaStream_t st);
// margin: [nq_pad] out, the rescoring margin of every query (see prep_queries_kernel);
// shadow_err: null when the tensor cores read a bf16 index's own rows, else the device word
// launch_normalize_rows left (how far rounding moved any unit row of the shadow)
cudaError_t launch_prep_queries(const float* q, uint32_t nq, uint32_t dim, uint32_t kpad,
                                uint32_t nq_pad, uint16_t* out, float* inv_qnorm, float* margin,
                                const unsigned int* shadow_err, cudaStream_t st);
cudaError_t launch_row_inv_norm(const void* rows, uint64_t n_rows, uint32_t stride_elems, float* out,
                                cudaStream_t st);
// margin: null, or what to subtract from every threshold (survivors will be re-scored)
cudaError_t launch_threshold(const float* tile_max, uint32_t sample_count, uint32_t nq_pad,
                             uint32_t nq, uint32_t k, const float* margin, float* thr,
                             cudaStream_t st);
// queries != null: re-score the keys within the margin of the k-th best with the scan's arithmetic
// (fp32 queries [nq][dim]; rows = the stored matrix, bf16 or fp32, stride_elems per row); null:
// top-k of the tensor-core scores
cudaError_t launch_select(const uint64_t* cand, const uint32_t* cand_count, uint32_t nslices,
                          uint32_t cap_s, const float* inv_qnorm, uint32_t nq, uint32_t k,
                          const float* queries, const float* margin, const void* rows,
                          bool rows_f32, uint32_t dim, uint32_t stride_elems, uint32_t row_base,
                          uint64_t n_rows, uint64_t* out, uint32_t* overflow, cudaStream_t st);

// Shadow prefilter (single queries, fp32 index): cand [nq][kc] = the scan's top-kc over the bf16
// shadow; re-scores them from the fp32 rows and writes the top-k, or flags the query incomplete
// (the kc-th shadow score is not provably below the fp32 top-k): see refine_kernel.
cudaError_t launch_refine(const uint64_t* cand, uint32_t kc, const float* queries, const void* rows_f32,
                          uint32_t dim, uint32_t stride_elems, uint32_t row_base, uint32_t nq,
                          uint64_t n_rows, const unsigned int* shadow_err, uint32_t k, uint64_t* out,
                          uint32_t* incomplete, cudaStream_t st);

}  // namespace tss
