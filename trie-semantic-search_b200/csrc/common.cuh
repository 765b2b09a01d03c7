// common.cuh -- shared device helpers for libtss (sm_100a only).
//
// Ordering key, score rule and the canonical fp32 reduction are the ones
// DESIGN.md section 3 defines and oracle/oracle.cpp restates on the CPU; the two
// must stay in lock-step, the parity tests compare them bit for bit.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tss {

constexpr unsigned FULL_MASK = 0xFFFFFFFFu;

// ---- ordering key: (orderable(score) << 32) | ~row; 0 = empty slot ----------
__host__ __device__ __forceinline__ uint32_t orderable_bits(uint32_t u) {
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t pack_key(float score, uint32_t row) {
  return ((uint64_t)orderable_bits(__float_as_uint(score)) << 32) |
         (uint64_t)(0xFFFFFFFFu - row);
}

// ---- score rule: zero / non-finite denominators and quotients score 0.0,
// -0.0 is canonicalised.  Explicit _rn intrinsics: never contracted by nvcc. ---
__device__ __forceinline__ float finish_score(float dot, float sqrt_nq2, float ne2) {
  float den = __fmul_rn(sqrt_nq2, __fsqrt_rn(ne2));
  float s = __fdiv_rn(dot, den);
  if (!(den > 0.0f) || !isfinite(s)) s = 0.0f;
  if (s == 0.0f) s = 0.0f;
  return s;
}

// ---- packed fp32 FMA (SASS FFMA2): two independent round-to-nearest fmas in one instruction,
// bit-identical to two __fmaf_rn calls, half the issue slots ------------------------------------
__device__ __forceinline__ void fma2(float ax, float ay, float bx, float by, float& cx, float& cy) {
  unsigned long long ra, rb, rc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(ax), "f"(ay));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(bx), "f"(by));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(cx), "f"(cy));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rc) : "l"(ra), "l"(rb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(cx), "=f"(cy) : "l"(rc));
}

// ---- transposed butterfly over RG rows ------------------------------------------
// v[r] is this lane's partial for row r.  On return v[0] holds the full 32-lane
// sum for row (lane >> (5 - log2 RG)); the addition tree per row is exactly
// the xor-butterfly 16,8,4,2,1 (fp add is commutative, so which lane of a pair
// "keeps" a row does not change any bit).
template <int RG>
__device__ __forceinline__ float reduce_rows(float (&v)[RG], int lane) {
  int m = 16;
#pragma unroll
  for (int c = RG; c > 1; c >>= 1) {
    const int half = c >> 1;
    const bool upper = (lane & m) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      float send = upper ? v[i] : v[i + half];
      float keep = upper ? v[i + half] : v[i];
      v[i] = __fadd_rn(keep, __shfl_xor_sync(FULL_MASK, send, m));
    }
    m >>= 1;
  }
#pragma unroll
  for (; m >= 1; m >>= 1) v[0] = __fadd_rn(v[0], __shfl_xor_sync(FULL_MASK, v[0], m));
  return v[0];
}

__device__ __forceinline__ float butterfly_sum(float x) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) x = __fadd_rn(x, __shfl_xor_sync(FULL_MASK, x, m));
  return x;
}

// ---- mbarrier + bulk async copy (TMA, 1-D) ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
// raise the pending transaction count of the current phase without arriving
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
// Completion is signalled on the mbarrier as complete_tx(bytes).  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                         uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}

}  // namespace tss
