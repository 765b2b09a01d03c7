// scan_launch.h -- host-side dispatch for the K1 scan kernel instances.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tss {
struct ScanParams;

// supported storage stripe counts (row = NS * 128 elements, zero padded)
inline int storage_stripes_for_dim(uint32_t dim) {
  static const int kNs[] = {1, 2, 3, 4, 6, 8};
  uint32_t need = (dim + 127) / 128;
  for (int ns : kNs)
    if ((uint32_t)ns >= need) return ns;
  return 0;  // dim > 1024 unsupported
}

// per-warp candidate list geometry for a given k
inline uint32_t kp_for_k(uint32_t k) {
  uint32_t kp = 8;
  while (kp < k) kp <<= 1;
  return kp;
}
inline uint32_t cap_for_k(uint32_t k) {
  uint32_t kp = kp_for_k(k);
  return 2 * (kp < 32 ? 32 : kp);
}
// queries one launch may carry for this k (shared-memory budget of the lists)
inline int max_bq_for_k(uint32_t k) { return cap_for_k(k) <= 64 ? 4 : 1; }

#define TSS_DECL_SCAN(NS)                                                                   \
  cudaError_t launch_scan_ns##NS(const ScanParams& p, int bq, bool bf16, bool masked, int grid, \
                                 int device, cudaStream_t st);
TSS_DECL_SCAN(1)
TSS_DECL_SCAN(2)
TSS_DECL_SCAN(3)
TSS_DECL_SCAN(4)
TSS_DECL_SCAN(6)
TSS_DECL_SCAN(8)
#undef TSS_DECL_SCAN

cudaError_t launch_scan(int ns, const ScanParams& p, int bq, bool bf16, bool masked, int grid,
                        int device, cudaStream_t st);
}  // namespace tss
