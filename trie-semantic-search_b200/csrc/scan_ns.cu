// scan_ns.cu -- instantiates the K1 scan kernel for one stripe count.
// Compiled once per supported NS with -DTSS_NS=<n> (see Makefile) so the
// instances build in parallel.
#include <mutex>

#include "scan.cuh"
#include "scan_launch.h"

#ifndef TSS_NS
#error "compile with -DTSS_NS=<stripes>"
#endif

namespace tss {

constexpr int kWarps = 16;
constexpr int kMaxDevices = 64;

template <int NS, int BQ, bool BF16, bool MASKED>
static cudaError_t launch_one(const ScanParams& p, int grid, int device, cudaStream_t st) {
  auto kern = scan_topk_kernel<NS, BQ, kWarps, BF16, MASKED>;
  size_t smem = (size_t)kWarps * TileGeom<NS, BF16>::TILE_BYTES +
                (size_t)kWarps * BQ * p.cap * 8 + (size_t)kWarps * 8 +
                (size_t)kWarps * 16 * 4;  // + per-warp slot -> row table of the masked scan
  // the last CTA stages gridDim.x partial lists of kp keys in shared memory
  if (grid > 256) return cudaErrorInvalidConfiguration;
  // (+ the staging rows and peer lists of the fused sharded merge)
  const size_t merge_bytes = (size_t)grid * p.kp * 8 + (size_t)kXchgMaxQ * kXchgMaxK * 8;
  const size_t xchg_bytes = (size_t)kXchgMaxQ * kXchgMaxRanks * kXchgMaxK * 8;
  if (smem < merge_bytes) smem = merge_bytes;
  if (smem < xchg_bytes) smem = xchg_bytes;
  // 227 KB per CTA minus the kernel's static shared memory (ticket word, padded)
  if (smem > 232448 - 256) return cudaErrorInvalidConfiguration;
  // per-instance cache of the attribute already set on each device; distinct index handles may
  // launch from distinct threads (include/tss.h), so the cache is guarded
  static size_t attr_bytes[kMaxDevices] = {};
  static std::mutex attr_mu;
  if (device < 0 || device >= kMaxDevices) return cudaErrorInvalidDevice;
  {
    std::lock_guard<std::mutex> lk(attr_mu);
    if (attr_bytes[device] < smem) {
      cudaError_t e =
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      attr_bytes[device] = smem;
    }
  }
  ScanParams q = p;
  q.smem_bytes = (uint32_t)smem;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kWarps * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = q.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, q);
}

template <int NS, int BQ>
static cudaError_t launch_bq(const ScanParams& p, bool bf16, bool masked, int grid, int device,
                             cudaStream_t st) {
  if (bf16)
    return masked ? launch_one<NS, BQ, true, true>(p, grid, device, st)
                  : launch_one<NS, BQ, true, false>(p, grid, device, st);
  return masked ? launch_one<NS, BQ, false, true>(p, grid, device, st)
                : launch_one<NS, BQ, false, false>(p, grid, device, st);
}

#define TSS_CAT2(a, b) a##b
#define TSS_CAT(a, b) TSS_CAT2(a, b)

cudaError_t TSS_CAT(launch_scan_ns, TSS_NS)(const ScanParams& p, int bq, bool bf16, bool masked,
                                            int grid, int device, cudaStream_t st) {
  switch (bq) {
    case 1: return launch_bq<TSS_NS, 1>(p, bf16, masked, grid, device, st);
    case 2: return launch_bq<TSS_NS, 2>(p, bf16, masked, grid, device, st);
    case 4: return launch_bq<TSS_NS, 4>(p, bf16, masked, grid, device, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace tss
