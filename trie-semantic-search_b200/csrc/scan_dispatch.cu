#include "scan.cuh"
#include "scan_launch.h"

namespace tss {
cudaError_t launch_scan(int ns, const ScanParams& p, int bq, bool bf16, bool masked, int grid,
                        int device, cudaStream_t st) {
  switch (ns) {
    case 1: return launch_scan_ns1(p, bq, bf16, masked, grid, device, st);
    case 2: return launch_scan_ns2(p, bq, bf16, masked, grid, device, st);
    case 3: return launch_scan_ns3(p, bq, bf16, masked, grid, device, st);
    case 4: return launch_scan_ns4(p, bq, bf16, masked, grid, device, st);
    case 6: return launch_scan_ns6(p, bq, bf16, masked, grid, device, st);
    case 8: return launch_scan_ns8(p, bq, bf16, masked, grid, device, st);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace tss
