#!/bin/bash
# K2 time dissection: cluster mode (2 pair, 1 single) x debug flags that keep the result
# valid enough not to trigger the exact-scan fallback (1 = no epilogue math, 5 = MMA only, 3 = TMA only)
for cl in ${CLUSTERS:-2 1}; do for d in ${DEBUGS:-0 1 5 3}; do
echo -n "cluster=$cl debug=$d: "; TSS_GEMM_CLUSTER=$cl TSS_GEMM_DEBUG=$d timeout 100 python benchmarks/gemm_bench.py --iters 5 2>&1 | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['ms_per_batch'])"
done; done
