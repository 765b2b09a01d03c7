#!/bin/bash
# K2 knob sweeps: each argument is "ENV=.. ENV=.. -- bench args"; prints ms per batch at 10M x 384
run() {
  local envs=() args=() seen=0
  for w in $1; do if [ "$w" = "--" ]; then seen=1; elif [ $seen = 0 ]; then envs+=("$w"); else args+=("$w"); fi; done
  echo -n "$1: "
  env "${envs[@]}" timeout 100 python benchmarks/gemm_bench.py --iters 5 "${args[@]}" 2>&1 | tail -1 |
    python -c "import sys,json; print(json.loads(sys.stdin.read())['ms_per_batch'])"
}
for a in "$@"; do run "$a"; done
