#!/bin/bash
# K2 knob sweeps (pair mode): each line is `env knobs...: ms per 1024-query batch at 10M x 384`
run() { echo -n "$*: "; env "$@" timeout 100 python benchmarks/gemm_bench.py --iters 5 2>&1 | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['ms_per_batch'])"; }
for a in "$@"; do run $a; done
