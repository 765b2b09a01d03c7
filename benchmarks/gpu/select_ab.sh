# same-box A/B of select_kernel builds: prev (walks the survivor lists in global memory for every
# radix pass), new (survivors staged in shared memory once; a lane per list when lists are many
# and short), minb4 (new with __launch_bounds__(256, 4): 64 registers, no spills)
for lib in prev new minb4; do
  if [ $lib = new ]; then unset TSS_LIB_PATH; else export TSS_LIB_PATH=$PWD/benchmarks/gpu/ab/libtss_$lib.so; fi
  for cfg in "16 10" "128 10" "1024 10" "1024 100"; do set -- $cfg
    timeout 100 python benchmarks/gemm_bench.py --nq $1 --k $2 --iters 20 > gpurun_out/sel_${lib}_$1_$2.json
    python -c "import json; d=json.load(open('gpurun_out/sel_${lib}_$1_$2.json')); print('$lib nq $1 k $2', round(d['ms_per_batch'],4))"
  done
done
