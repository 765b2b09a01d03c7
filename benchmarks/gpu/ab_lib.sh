# same-box A/B of two builds: benchmarks/gpu/ab/libtss_prev.so (git HEAD when it was built) vs the tree's
for rep in 1 2; do for lib in prev new; do for it in 5 300; do
  if [ $lib = prev ]; then export TSS_LIB_PATH=$PWD/benchmarks/gpu/ab/libtss_prev.so; else unset TSS_LIB_PATH; fi
  timeout 100 python benchmarks/gemm_bench.py --iters $it ${AB_ARGS:-} > gpurun_out/ab_${lib}_$it.json
  python -c "import json; d=json.load(open('gpurun_out/ab_${lib}_$it.json')); print('$lib rep $rep iters $it', round(d['ms_per_batch'],3), round(d['frac_of_burst'],3), round(d['frac_of_sustained'],3))"
done; done; done
