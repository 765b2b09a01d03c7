# select_kernel after survivors are staged in shared memory: full capture at 16 and 1 024 queries
# (top-10 / top-100), and the launch list of a 16-query batch
set -x
for cfg in "16 10" "1024 100"; do set -- $cfg
  G="python benchmarks/gemm_bench.py --iters 1 --nq $1 --k $2"
  $G > gpurun_out/r02u_plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k 'regex:select_kernel' -s 2 -c 1 -f -o gpurun_out/r02u_select_nq$1 $G > gpurun_out/r02u_select_nq$1_ncu.log 2>&1
done
G="python benchmarks/gemm_bench.py --iters 1 --nq 16 --k 10"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02u_gemm_nq16_launches.csv $G > /dev/null 2>&1
ls -la gpurun_out/r02u*
