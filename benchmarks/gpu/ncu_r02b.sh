# final ncu pass of round 2: K2 after the measured margins, and the launch list of bench.py
set -x
G="python benchmarks/gemm_bench.py --iters 1"
$G > gpurun_out/r02n_gemm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:gemm_topk|select_kernel' -s 6 -c 3 -f -o gpurun_out/r02n_gemm $G > gpurun_out/r02n_gemm_ncu.log 2>&1
$G > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 30 --csv --log-file gpurun_out/r02n_gemm_launches.csv $G > gpurun_out/r02n_gemm_launches_ncu.log 2>&1
B="python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline"
$B > gpurun_out/r02n_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02n_launches.csv $B > gpurun_out/r02n_launches_ncu.log 2>&1
ls -la gpurun_out/r02n*
