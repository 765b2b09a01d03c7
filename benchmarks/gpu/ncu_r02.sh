# ncu captures of round 2 (one gpurun call; every profiled command first exits 0 without ncu)
set -x
P="python benchmarks/masked_probe.py --iters 1 --sel 0.001 0.01 0.11"
$P > gpurun_out/r02_masked_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_topk -c 12 -f -o gpurun_out/r02_masked_scan $P > gpurun_out/r02_masked_ncu.log 2>&1
$P > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:prefix_mask -c 18 -f -o gpurun_out/r02_k4 $P > gpurun_out/r02_k4_ncu.log 2>&1
G="python benchmarks/gemm_bench.py --iters 1"
$G > gpurun_out/r02_gemm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:gemm_topk|select_kernel' -s 6 -c 3 -f -o gpurun_out/r02_gemm $G > gpurun_out/r02_gemm_ncu.log 2>&1
B="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
$B > gpurun_out/r02_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_launches_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
