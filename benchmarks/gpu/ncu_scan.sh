# refresh profiles/scan_traffic.json: DRAM bytes per launch of the headline scan (10M x 384 fp32)
set -x
B="python bench.py --steps 6 --warmup 3 --no-extras --no-batched --no-cpu-baseline"
$B > gpurun_out/r02o_bench_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 4 -c 2 -f -o gpurun_out/r02o_scan $B > gpurun_out/r02o_scan_ncu.log 2>&1
ls -la gpurun_out/r02o*
