set -x
P="python benchmarks/exclude_probe.py --iters 2"
$P > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:scan_topk' -s 8 -c 12 -f -o gpurun_out/r02z_dense_masked $P > gpurun_out/r02z_dense_ncu.log 2>&1
