# the masked scan after the run-based walk with claimed runs: full capture at 11 % and 60 % random masks
set -x
P="python benchmarks/masked_probe.py --iters 1 --sel 0.11 0.6"
$P > gpurun_out/r02z_masked_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:scan_topk' -c 10 -f -o gpurun_out/r02z_masked_scan $P > gpurun_out/r02z_masked_ncu.log 2>&1
ls -la gpurun_out/r02z*
