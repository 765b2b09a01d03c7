(timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_prefix_gpu.py -m gpu -x -q 2>&1 | tail -12) > gpurun_out/r02i_pytest.log; cat gpurun_out/r02i_pytest.log
for cfg in "bf16 0 0" "bf16 1 0" "bf16 1 1" "f32 1 0" "f32 1 1"; do set -- $cfg; for it in 5 300; do
TSS_GEMM_UNIT_SHADOW=$2 TSS_GEMM_WEIGHTS=$3 timeout 100 python benchmarks/gemm_bench.py --storage $1 --iters $it > gpurun_out/r02i_gemm_$1_u$2_w$3_i$it.json
python -c "import sys,json; d=json.load(open('gpurun_out/r02i_gemm_$1_u$2_w$3_i$it.json')); print('$1 unit=$2 weights=$3 iters $it', round(d['ms_per_batch'],3), round(d['frac_of_burst'],3), round(d['frac_of_sustained'],3))"
done; done
