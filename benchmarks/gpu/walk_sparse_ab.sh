for st in 1.0 0.5; do TSS_WALK_STATIC=$st python benchmarks/exclude_probe.py --iters 30 > gpurun_out/excl2_$st.json 2>gpurun_out/excl2_$st.err; python -c "
import json; d=json.load(open('gpurun_out/excl2_$st.json'))
for c in d['cases']: print('static $st', c['case'], round(c['us'],1), 'us', round(c['gbs']), 'GB/s')"; done
