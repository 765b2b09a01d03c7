# share of the mask walk's runs assigned statically (the rest claimed from the tile counter)
for st in ${STATICS:-1.0 0.5 0.0}; do
  export TSS_WALK_STATIC=$st
  python benchmarks/exclude_probe.py > gpurun_out/excl_$st.json 2> gpurun_out/excl_$st.err
  python -c "
import json; d=json.load(open('gpurun_out/excl_$st.json'))
for c in d['cases'][:3]: print('static $st', c['case'], round(c['us'],1), 'us', round(c['gbs']), 'GB/s')"
  python benchmarks/masked_probe.py --sel 0.002 0.01 0.02 0.05 0.11 0.3 0.6 --contig 0.11 0.3 > gpurun_out/walk_s$st.json 2> gpurun_out/walk_s$st.err
  python -c "
import json; d=json.load(open('gpurun_out/walk_s$st.json'))
for c in d['cases']: print('static $st', c['selectivity'], c['live_rows'], c['mask_ok'], 'list' if c['list_driven'] else 'walk', round(c['masked_scan_us'],1), 'us', round(c['live_gbs']), 'GB/s')"
done
