# mask-walk order sweep: TSS_WALK_RUN = log2 of the consecutive tiles per run (0 = the old one-tile
# interleave; 5 = the default: runs of 32 tiles, consecutive runs to the 16 warps of one CTA); +8 =
# consecutive runs to different CTAs.  (profiles/r02x_walk_sweep2.txt was taken when +8 meant the
# opposite: there "13" is today's default and "5" is today's 13.)
for w in ${WALKS:-0 3 5}; do
  TSS_WALK_RUN=$w python benchmarks/masked_probe.py --sel 0.02 0.05 0.11 0.3 0.6 --contig 0.001 0.01 0.11 0.3 > gpurun_out/walk_$w.json 2> gpurun_out/walk_$w.err
  python -c "
import json; d=json.load(open('gpurun_out/walk_$w.json'))
for c in d['cases']: print('TSS_WALK_RUN=$w', c['selectivity'], c['live_rows'], c['mask_ok'], round(c['masked_scan_us'],1), 'us', round(c['live_gbs']), 'GB/s')"
done
