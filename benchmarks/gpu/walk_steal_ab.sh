# mask walk: moving on to other claim counters when the home counter runs dry (TSS_WALK_RUN=5, the
# default) vs staying on the home counter (TSS_WALK_RUN=21), same box, twice
for rep in 1 2; do for w in 21 5; do
  TSS_WALK_RUN=$w python benchmarks/exclude_probe.py --iters 30 > gpurun_out/steal_$w.json 2> gpurun_out/steal_$w.err
  python -c "
import json; d=json.load(open('gpurun_out/steal_$w.json'))
for c in d['cases'][1:3]+d['cases'][5:]: print('rep $rep TSS_WALK_RUN=$w', c['case'], round(c['us'],1), 'us')"
  TSS_WALK_RUN=$w python benchmarks/masked_probe.py --sel 0.02 0.11 0.6 --contig 0.3 > gpurun_out/steal_m$w.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/steal_m$w.json'))
for c in d['cases']: print('rep $rep TSS_WALK_RUN=$w', c['selectivity'], round(c['masked_scan_us'],1), 'us', round(c['live_gbs']), 'GB/s')"
done; done
