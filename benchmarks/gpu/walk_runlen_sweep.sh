for w in 3 4 5; do export TSS_WALK_RUN=$w
  python benchmarks/exclude_probe.py --iters 30 > gpurun_out/rl_e$w.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/rl_e$w.json'))
for c in d['cases'][1:3]+d['cases'][5:]: print('run 2^$w', c['case'], round(c['us'],1))"
  python benchmarks/masked_probe.py --sel 0.02 0.11 0.6 --contig 0.3 > gpurun_out/rl_m$w.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/rl_m$w.json'))
for c in d['cases']: print('run 2^$w', c['selectivity'], round(c['masked_scan_us'],1), 'us', round(c['live_gbs']), 'GB/s')"
done
