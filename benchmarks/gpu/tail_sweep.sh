# tile-schedule knobs at the per-GPU shard size of the 8-GPU headline (1.25M rows) and at 10M
for rows in 1250000 10000000; do for cfg in "8 2" "8 4" "4 2" "4 4" "2 4" "4 8" "16 2"; do set -- $cfg
  TSS_DYN_CHUNK=$1 TSS_FINE_ROUNDS=$2 python bench.py --rows $rows --steps 300 --warmup 20 --no-extras --no-batched --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('rows $rows chunk $1 fine $2: ms', round(d['ms_per_step'],4), 'GB/s', round(d['roofline']['achieved'],1), d['check'])"
done; done
