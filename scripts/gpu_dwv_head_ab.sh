for rep in 1 2; do
for cfg in "0 0" "1 0" "1 1"; do
  set -- $cfg
  VQA_PREFETCH_EARLY=$1 VQA_DWV_HEAD=$2 python bench.py --no-fp32 --no-infer --no-memft --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); c=d['critical_path_ms']
        print('early_pf $1 dwv_head $2', 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k:round(x,4) for k,x in c.items() if x>0})
"
done; done
