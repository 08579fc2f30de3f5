for m in 0 3 4 8 12 15 79 127; do
  VQA_LINEAR_LN_SITES=$m python bench.py --no-fp32 --no-infer --no-memft --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); c=d['critical_path_ms']
        print('sites', $m, 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'qheads', round(c['qheads_fwd'],4), 'head_fwd', round(c['head_fwd'],4), 'head_bwd', round(c['head_bwd'],4), 'launches', d['gpu_launches'])
"
done
