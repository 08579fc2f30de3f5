# generic A/B of one environment switch inside the cfg1 train step: bash scripts/gpu_env_ab.sh VAR v1 v2 ...  (each value twice, interleaved)
var=$1; shift
for rep in 1 2; do for v in "$@"; do
  env $var=$v python bench.py --no-fp32 --no-infer --no-memft --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); c=d['critical_path_ms']
        print('$var', '$v', 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k:round(x,4) for k,x in c.items() if x>0})
"
done; done
