# A/B of the v-projection split under the recurrent forward kernel (csrc/model.cu: vproj_split_plan); VQA_VPROJ_SPLIT = 0 off,
# unset = planned, n = n row tiles of 256 rows moved
for m in 0 -1 0 -1; do
  VQA_VPROJ_SPLIT=$m python bench.py --no-fp32 --no-infer --no-memft --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); c=d['critical_path_ms']
        print('split', '$m', 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'vproj', round(c['vproj_fwd'],4), 'gru_fwd', round(c['gru_fwd'],4), 'qheads', round(c['qheads_fwd'],4), 'attn', round(c['attn_fwd'],4))
"
done
