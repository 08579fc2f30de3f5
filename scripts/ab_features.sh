run() { env "$@" python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-fp32 --no-infer 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
c=d['critical_path_ms']
print('  step %.4f e2e %.4f | attn_fwd %.4f attn_bwd %.4f gru_fwd %.4f vproj %.4f head_fwd %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step'], c['attn_fwd'], c['attn_bwd'], c['gru_fwd'], c['vproj_fwd'], c['head_fwd']))"; }
echo default; run A=1
echo KEEP_BITS=0; run VQA_ATTN_KEEP_BITS=0
echo VRING=0; run VQA_ATTN_BWD_VRING=0
echo PREFETCH=0; run VQA_PREFETCH_FEATURES=0
echo default; run A=1
