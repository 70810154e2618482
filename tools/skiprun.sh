for fam in none norm_stats norm_finalize norm_apply norm_bwd_reduce norm_bwd_finalize norm_bwd_apply act_bwd colsum halo_fill "gather_cast,cast_bf16" wgrad tapgemm "image_to_kwexp,kwexp_to_image_grad,rspace_combine,rspace_expand" "linear_fwd,linear_bwd" adam; do
  MUNIT_SKIP=$fam python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('SKIP %-70s ms/step %.2f' % ('$fam', d['ms_per_step']))
except Exception as e: print('SKIP $fam failed', e)
"
done
