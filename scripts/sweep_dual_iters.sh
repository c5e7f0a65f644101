for it in 10 30 60 120 300; do
MHT_DUAL_ITERS=$it python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['scan_stats']; print($it, 'ms/step %.1f'%d['ms_per_step'], 'assoc %.1f'%d['stage_ms']['ms_assoc'], 'iters',s['dual_iters'],'LB %.3f UB %.3f gap %.3f'%(s['lower_bound'],s['objective'],s['objective']-s['lower_bound']), 'children %.3g'%s['n_children'])"
done
