for it in 20 40 80 120 240; do
MHT_DUAL_ITERS=$it python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['scan_stats']; print('iters $it', 'value %.1f e2e %.1f'%(d['value'], d['e2e']['value']), {k:round(v,2) for k,v in d['stage_ms'].items()}, 'LB %.2f OBJ %.2f gap %.2f dual_iters %.0f'%(s['lower_bound'], s['objective'], s['objective']-s['lower_bound'], s['dual_iters']))"
done
