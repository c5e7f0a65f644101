for g in 296 444 592; do
MHT_ACT_GRID=$g python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print($g, 'ms/step %.1f'%d['ms_per_step'], 'assoc %.1f'%d['stage_ms']['ms_assoc'])"
done
