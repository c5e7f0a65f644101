python -m pytest tests -q -m gpu -s 2>&1 | grep -E "^cfg3.*scan [23]|passed|failed|FAILED|^E " | sed -E "s/.n_parents.*.certified./certified/; s/.ms_gate.[^,]*, //; s/'dual_iters.*'lower_bound'/LB/" | cut -c1-200
python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['scan_stats']; print('value %.1f'%d['value'], 'assoc %.2f'%d['stage_ms']['ms_assoc'], 'LB %.2f OBJ %.2f gap %.2f'%(s['lower_bound'], s['objective'], s['objective']-s['lower_bound']))"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ls_" -c 70 --csv --log-file gpurun_out/ls_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/ls_launches.csv | head -12
