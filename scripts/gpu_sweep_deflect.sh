for a in 1.0 0.7 0.5 0.3; do
echo "=== MHT_DEFLECT=$a"
MHT_DEFLECT=$a python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "cfg3_vs_reference" 2>&1 | grep -E "^cfg3|identical|passed|failed" | sed -E "s/'n_parents.*'certified'/'certified'/; s/'ms_gate.*//" | cut -c1-260
MHT_DEFLECT=$a python -m pytest tests/test_gpu_parity.py -q -m gpu -k "replays or assoc_vs_oracle" 2>&1 | tail -1
MHT_DEFLECT=$a python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['scan_stats']; print('bench value %.1f'%d['value'], 'assoc %.2f ms'%d['stage_ms']['ms_assoc'], 'LB %.2f OBJ %.2f gap %.2f'%(s['lower_bound'], s['objective'], s['objective']-s['lower_bound']))"
done
