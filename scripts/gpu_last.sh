MHT_SIFT_ROUNDS=2 python bench.py > gpurun_out/bench_sift2.json 2> gpurun_out/bench_sift2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_sift2.json')); s=d['scan_stats']
print('sift2', 'value %.1f e2e %.1f'%(d['value'], d['e2e']['value']), d['stage_ms'], 'LB %.2f OBJ %.2f gap %.2f'%(s['lower_bound'], s['objective'], s['objective']-s['lower_bound']), 'frac %.3f'%d['roofline']['frac'])
PY
