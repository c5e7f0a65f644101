run() { env "$@" python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['scan_stats']; print('$*', 'value %.1f'%d['value'], 'assoc %.2f ms'%d['stage_ms']['ms_assoc'], 'LB %.2f OBJ %.2f gap %.2f iters %.0f'%(s['lower_bound'], s['objective'], s['objective']-s['lower_bound'], s['dual_iters']))"; }
run MHT_SIFT_ROUNDS=3 MHT_DUAL_ITERS=120 MHT_GREEDY_EVERY=40
run MHT_SIFT_ROUNDS=3 MHT_DUAL_ITERS=120 MHT_GREEDY_EVERY=120
run MHT_SIFT_ROUNDS=3 MHT_DUAL_ITERS=120 MHT_GREEDY_EVERY=20
run MHT_SIFT_ROUNDS=2 MHT_DUAL_ITERS=180 MHT_GREEDY_EVERY=60
run MHT_SIFT_ROUNDS=6 MHT_DUAL_ITERS=60 MHT_GREEDY_EVERY=30
run MHT_SIFT_ROUNDS=4 MHT_DUAL_ITERS=60 MHT_GREEDY_EVERY=30
run MHT_SIFT_ROUNDS=2 MHT_DUAL_ITERS=120 MHT_GREEDY_EVERY=40
