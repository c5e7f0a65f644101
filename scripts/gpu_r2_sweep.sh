mkdir -p gpurun_out
run() { # label, env...
  lab=$1; shift
  env "$@" MHT_BENCH_SKIP_E2E=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lab', 'value %.1f' % d['value'], 'p50 %.1f p95 %.1f max %.1f' % (d['scan_ms']['ms_total']['p50'], d['scan_ms']['ms_total']['p95'], d['scan_ms']['ms_total']['max']), 'certified', d['ilp']['certified_scans'], 'exact_ms %.2f' % d['stage_ms']['ms_exact'], 'dual %.2f' % d['stage_ms']['ms_dual'], 'gap_mean %.2f' % d['ilp']['gap_mean'])"
}
run "default          " X=1
run "knode30          " MHT_BB_KNODE=30
run "knode30 kroot100 " MHT_BB_KNODE=30 MHT_BB_KROOT=100
run "ms20             " MHT_BB_MS=20
run "ms20 knode30     " MHT_BB_MS=20 MHT_BB_KNODE=30
run "ms5              " MHT_BB_MS=5
run "greedy60         " MHT_GREEDY_EVERY=60
run "iters90          " MHT_DUAL_ITERS=90
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('full e2e check: value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))"
