#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/bench_initiator.py > gpurun_out/initiator_r2.json 2> gpurun_out/initiator_r2.err
echo "exit $?"; tail -3 gpurun_out/initiator_r2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/initiator_r2.json").read().strip().splitlines()[-1])
print({k: v for k, v in d.items() if k != "gpu_full"})
g = d["gpu_full"]
print({k: v for k, v in g.items() if k != "scans"})
for r in g["scans"]:
    print(r["unused"], r["ms"], r["new_targets"], r["preliminary"], "| tracks gnn:", {k: r["tracks"].get(k) for k in ("n_edges", "largest_component", "batches", "searches", "ms_gate", "ms_solve")}, "| initiators gnn:", {k: r["initiators_gnn"].get(k) for k in ("n_edges", "largest_component", "batches", "searches", "ms_gate", "ms_solve")})
PY
for v in "MHT_GNN_SIZE_FIRST=1" "MHT_GNN_ROWCAP=128" "MHT_GNN_ROWCAP=64 MHT_GNN_SIZE_FIRST=1"; do
  echo "== $v"; env $v timeout 600 python -m pytest tests/test_gpu_initiator.py -x -q -s -k "giant" 2>&1 | grep -E "config-3|passed|failed"
done
