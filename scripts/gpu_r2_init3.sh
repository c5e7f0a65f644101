#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_initiator.py -x -q -s 2>&1 | grep -E "config-3|passed|failed|Error|assert" | head
timeout 900 python scripts/bench_initiator.py > gpurun_out/initiator_r2.json 2> gpurun_out/initiator_r2.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/initiator_r2.json").read().strip().splitlines()[-1])
print({k: v for k, v in d.items() if k != "gpu_full"})
g = d["gpu_full"]
print({k: v for k, v in g.items() if k != "scans"})
for r in g["scans"][:5]:
    print(r["unused"], r["ms"], r["new_targets"], r["preliminary"], "| tracks:", {k: r["tracks"].get(k) for k in ("n_edges", "batches", "ms_gate", "ms_solve")}, "| initiators:", {k: r["initiators_gnn"].get(k) for k in ("n_edges", "largest_component", "batches", "ms_gate", "ms_solve")})
PY
