#!/bin/bash
for ms in 1 3 16 40; do
  echo "== MHT_EXACT_MS=$ms"; MHT_EXACT_MS=$ms timeout 300 python scripts/debug_lb.py 8 2>&1 | grep -v "NOT optimal" | grep -E "repaired [1-9]|LB > OBJ|conflicts [1-9]|bad scans|rep 0 scan 3"
done
