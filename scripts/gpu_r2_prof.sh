mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "dynamic or golden or dropin" 2>&1 | grep -E "^E |passed|failed|Error" | head -20
timeout 1700 bash profiles/prof_r2.sh
