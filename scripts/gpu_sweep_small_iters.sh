for it in 300 1200; do
echo "=== MHT_DUAL_ITERS=$it"
MHT_DUAL_ITERS=$it python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "cfg3_vs_reference" 2>&1 | grep -E "^cfg3|identical|passed|failed" | sed -E "s/.n_parents.*.certified./certified/; s/.ms_gate.[^,]*, //" | cut -c1-300
done
