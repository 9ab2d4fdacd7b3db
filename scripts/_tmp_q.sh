(timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x --tb=short -k "conv or doc_plan" 2>&1 | tail -4)
(timeout 120 python scripts/conv_repro.py --reps 20 2>&1 | tail -3)
for d in amazon nopad; do
  (CONV_PROF=1 timeout 100 python scripts/conv_bench.py --dist $d 2>&1 | tail -8)
done
