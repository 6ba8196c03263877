#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_abstraction.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r02y_tests.log
python - > gpurun_out/r02y_hist.log 2>&1 <<'PY'
import rustsolver_b200 as rb
h, cards, ms = rb.generate_histograms(1, 0, 20000, 100, 50, seed=1, return_ms=True)
print(f"generate_histograms flop: 20000 hands x 100 samples x 990 opponent combos in {ms:.1f} ms = {20000*100*990/ms/1e6:.1f} G evaluations/s")
PY
