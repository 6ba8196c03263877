#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "street or config2_full or config4" 2>&1 | grep -E "AssertionError|passed|failed|FAILED" > gpurun_out/r02g_tests.log
