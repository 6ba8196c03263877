#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "config4 or every_shard" 2>&1 | tail -30 > gpurun_out/r02f_tests.log
