#!/bin/bash
mkdir -p gpurun_out
( python -c "import __graft_entry__ as g; g.smoke()"; RS_BOARD_MAJOR=1 timeout 80 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "flop_rooted or turn_river_small or river_small" 2>&1 | tail -3 ) > gpurun_out/r02i2.log 2>&1
