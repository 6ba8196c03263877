#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "street_kernel_matches or batch" 2>&1 | grep -E "AssertionError|passed|failed|FAILED|Error" > gpurun_out/r02i_tests.log
run() { # name, workload, env...
  name=$1; w=$2; shift 2
  env RS_ENGINE_FLAGS=4 "$@" timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02i_${name}.json 2> gpurun_out/r02i_${name}.err
}
run c5_default config5
run c2_default config2
run c5_t256 config5 RS_STREET_THREADS=256
run c5_t192 config5 RS_STREET_THREADS=192
run c2_t192 config2 RS_STREET_THREADS=192
run c2_t256 config2 RS_STREET_THREADS=256
run c5_smem110 config5 RS_STREET_SMEM_KB=110
run c2_smem110 config2 RS_STREET_SMEM_KB=110
