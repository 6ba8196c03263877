#!/bin/bash
mkdir -p gpurun_out
RS_ENGINE_FLAGS=4 timeout 600 python bench.py --workload config4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_c4_street.json 2> gpurun_out/r02j_c4_street.err
RS_ENGINE_FLAGS=4 ncu --set full --clock-control none --import-source on -k regex:street_kernel -s 6 -c 1 -o gpurun_out/r02j_prof_c5_street python bench.py --workload config5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_prof_c5.log 2>&1
RS_ENGINE_FLAGS=4 ncu --set full --clock-control none --import-source on -k regex:street_kernel -s 6 -c 1 -o gpurun_out/r02j_prof_c2_street python bench.py --workload config2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_prof_c2.log 2>&1
