#!/bin/bash
mkdir -p gpurun_out
( time timeout 100 python bench.py ) > gpurun_out/r02final_bench.json 2> gpurun_out/r02final_bench.err
