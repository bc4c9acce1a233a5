#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_fused.py --prof > gpurun_out/r2c_bench_fused.jsonl 2> gpurun_out/r2c_bench_fused.err; echo "rc=$?" >> gpurun_out/r2c_bench_fused.err
cat gpurun_out/r2c_bench_fused.jsonl; tail -5 gpurun_out/r2c_bench_fused.err
timeout 900 python -m pytest tests/test_tp_fused_gpu.py -x -q > gpurun_out/r2c_fused_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2c_fused_tests.txt
tail -4 gpurun_out/r2c_fused_tests.txt
