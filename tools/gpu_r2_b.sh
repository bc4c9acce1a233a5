#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tp_fused_gpu.py -x -q -k "forward_matches" > gpurun_out/r2b_fused_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2b_fused_tests.txt
tail -15 gpurun_out/r2b_fused_tests.txt
timeout 600 python tools/bench_fused.py > gpurun_out/r2b_bench_fused.jsonl 2> gpurun_out/r2b_bench_fused.err; echo "rc=$?" >> gpurun_out/r2b_bench_fused.err
cat gpurun_out/r2b_bench_fused.jsonl; tail -5 gpurun_out/r2b_bench_fused.err
timeout 900 python -m pytest tests/test_tp_fused_gpu.py -x -q -k "model_with" > gpurun_out/r2b_fused_model_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2b_fused_model_tests.txt
tail -8 gpurun_out/r2b_fused_model_tests.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
cut -c1-300 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
