#!/bin/bash
# round 2, run A: validate the phase-0 changes and re-baseline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2a_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r2a_smoke.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload water_1k_l2_f32 --no-cpu-baseline > gpurun_out/r2a_bench_water.json 2> gpurun_out/r2a_bench_water.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload asi_50k_l3_f32 --no-cpu-baseline > gpurun_out/r2a_bench_asi.json 2> gpurun_out/r2a_bench_asi.err
tail -3 gpurun_out/r2a_tests.txt; cat gpurun_out/r2a_smoke.txt | tail -2; cat gpurun_out/r2a_bench.json | cut -c1-400
