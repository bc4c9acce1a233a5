#!/bin/bash
# Round-end measurement on a GPU box: smoke, full GPU test suite, bench (both arms), ncu launch list.
# FULL=1 adds the `ncu --set full` captures of the dominant kernels (profiles/r01_ncu_full_*).
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_tests.txt
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2>> gpurun_out/final_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --no-graph --profile-step > /dev/null 2>&1
if [ "$FULL" = "1" ]; then
  for k in tp_fwd2_kernel tp_bwd2_kernel; do
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -c 2 -f -o gpurun_out/final_$k \
        python bench.py --no-graph --profile-step > /dev/null 2>&1
    ncu -i gpurun_out/final_$k.ncu-rep --page raw --csv > gpurun_out/final_$k.raw.csv 2>/dev/null
  done
  timeout 500 ncu --set full --clock-control none -k regex:k_gemm3x -c 14 -f -o gpurun_out/final_k_gemm3x \
      python bench.py --no-graph --profile-step > /dev/null 2>&1
  ncu -i gpurun_out/final_k_gemm3x.ncu-rep --page raw --csv > gpurun_out/final_k_gemm3x.raw.csv 2>/dev/null
fi
tail -2 gpurun_out/final_smoke.txt
cat gpurun_out/final_tests.txt
cut -c1-300 gpurun_out/final_bench.json
