#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tp_fused_gpu.py -x -q > gpurun_out/r2f_fused_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2f_fused_tests.txt
tail -4 gpurun_out/r2f_fused_tests.txt
timeout 600 python tools/bench_fused.py --prof > gpurun_out/r2f_bench_fused.jsonl 2> gpurun_out/r2f_bench_fused.err; echo "rc=$?" >> gpurun_out/r2f_bench_fused.err
python - <<'PY'
import json
for l in open('gpurun_out/r2f_bench_fused.jsonl'):
    d=json.loads(l); p=d.pop('prof_first_cta_of_slice',None); c=d.pop('cta_total_Mcycles',None)
    print(d)
    if p:
        for r in p: print('   ', r)
    if c: print('    cta Mcycles min/max', min(c), max(c))
PY
tail -3 gpurun_out/r2f_bench_fused.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tp_fused -c 1 -o gpurun_out/r2f_prof_fused python tools/bench_fused.py --layers 2 --reps 1 > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_ncu.log
