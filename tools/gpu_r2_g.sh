#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tp_fused_gpu.py -x -q > gpurun_out/r2g_fused_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2g_fused_tests.txt
tail -4 gpurun_out/r2g_fused_tests.txt
timeout 600 python tools/bench_fused.py --prof > gpurun_out/r2g_bench_fused.jsonl 2> gpurun_out/r2g_bench_fused.err; echo "rc=$?" >> gpurun_out/r2g_bench_fused.err
python - <<'PY'
import json
for l in open('gpurun_out/r2g_bench_fused.jsonl'):
    d=json.loads(l); p=d.pop('prof_first_cta_of_slice',None); c=d.pop('cta_total_Mcycles',None)
    print(d)
    if p:
        for r in p: print('   ', r)
    if c: print('    cta Mcycles min/max', min(c), max(c))
PY
tail -3 gpurun_out/r2g_bench_fused.err
timeout 600 python -m pytest tests/test_modifiers_gpu.py tests/test_torch_library.py -q -m gpu > gpurun_out/r2g_misc_tests.txt 2>&1; tail -3 gpurun_out/r2g_misc_tests.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
cut -c1-300 gpurun_out/r2g_bench.json; tail -3 gpurun_out/r2g_bench.err
