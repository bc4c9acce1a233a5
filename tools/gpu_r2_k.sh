#!/bin/bash
# round 2, run K: full validation + the measurements that go into profiles/
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2k_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2k_tests.txt
grep -E "bench-size|1000 atoms vs oracle|passed|failed|rc=" gpurun_out/r2k_tests.txt | tail -8
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2k_smoke.txt 2>&1; tail -2 gpurun_out/r2k_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_li3po4.json 2> gpurun_out/r2k_bench_li3po4.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2k_bench_reference_arm.json 2> gpurun_out/r2k_bench_reference_arm.err
for wl in water_1k_l2_f32 asi_50k_l3_f32; do
  timeout 900 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/r2k_bench_$wl.json 2> gpurun_out/r2k_bench_$wl.err
done
python - <<'PY'
import json
for wl in ('li3po4','water_1k_l2_f32','asi_50k_l3_f32'):
    try:
        d=json.loads(open(f'gpurun_out/r2k_bench_{wl}.json').read().strip().splitlines()[-1])
        print(wl, round(d['ms_per_step'],3),'ms', round(d['value']), 'e2e', round(d['e2e']['value']), 'nl-e2e', d.get('e2e_device_neighbor_list') and round(d['e2e_device_neighbor_list']['value']), d['clocks'])
        r=d['roofline']; print('  top', r['kernel'][:50], round(r['frac'],3), round(r['share_of_step'],3), 'cpu', d.get('cpu_baseline'))
    except Exception as e: print(wl,'ERR',e); print(open(f'gpurun_out/r2k_bench_{wl}.err').read()[-800:])
PY
cut -c1-400 gpurun_out/r2k_bench_reference_arm.json
# launch list of one eager step (shares) and full captures of the four hot kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2k_launches_li3po4.csv python bench.py --profile-step --no-graph > gpurun_out/r2k_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm3x -s 8 -c 2 -o gpurun_out/r2k_full_gemm python tools/bench_fused.py --layers 2 --reps 1 > gpurun_out/r2k_ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:tp_fwd2_kernel -s 1 -c 1 -o gpurun_out/r2k_full_tpfwd python tools/bench_fused.py --layers 2 --reps 1 > gpurun_out/r2k_ncu_tpfwd.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:tp_bwd2_kernel -s 1 -c 1 -o gpurun_out/r2k_full_tpbwd python tools/bench_kernels.py --skip-mlp --reps 1 > gpurun_out/r2k_ncu_tpbwd.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:tp_fused -s 1 -c 1 -o gpurun_out/r2k_full_fused python tools/bench_fused.py --layers 2 --reps 1 > gpurun_out/r2k_ncu_fused.log 2>&1
ls -la gpurun_out/*.ncu-rep
