#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_fused.py --workload asi_50k_l3_f32 --reps 3 > gpurun_out/r2h_bench_fused_asi.jsonl 2> gpurun_out/r2h_bench_fused_asi.err; cat gpurun_out/r2h_bench_fused_asi.jsonl; tail -2 gpurun_out/r2h_bench_fused_asi.err
timeout 600 python tools/bench_fused.py --workload water_1k_l2_f32 > gpurun_out/r2h_bench_fused_water.jsonl 2>&1; cat gpurun_out/r2h_bench_fused_water.jsonl
for wl in li3po4_10k_l2_f64 water_1k_l2_f32 asi_50k_l3_f32; do
  timeout 900 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/r2h_bench_$wl.json 2> gpurun_out/r2h_bench_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2h_bench_$wl.json').read().strip().splitlines()[-1])
    print('$wl', round(d['ms_per_step'],3),'ms', round(d['value']), 'e2e', round(d['e2e']['value']), d['config']['radial_tp_path'])
    r=d['roofline']; print('  top', r['kernel'], round(r['frac'],3), round(r['share_of_step'],3)); 
    for k,v in r['by_kernel'].items(): print('   ', k, round(v['ms_per_step_isolated'],3), round(v['share_of_step'],3), 'hbm', round(v['frac'],3), v.get('tensor',{}).get('frac'))
except Exception as e: print('$wl ERR', e); print(open('gpurun_out/r2h_bench_$wl.err').read()[-1500:])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2h_launches_asi.csv python bench.py --workload asi_50k_l3_f32 --profile-step --no-graph > gpurun_out/r2h_ncu_asi.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2h_launches_asi.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.Counter(); cnt=collections.Counter()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    k=r[ki][:60]; agg[k]+=v; cnt[k]+=1
tot=sum(agg.values())
for k,v in agg.most_common(12): print(round(v/1e6,3),'ms',cnt[k],round(100*v/tot,1),'%',k)
PY
