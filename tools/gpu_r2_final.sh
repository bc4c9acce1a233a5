#!/bin/bash
# Round 2, last GPU call (a few minutes of budget left): the bench line of the committed code, the full
# GPU test suite, and the ncu launch list of one (graph-replayed) step -- in that order of priority.
mkdir -p gpurun_out
O=gpurun_out/r2z
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > ${O}_gpu.txt 2>&1
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > ${O}_bench_li3po4.json 2> ${O}_bench_li3po4.err
echo "bench rc=$?"; cut -c1-330 ${O}_bench_li3po4.json; tail -3 ${O}_bench_li3po4.err
# the tests added by the last commit first (test_model_gpu.py), then everything else
( timeout 600 python -m pytest tests/test_model_gpu.py tests -m gpu -q -p no:cacheprovider > ${O}_tests.txt 2>&1; echo "rc=$?" >> ${O}_tests.txt ) &
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file ${O}_launches_li3po4.csv python bench.py --profile-step > ${O}_ncu_list.log 2>&1
echo "ncu list rc=$?"; wc -l ${O}_launches_li3po4.csv
wait
tail -5 ${O}_tests.txt
for wl in asi_50k_l3_f32 water_1k_l2_f32; do
  timeout 200 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > ${O}_bench_$wl.json 2> ${O}_bench_$wl.err
  echo "$wl rc=$?"; cut -c1-200 ${O}_bench_$wl.json
done
timeout 120 python __graft_entry__.py smoke > ${O}_smoke.txt 2>&1; tail -1 ${O}_smoke.txt
