#!/bin/bash
# Round 2, the very last GPU call (~3 min): hidden-layer kernels v2 (built as the default) -- parity + A/B timing,
# the bench line with them, the model-level GPU tests on the final binary, then `ncu --set full` of the hot kernels.
mkdir -p gpurun_out
O=gpurun_out/r2y
timeout 80 python -m pytest tests/test_hidden_variants_gpu.py tests/test_radial_mlp_gpu.py -q -s -p no:cacheprovider > ${O}_hidden_tests.txt 2>&1; echo "rc=$?" >> ${O}_hidden_tests.txt
grep -E "hidden variant|passed|failed|rc=" ${O}_hidden_tests.txt | tail -5
timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > ${O}_bench_li3po4.json 2> ${O}_bench_li3po4.err
echo "bench rc=$?"; cut -c1-300 ${O}_bench_li3po4.json; tail -2 ${O}_bench_li3po4.err
timeout 150 python -m pytest tests/test_model_gpu.py tests/test_graph_gpu.py tests/test_modifiers_gpu.py tests/test_tp_fused_gpu.py -m gpu -q -p no:cacheprovider > ${O}_tests.txt 2>&1; echo "rc=$?" >> ${O}_tests.txt
tail -3 ${O}_tests.txt
timeout 100 ncu --set full --clock-control none --profile-from-start off -k regex:'k_gemm3x|tp_fwd2_kernel|tp_bwd2_kernel|k_hidden' \
    -f -o ${O}_full python bench.py --profile-step --no-graph > ${O}_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la ${O}_full.ncu-rep 2>/dev/null
