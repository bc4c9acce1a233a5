#!/bin/bash
# 4 GPUs: halo-decomposition parity tests (2 and 4 ranks, eager + graphed) and the bench at 2 / 4 ranks
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2mg4_gpus.txt
timeout 900 python -m pytest tests/test_parallel_nccl_gpu.py -q -s > gpurun_out/r2mg4_nccl_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2mg4_nccl_tests.txt
grep -E "halo world|passed|failed|rc=" gpurun_out/r2mg4_nccl_tests.txt | tail -8
for N in 2 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2mg4_bench_halo_weak_$N.json 2> gpurun_out/r2mg4_bench_halo_weak_$N.err
  echo "weak $N rc=$?"; cut -c1-260 gpurun_out/r2mg4_bench_halo_weak_$N.json; tail -2 gpurun_out/r2mg4_bench_halo_weak_$N.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 4 --steps 10 --warmup 3 --scaling strong > gpurun_out/r2mg4_bench_halo_strong_4.json 2> gpurun_out/r2mg4_bench_halo_strong_4.err
echo "strong 4 rc=$?"; cut -c1-260 gpurun_out/r2mg4_bench_halo_strong_4.json; tail -2 gpurun_out/r2mg4_bench_halo_strong_4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 --no-graph > gpurun_out/r2mg4_bench_halo_weak_4_eager.json 2> gpurun_out/r2mg4_bench_halo_weak_4_eager.err
echo "weak 4 eager rc=$?"; cut -c1-200 gpurun_out/r2mg4_bench_halo_weak_4_eager.json
