#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/tune_tp.py --cfg 3,32,5 --layers 2,3 --variants irmul,irmul_noring,irmul_noring_split,irmul_noring_split_a16,irmul_noring_split_a64,irmul_ring_a16 > gpurun_out/r2i_tune_asi.jsonl 2> gpurun_out/r2i_tune_asi.err
cat gpurun_out/r2i_tune_asi.jsonl; tail -3 gpurun_out/r2i_tune_asi.err
