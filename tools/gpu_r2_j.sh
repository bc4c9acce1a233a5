#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tune_tp.py --cfg 2,32,4 --layers 1,2 --variants irmul,irmul_noring > gpurun_out/r2j_tune_water.jsonl 2> gpurun_out/r2j_tune.err; cat gpurun_out/r2j_tune_water.jsonl
timeout 600 python tools/tune_tp.py --cfg 3,32,5 --layers 1 --variants irmul,irmul_noring > gpurun_out/r2j_tune_asi1.jsonl 2>> gpurun_out/r2j_tune.err; cat gpurun_out/r2j_tune_asi1.jsonl
timeout 600 python tools/tune_tp.py --cfg 2,64,4 --layers 1,2 --variants irmul,irmul_noring > gpurun_out/r2j_tune_li.jsonl 2>> gpurun_out/r2j_tune.err; cat gpurun_out/r2j_tune_li.jsonl
tail -3 gpurun_out/r2j_tune.err
