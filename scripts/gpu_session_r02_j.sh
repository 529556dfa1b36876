#!/bin/bash
# round 2 session J: final evidence — full GPU suite, default bench, MSA Transformer launch list + ncu capture of the grouped GEMMs
mkdir -p gpurun_out
nvidia-smi -L
echo "== 1. full GPU suite"
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | grep -v "^$" > gpurun_out/j1_suite.log; grep "passed\|failed\|MSA\|Error" gpurun_out/j1_suite.log | tail -25
echo "== 2. bench (defaults)"
timeout 1500 python bench.py > gpurun_out/bench_r02_final_n1.json 2> gpurun_out/bench_r02_final_n1.err; tail -c 600 gpurun_out/bench_r02_final_n1.json
echo "== 3. MSA launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r02_msa.csv \
  python scripts/bench_msa_transformer.py --rows 400 --length 512 --positions 4 --per-pass 4 > gpurun_out/j3.log 2>&1; tail -2 gpurun_out/j3.log | cut -c1-300
echo "== 4. ncu --set full of the tied-attention grouped GEMMs (layer 0: launches 2 and 3 of gemm_tc_kernel)"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:gemm_tc_kernel --launch-skip 1 --launch-count 2 \
  -o gpurun_out/prof_r02_msa_tied_gemms python scripts/bench_msa_transformer.py --rows 400 --length 512 --positions 4 --per-pass 4 > gpurun_out/j4.log 2>&1; tail -2 gpurun_out/j4.log | cut -c1-200
echo "== 5. msa prior kernel"; timeout 300 python scripts/bench_msa.py 2>&1 | tail -3 | tee gpurun_out/bench_msa_r02.jsonl
echo "== done"
