#!/bin/bash
# round 2 session I: MSA Transformer — pruned last layer, pointer-walking regroup kernels, true-size parity, timing
mkdir -p gpurun_out
nvidia-smi -L
echo "== 1. MSA Transformer tests"
timeout 1500 python -m pytest tests/test_gpu_z_msa_transformer.py -q -s 2>&1 | grep -v "^$" > gpurun_out/i1_msa_tests.log; grep "MSA\|passed\|failed\|Error" gpurun_out/i1_msa_tests.log | tail -30
echo "== 2. MSA-1b timing"
timeout 900 python scripts/bench_msa_transformer.py --rows 400 --length 512 --positions 8 2>&1 | tail -1 | tee gpurun_out/i2_msa_bench.jsonl
timeout 900 python scripts/bench_msa_transformer.py --rows 400 --length 512 --positions 8 --precision f16x3 2>&1 | tail -1 | tee -a gpurun_out/i2_msa_bench.jsonl
timeout 900 python scripts/bench_msa_transformer.py --rows 128 --length 255 --positions 16 --per-pass 8 2>&1 | tail -1 | tee -a gpurun_out/i2_msa_bench.jsonl
echo "== done"
