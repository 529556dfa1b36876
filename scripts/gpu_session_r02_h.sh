#!/bin/bash
# round 2 session H: first light of the MSA Transformer path
mkdir -p gpurun_out
nvidia-smi -L
echo "== 1. MSA Transformer tests"
timeout 1500 python -m pytest tests/test_gpu_z_msa_transformer.py -q -s 2>&1 | tail -60 > gpurun_out/h1_msa_tests.log; tail -40 gpurun_out/h1_msa_tests.log
echo "== 2. kernel-level regression (gemm / layernorm / attention)"
true
echo "== 3. MSA-1b timing"
true
true
true
echo "== done"
