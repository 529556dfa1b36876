#!/bin/bash
# round 2 session N: un-fenced remote accumulator release in the CTA-pair GEMM (A/B through PG_GEMM_FENCED_RELEASE), EVE log prior on
# the library's kernels
mkdir -p gpurun_out
nvidia-smi -L
echo "== 1. GEMM parity (pair mode and single), model parity"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm_matches_fp64 or gemm_f16f8 or golden_small or model_matches" 2>&1 | tail -3 | tee gpurun_out/n1_gemm_parity.log
echo "== 2. GEMM microbench: fenced (1) vs plain (0) remote release"
for v in 1 0 1 0; do echo "-- PG_GEMM_FENCED_RELEASE=$v"; PG_GEMM_FENCED_RELEASE=$v timeout 300 python scripts/bench_gemm.py 1280 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    if d.get('cta2')==1 or 'cta2' not in d: print(d['gemm'], 'nseg', d['nseg'], 'cta2', d.get('cta2'), d['ms'], 'ms', d['issued_tflops'], 'issued TF/s', d['issued_frac_of_burst_peak'])
"; done | tee gpurun_out/n2_bench_gemm_ab.txt
echo "== 3. bench A/B (3 steps each, same box)"
for cfg in "PG_GEMM_FENCED_RELEASE=1" "PG_GEMM_FENCED_RELEASE=0" "PG_GEMM_FENCED_RELEASE=1" "PG_GEMM_FENCED_RELEASE=0"; do
  echo "-- $cfg"; env $cfg timeout 300 python bench.py --steps 3 --warmup 3 --no-other-workloads --no-cpu-baseline --no-other-modes 2> gpurun_out/n3_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['roofline']['kernel_ms_in_timed_region']
print(round(d['value']), 'mut/s', round(d['ms_per_step'],1),'ms/step', 'clk', d['clocks']['sm_mhz'], 'frac', round(d['roofline']['frac'],3), 'issued', round(d['roofline']['issued_frac'],3), {k: round(v['ms']/d['steps'],1) for k,v in c.items() if v['ms']>1})
" || tail -3 gpurun_out/n3_err.log; done 2>&1 | tee gpurun_out/n3_ab.txt
echo "== 4. EVE log prior on pg_gemm / pg_eve_output_conv"
timeout 600 python -m pytest tests/test_gpu_trancepteve.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/n4_trancepteve.log
timeout 300 python scripts/bench_eve_prior.py 500 200000 2>&1 | tail -2 | tee gpurun_out/bench_eve_prior_r02.jsonl
echo "== done"
