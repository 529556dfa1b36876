#!/bin/bash
# round 2 session L: packed-fp32 arithmetic (FFMA2/FADD2/FMUL2) in the attention softmax (attn_softmax=2) and the GEMM epilogue
# (gemm_epi=2): kernel parity, isolated throughput old vs new, same-box bench A/B, model parity, ncu of the new attention kernel.
mkdir -p gpurun_out
nvidia-smi -L
echo "== 1. kernel parity (new defaults)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "attention_matches or gemm_matches_fp64 or gemm_f16f8 or layernorm" 2>&1 | tail -3 | tee gpurun_out/l1_kernel_parity.log
echo "== 2. attention: 19 parity cases + isolated throughput, scalar (1) vs packed (2) softmax"
for v in 1 2; do echo "-- PG_ATTN_SOFTMAX=$v"; PG_ATTN_SOFTMAX=$v timeout 300 python scripts/check_attention_impl.py 0 2>&1 | grep -v "  ok" ; done | tee gpurun_out/l2_attention_ab.txt
echo "== 3. GEMM microbench: scalar (1) vs packed (2) epilogue"
for v in 1 2; do echo "-- PG_GEMM_EPI=$v"; PG_GEMM_EPI=$v timeout 300 python scripts/bench_gemm.py 1280 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['gemm'], 'nseg', d['nseg'], d['ms'], 'ms', d['issued_tflops'], 'issued TF/s', d['issued_frac_of_burst_peak'])
"; done | tee gpurun_out/l3_bench_gemm_ab.txt
echo "== 4. bench A/B (3 steps each, same box)"
for cfg in "PG_ATTN_SOFTMAX=1 PG_GEMM_EPI=1" "PG_ATTN_SOFTMAX=2 PG_GEMM_EPI=1" "PG_ATTN_SOFTMAX=2 PG_GEMM_EPI=2" "PG_ATTN_SOFTMAX=1 PG_GEMM_EPI=1"; do
  echo "-- $cfg"; env $cfg timeout 300 python bench.py --steps 3 --warmup 3 --no-other-workloads --no-cpu-baseline --no-other-modes 2> gpurun_out/l4_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['roofline']['kernel_ms_in_timed_region']
print(round(d['value']), 'mut/s', round(d['ms_per_step'],1),'ms/step', 'clk', d['clocks']['sm_mhz'], 'frac', round(d['roofline']['frac'],3), 'issued', round(d['roofline']['issued_frac'],3), {k: round(v['ms']/d['steps'],1) for k,v in c.items() if v['ms']>1})
" || tail -3 gpurun_out/l4_err.log; done 2>&1 | tee gpurun_out/l4_ab.txt
echo "== 5. model parity with the new defaults"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "blat or model_matches or golden_small or multi_site or esm2_3b_true_size_rows" 2>&1 | grep -v "^$" | tail -8 | tee gpurun_out/l5_model_parity.log
echo "== 6. ncu of the packed-softmax attention kernel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc4 -s 2 -c 1 -f -o gpurun_out/prof_r02_attn_tc4_x3_packed python scripts/prof_attn.py 3 0 > gpurun_out/l6_ncu_attn.log 2>&1; tail -1 gpurun_out/l6_ncu_attn.log
echo "== done"
