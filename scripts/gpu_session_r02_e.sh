#!/bin/bash
# Round 2, GPU session E: CTA-pair (cta_group::2) GEMM first light, Tranception prefix-reuse timing.
mkdir -p gpurun_out
echo "== 1. GEMM tests, single-CTA then CTA-pair"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm and not f16f8 and cta2" 2>&1 | tail -4
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gemm_matches_fp64 or gemm_f16f8" 2>&1 | grep -v "^$" | tail -12 | tee gpurun_out/e1_gemm_cta2.log
echo "== 2. GEMM microbench single vs pair"; timeout 300 python scripts/bench_gemm.py 1280 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['gemm'], 'nseg', d['nseg'], 'cta2', d.get('cta2'), d['ms'], 'ms', d['issued_tflops'], 'issued TF/s', d['issued_frac_of_burst_peak'])
" | tee gpurun_out/e2_bench_gemm.log
echo "== 3. bench A/B"
for cfg in "" "PG_GEMM_CTA2=1"; do
  echo "-- $cfg"; env $cfg timeout 300 python bench.py --steps 3 --warmup 3 --no-other-workloads --no-cpu-baseline --no-other-modes 2> gpurun_out/e3_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['roofline']['kernel_ms_in_timed_region']
print(round(d['value']), 'mut/s', round(d['ms_per_step'],1),'ms/step', 'clk', d['clocks']['sm_mhz'], 'frac', round(d['roofline']['frac'],3), 'issued', round(d['roofline']['issued_frac'],3), {k: round(v['ms']/d['steps'],1) for k,v in c.items() if v['ms']>1})
" || tail -3 gpurun_out/e3_err.log; done 2>&1 | tee gpurun_out/e3_ab.log
echo "== 4. model parity with the pair kernel"; PG_GEMM_CTA2=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "blat or model_matches or golden_small or multi_site" 2>&1 | grep -v "^$" | tail -8 | tee gpurun_out/e4_parity_cta2.log
echo "== 5. Tranception prefix reuse on/off"
for r in 1 0; do echo "-- PG_PREFIX_REUSE=$r"; PG_PREFIX_REUSE=$r PG_BENCH_PRECS=f16f8 PG_BENCH_CASES=subs_L512_1000,subs_L1500_windowed_300 timeout 200 python scripts/bench_tranception.py 2>&1 | grep "^{"; done | tee gpurun_out/e5_prefix.jsonl
echo "== done"
