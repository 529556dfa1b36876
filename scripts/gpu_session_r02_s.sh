#!/bin/bash
# round 2 session S: delta GEMM epilogue back to its first form (base row added in the epilogue, not as accumulator start); parity,
# bench leg, ncu capture of the delta fc1 GEMM
mkdir -p gpurun_out
nvidia-smi -L
echo "== 1. delta + GEMM parity"
timeout 600 python -m pytest tests/test_gpu_delta.py tests/test_gpu_parity.py -m gpu -q -s -k "delta or gemm_matches_fp64" 2>&1 | grep -E "BLAT|passed|failed" | tee gpurun_out/s1_tests.log
echo "== 2. bench leg f16d (3 steps)"
for prec in f16d f16d; do
  timeout 400 python bench.py --steps 3 --warmup 3 --precision $prec --no-other-workloads --no-cpu-baseline --no-other-modes 2> gpurun_out/s2_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['roofline']['kernel_ms_in_timed_region']
print(round(d['value']), 'mut/s', round(d['ms_per_step'],1),'ms/step', 'e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], 'frac', round(d['roofline']['frac'],3), 'issued', round(d['roofline']['issued_frac'],3), {k: round(v['ms']/d['steps'],1) for k,v in c.items() if v['ms']>1})
" || tail -5 gpurun_out/s2_err.log; done 2>&1 | tee gpurun_out/s2_bench.txt
echo "== 3. ncu of the delta fc1 GEMM inside a bench step"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel<\(int\)1, \(int\)1, \(int\)1>" -s 40 -c 1 -f -o gpurun_out/prof_r02_gemm_fc1_delta python bench.py --steps 1 --warmup 1 --no-other-workloads --no-cpu-baseline --no-other-modes > gpurun_out/s3_ncu_fc1.log 2>&1; tail -1 gpurun_out/s3_ncu_fc1.log | cut -c1-200
echo "== done"
