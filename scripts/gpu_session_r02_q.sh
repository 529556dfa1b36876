#!/bin/bash
# round 2 session Q: delta-operand mode after hiding the base-row load latency (LayerNorm, GEMM), templated delta kernels
mkdir -p gpurun_out
nvidia-smi -L
echo "== 1. delta-mode tests + multi-site error growth + GEMM / attention kernel parity"
timeout 900 python -m pytest tests/test_gpu_delta.py tests/test_gpu_parity.py -m gpu -q -s -k "delta or multi_site or gemm_matches_fp64 or attention_matches or gemm_f16f8" 2>&1 | grep -v "^$" | tail -12 | tee gpurun_out/q1_delta_tests.log
echo "== 2. bench legs (3 steps): f16d / f16f8, same box"
for prec in f16d f16f8 f16d; do
  echo "-- --precision $prec"; timeout 400 python bench.py --steps 3 --warmup 3 --precision $prec --no-other-workloads --no-cpu-baseline --no-other-modes 2> gpurun_out/q2_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['roofline']['kernel_ms_in_timed_region']
print(round(d['value']), 'mut/s', round(d['ms_per_step'],1),'ms/step', 'e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], 'frac', round(d['roofline']['frac'],3), 'issued', round(d['roofline']['issued_frac'],3), {k: round(v['ms']/d['steps'],1) for k,v in c.items() if v['ms']>1})
" || tail -5 gpurun_out/q2_err.log; done 2>&1 | tee gpurun_out/q2_ab.txt
echo "== done"
