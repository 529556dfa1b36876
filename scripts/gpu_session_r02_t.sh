#!/bin/bash
# round 2 session T (last GPU minutes): fc1 delta epilogue with the post-activation base row requested before the GELU arithmetic
mkdir -p gpurun_out
echo "== 1. delta tests"
timeout 200 python -m pytest tests/test_gpu_delta.py -m gpu -q -s 2>&1 | grep -E "BLAT|passed|failed" | tee gpurun_out/t1_tests.log
echo "== 2. bench leg f16d (3 steps)"
timeout 200 python bench.py --steps 3 --warmup 3 --precision f16d --no-other-workloads --no-cpu-baseline --no-other-modes 2> gpurun_out/t2_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['roofline']['kernel_ms_in_timed_region']
print(round(d['value']), 'mut/s', round(d['ms_per_step'],1),'ms/step', 'e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], 'frac', round(d['roofline']['frac'],3), 'issued', round(d['roofline']['issued_frac'],3), {k: round(v['ms']/d['steps'],1) for k,v in c.items() if v['ms']>1})
" 2>&1 | tee gpurun_out/t2_bench.txt
echo "== done"
