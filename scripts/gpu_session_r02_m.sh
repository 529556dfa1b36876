#!/bin/bash
# round 2 session M: attention softmax version 3 (exponent + packing of a 32-column chunk in one basic block, MUFUs interleaved)
mkdir -p gpurun_out
nvidia-smi -L
echo "== 1. attention: 19 parity cases + isolated throughput, softmax versions 1 / 2 / 3"
for v in 1 2 3; do echo "-- PG_ATTN_SOFTMAX=$v"; PG_ATTN_SOFTMAX=$v timeout 300 python scripts/check_attention_impl.py 0 2>&1 | grep -v "  ok" ; done | tee gpurun_out/m1_attention_ab.txt
echo "== 2. kernel parity (defaults), incl. causal / ALiBi / prefix cases"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tranception.py -m gpu -q -x -k "attention or prefix" 2>&1 | tail -3 | tee gpurun_out/m2_kernel_parity.log
echo "== 3. bench A/B (3 steps each, same box)"
for cfg in "PG_ATTN_SOFTMAX=1" "PG_ATTN_SOFTMAX=3" "PG_ATTN_SOFTMAX=2" "PG_ATTN_SOFTMAX=3"; do
  echo "-- $cfg"; env $cfg timeout 300 python bench.py --steps 3 --warmup 3 --no-other-workloads --no-cpu-baseline --no-other-modes 2> gpurun_out/m3_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['roofline']['kernel_ms_in_timed_region']
print(round(d['value']), 'mut/s', round(d['ms_per_step'],1),'ms/step', 'clk', d['clocks']['sm_mhz'], 'frac', round(d['roofline']['frac'],3), 'issued', round(d['roofline']['issued_frac'],3), {k: round(v['ms']/d['steps'],1) for k,v in c.items() if v['ms']>1})
" || tail -3 gpurun_out/m3_err.log; done 2>&1 | tee gpurun_out/m3_ab.txt
echo "== 4. ncu of the version-3 attention kernel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc4 -s 2 -c 1 -f -o gpurun_out/prof_r02_attn_tc4_x3_sv3 python scripts/prof_attn.py 3 0 > gpurun_out/m4_ncu_attn.log 2>&1; tail -1 gpurun_out/m4_ncu_attn.log
echo "== done"
