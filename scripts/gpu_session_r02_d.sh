#!/bin/bash
# Round 2, GPU session D: warp-uniform MMA issue (GEMM + attention), precision policy tests, un-gated seam test, ncu captures.
mkdir -p gpurun_out
echo "== 1. kernel tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "gemm or attention_matches or layernorm" 2>&1 | grep -v "^$" | tail -6 | tee gpurun_out/d1_kernels.log
echo "== 2. bench A/B (3 steps, headline mode only)"
for cfg in "" "PG_ATTN_TC3=1" "PG_GEMM_KCHUNK=640"; do
  echo "-- $cfg"; env $cfg timeout 300 python bench.py --steps 3 --warmup 3 --no-other-workloads --no-cpu-baseline --no-other-modes 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['roofline']['kernel_ms_in_timed_region']
print(round(d['value']), 'mut/s', round(d['ms_per_step'],1),'ms/step', 'clk', d['clocks']['sm_mhz'], 'frac', round(d['roofline']['frac'],3), 'issued', round(d['roofline']['issued_frac'],3), {k: round(v['ms']/d['steps'],1) for k,v in c.items() if v['ms']>1})
"; done 2>&1 | tee gpurun_out/d2_ab.log
echo "== 3. attention + gemm microbench"; timeout 200 python scripts/check_attention_impl.py 2>&1 | tail -6 | tee gpurun_out/d3_attn.log
timeout 300 python scripts/bench_gemm.py 1280 2>&1 | tee gpurun_out/d3_bench_gemm.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['gemm'], 'nseg', d['nseg'], d['ms'], 'ms', d['issued_tflops'], 'issued TF/s', d['issued_frac_of_burst_peak'])
"
echo "== 4. full GPU suite"; timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" > gpurun_out/d4_suite.log; tail -8 gpurun_out/d4_suite.log; grep -i "BLAT\|ESM2-3B\|tranception_L\|by sites\|trancepteve_L" gpurun_out/d4_suite.log | head -30
echo "== 5. ncu captures"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/prof_r02_gemm_fc1_f16f8_v2 python scripts/prof_gemm.py 2 fc1 > gpurun_out/d5_ncu_fc1.log 2>&1; tail -1 gpurun_out/d5_ncu_fc1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc4 -s 2 -c 1 -f -o gpurun_out/prof_r02_attn_tc4_x3 python scripts/prof_attn.py 3 0 > gpurun_out/d5_ncu_attn.log 2>&1; tail -1 gpurun_out/d5_ncu_attn.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_f16f8_bench_steps1.csv python bench.py --steps 1 --warmup 1 --no-other-workloads --no-cpu-baseline --no-other-modes > /dev/null 2>&1; wc -l gpurun_out/launches_r02_f16f8_bench_steps1.csv
echo "== 6. bench (full)"; timeout 900 python bench.py > gpurun_out/bench_r02_d.json 2> gpurun_out/bench_r02_d.err; tail -3 gpurun_out/bench_r02_d.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_d.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "precision_mode")}, "e2e", d["e2e"]["value"], d["clocks"])
    print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "issued_frac")}, d["roofline"]["secondary"])
    print("cats", {k: round(v["ms"], 1) for k, v in d["roofline"]["kernel_ms_in_timed_region"].items()})
    for o in d["other_precision_modes"]:
        print(o["precision_mode"], round(o["value"]), o["roofline"]["frac"], o["roofline"]["issued_frac"])
    for o in d.get("other_workloads", []):
        print({k: o.get(k) for k in ("value", "seconds", "algorithmic_tflops", "frac_of_peak", "precision_mode", "prefix_reuse_token_rows")}, o.get("config", "")[:40])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
echo "== done"
