#!/bin/bash
# Round 2, GPU session C: new attention kernel (2 CTAs/SM), L2 look-ahead in the GEMM, Tranception prefix reuse, true-size goldens.
mkdir -p gpurun_out
echo "== 1. kernel tests (attention impl 0/1/2, prefix reuse)"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tranception.py -m gpu -q -s -k "attention_matches or prefix or true_size_tranception" 2>&1 | grep -v "^$" | tail -25 | tee gpurun_out/c1_kernels.log
echo "== 2. bench A/B (3 steps, headline mode only)"
for cfg in "" "PG_ATTN_TC3=1" "PG_GEMM_PREFETCH=0" "PG_GEMM_KCHUNK=1024" "PG_GEMM_PREFETCH=16"; do
  echo "-- $cfg"; env $cfg timeout 300 python bench.py --steps 3 --warmup 3 --no-other-workloads --no-cpu-baseline --no-other-modes 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['roofline']['kernel_ms_in_timed_region']
print(round(d['value']), 'mut/s', round(d['ms_per_step'],1),'ms/step', 'clk', d['clocks']['sm_mhz'], 'frac', round(d['roofline']['frac'],3), 'issued', round(d['roofline']['issued_frac'],3), {k: round(v['ms']/d['steps'],1) for k,v in c.items() if v['ms']>1})
"; done 2>&1 | tee gpurun_out/c2_ab.log
echo "== 3. attention microbench"; timeout 200 python scripts/check_attention_impl.py 2>&1 | tail -12 | tee gpurun_out/c3_attn.log
echo "== 4. full GPU suite"; timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" > gpurun_out/c4_suite.log; tail -6 gpurun_out/c4_suite.log; grep -i "BLAT\|ESM2-3B\|tranception_L\|prefix reuse" gpurun_out/c4_suite.log | head -30
echo "== 5. bench (full)"; timeout 900 python bench.py > gpurun_out/bench_r02_c.json 2> gpurun_out/bench_r02_c.err; tail -3 gpurun_out/bench_r02_c.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_c.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "precision_mode")}, "e2e", d["e2e"]["value"], d["clocks"])
    print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "issued_frac")}, d["roofline"]["secondary"])
    print("cats", {k: round(v["ms"], 1) for k, v in d["roofline"]["kernel_ms_in_timed_region"].items()})
    for o in d["other_precision_modes"]:
        print(o["precision_mode"], round(o["value"]), o["roofline"]["frac"], o["roofline"]["issued_frac"])
    for o in d.get("other_workloads", []):
        print({k: o.get(k) for k in ("value", "seconds", "algorithmic_tflops", "frac_of_peak")}, o.get("config", "")[:40])
except Exception as e:
    print("bench parse failed", e)
PY
echo "== done"
