#!/bin/bash
# Round 2, GPU session F: everything at its defaults (CTA-pair GEMM, attention_tc4, auto precision): suite, bench, profiles.
mkdir -p gpurun_out
echo "== 1. full GPU suite"; timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" > gpurun_out/f1_suite.log; tail -6 gpurun_out/f1_suite.log; grep -i "BLAT\|ESM2-3B\|tranception_L\|by sites" gpurun_out/f1_suite.log | head -20
echo "== 2. GEMM microbench"; timeout 300 python scripts/bench_gemm.py 1280 > gpurun_out/f2_bench_gemm.jsonl 2>&1; python - <<'PY'
import json
for l in open("gpurun_out/f2_bench_gemm.jsonl"):
    try: d = json.loads(l)
    except Exception: continue
    if d["nseg"] == 2: print(d["gemm"], "cta2", d["cta2"], d["ms"], "ms", d["issued_tflops"], d["issued_frac_of_burst_peak"])
PY
echo "== 3. ncu: pair GEMM fc1/fc2 (f16f8), launch list"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/prof_r02_gemm_fc1_f16f8_cta2 python scripts/prof_gemm.py 2 fc1 > gpurun_out/f3_ncu_fc1.log 2>&1; tail -1 gpurun_out/f3_ncu_fc1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/prof_r02_gemm_fc2_f16f8_cta2 python scripts/prof_gemm.py 2 fc2 > gpurun_out/f3_ncu_fc2.log 2>&1; tail -1 gpurun_out/f3_ncu_fc2.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_final_bench_steps1.csv python bench.py --steps 1 --warmup 1 --no-other-workloads --no-cpu-baseline --no-other-modes > /dev/null 2>&1; wc -l gpurun_out/launches_r02_final_bench_steps1.csv
echo "== 4. Tranception bench (reuse on)"; PG_BENCH_PRECS=f16f8 timeout 300 python scripts/bench_tranception.py 2>&1 | grep "^{" | tee gpurun_out/f4_tranception.jsonl
echo "== 5. bench (full, defaults)"; timeout 900 python bench.py > gpurun_out/bench_r02_f.json 2> gpurun_out/bench_r02_f.err; tail -3 gpurun_out/bench_r02_f.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_f.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "precision_mode")}, "e2e", d["e2e"]["value"], d["clocks"])
    print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "issued_frac")}, d["roofline"]["secondary"])
    print("cats", {k: round(v["ms"], 1) for k, v in d["roofline"]["kernel_ms_in_timed_region"].items()})
    for o in d["other_precision_modes"]:
        print(o["precision_mode"], round(o["value"]), o["roofline"]["frac"], o["roofline"]["issued_frac"], o["clocks"]["sm_mhz"])
    for o in d.get("other_workloads", []):
        print({k: o.get(k) for k in ("value", "seconds", "algorithmic_tflops", "frac_of_peak", "precision_mode", "prefix_reuse_token_rows")}, o.get("config", "")[:40])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
echo "== 6. reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2> /dev/null | cut -c1-600
echo "== done"
