#!/bin/bash
# round 2 session O: final evidence at HEAD — full GPU suite, smoke, default bench, balanced-pass A/B, launch list + ncu of the GEMMs
mkdir -p gpurun_out
nvidia-smi -L
echo "== 1. full GPU suite"
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | grep -v "^$" > gpurun_out/o1_suite.log; tail -4 gpurun_out/o1_suite.log
echo "== 2. smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== 3. passes A/B (3 steps each, same box): 255+255+2 vs 171+171+170 sequences per pass"
for cfg in "PG_BALANCED_PASSES=0" "PG_BALANCED_PASSES=1" "PG_BALANCED_PASSES=0" "PG_BALANCED_PASSES=1"; do
  echo "-- $cfg"; env $cfg timeout 300 python bench.py --steps 3 --warmup 3 --no-other-workloads --no-cpu-baseline --no-other-modes 2> gpurun_out/o3_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['roofline']['kernel_ms_in_timed_region']
print(round(d['value']), 'mut/s', round(d['ms_per_step'],1),'ms/step', 'clk', d['clocks']['sm_mhz'], 'frac', round(d['roofline']['frac'],3), 'issued', round(d['roofline']['issued_frac'],3), {k: round(v['ms']/d['steps'],1) for k,v in c.items() if v['ms']>1})
" || tail -3 gpurun_out/o3_err.log; done 2>&1 | tee gpurun_out/o3_ab.txt
echo "== 4. bench (defaults)"
timeout 1500 python bench.py > gpurun_out/bench_r02_final3_n1.json 2> gpurun_out/bench_r02_final3_n1.err; tail -c 400 gpurun_out/bench_r02_final3_n1.json; echo
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_final3_n1.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "precision_mode")}, "e2e", d["e2e"]["value"], d["clocks"])
    print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "issued_frac")}, d["roofline"]["secondary"])
    print("cats", {k: round(v["ms"], 1) for k, v in d["roofline"]["kernel_ms_in_timed_region"].items()})
    for o in d["other_precision_modes"]:
        print(o["precision_mode"], round(o["value"]), o["roofline"]["frac"], o["roofline"]["issued_frac"], o["clocks"]["sm_mhz"])
    for o in d.get("other_workloads", []):
        print({k: o.get(k) for k in ("value", "seconds", "algorithmic_tflops", "frac_of_peak", "precision_mode", "error")}, o.get("config", "")[:40])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
echo "== 5. launch list + ncu of the pair GEMMs"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_final3_bench_steps1.csv python bench.py --steps 1 --warmup 1 --no-other-workloads --no-cpu-baseline --no-other-modes > /dev/null 2>&1; wc -l gpurun_out/launches_r02_final3_bench_steps1.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/prof_r02_gemm_fc1_f16f8_cta2_v2 python scripts/prof_gemm.py 2 fc1 > gpurun_out/o5_ncu_fc1.log 2>&1; tail -1 gpurun_out/o5_ncu_fc1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/prof_r02_gemm_fc2_f16f8_cta2_v2 python scripts/prof_gemm.py 2 fc2 > gpurun_out/o5_ncu_fc2.log 2>&1; tail -1 gpurun_out/o5_ncu_fc2.log
echo "== done"
