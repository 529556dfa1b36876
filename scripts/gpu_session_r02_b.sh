#!/bin/bash
# Round 2, GPU session B: GEMM microbench over modes / chunk lengths, ncu captures of the default-mode GEMMs, new bench.py.
mkdir -p gpurun_out
echo "== 1. failed-last-time kernel tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "gemm_f16f8" 2>&1 | grep -v "^$" | tail -15 | tee gpurun_out/b1_gemm_tests.log
echo "== 2. GEMM microbench"; timeout 400 python scripts/bench_gemm.py 0,640,1024,1280 2>&1 | tee gpurun_out/b2_bench_gemm.jsonl | tail -50
echo "== 3. ncu fc1 / fc2 f16f8"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/prof_r02_gemm_fc1_f16f8 python scripts/prof_gemm.py 2 fc1 > gpurun_out/b3_ncu_fc1.log 2>&1; tail -2 gpurun_out/b3_ncu_fc1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/prof_r02_gemm_fc2_f16f8 python scripts/prof_gemm.py 2 fc2 > gpurun_out/b3_ncu_fc2.log 2>&1; tail -2 gpurun_out/b3_ncu_fc2.log
echo "== 4. bench (new layout, all workloads)"; timeout 900 python bench.py > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err; tail -5 gpurun_out/bench_r02_b.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_b.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "precision_mode")}, "e2e", d["e2e"]["value"], d["clocks"])
    print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "issued_frac", "timed_in")})
    print("cats", {k: round(v["ms"], 1) for k, v in d["roofline"]["kernel_ms_in_timed_region"].items()})
    for o in d["other_precision_modes"]:
        print(o["precision_mode"], round(o["value"]), o["roofline"]["frac"], o["roofline"]["issued_frac"])
    for o in d.get("other_workloads", []):
        print(json.dumps(o)[:900])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
echo "== 5. full GPU suite"; timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" > gpurun_out/b5_suite.log; tail -5 gpurun_out/b5_suite.log; grep -i "BLAT\|ESM2-3B\|f16f8 gemm" gpurun_out/b5_suite.log | head -30
echo "== done"
