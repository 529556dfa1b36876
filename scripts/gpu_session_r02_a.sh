#!/bin/bash
# Round 2, GPU session A: first light of the chunked / fp16+e4m3 GEMM, then the whole suite, accumulation probe, bench.
mkdir -p gpurun_out
echo "== 1. kernel-level tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "gemm or layernorm or attention_matches" 2>&1 | tail -25 | tee gpurun_out/a1_kernels.log
echo "== 2. accumulation probe vs chunk size"
for kc in 0 2048 1024 512; do echo "-- PG_GEMM_KCHUNK=$kc"; PG_GEMM_KCHUNK=$kc timeout 120 python scripts/accum_probe.py 2>&1 | tail -6; done | tee gpurun_out/a2_accum_probe.log
echo "== 3. full GPU suite"; timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -40 | tee gpurun_out/a3_suite.log
echo "== 4. bench"; timeout 400 python bench.py > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; tail -3 gpurun_out/bench_r02_a.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_a.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "precision_mode")}, "e2e", d["e2e"]["value"], d["clocks"])
    print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "issued_frac")})
    print("cats", d["roofline"]["kernel_ms_in_timed_region"])
    for o in d["other_precision_modes"]:
        print(o["precision_mode"], o["value"], o["roofline"]["frac"], o["roofline"]["kernel_ms_in_timed_region"])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
echo "== done"
