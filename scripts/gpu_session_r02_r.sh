#!/bin/bash
# round 2 session R: final evidence at HEAD with the delta-operand mode as the headline — full GPU suite, smoke, default bench,
# launch list and ncu captures of the delta GEMM / attention kernels
mkdir -p gpurun_out
nvidia-smi -L
echo "== 1. full GPU suite"
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | grep -v "^$" > gpurun_out/r1_suite.log; tail -4 gpurun_out/r1_suite.log
timeout 300 python -m pytest tests/test_gpu_delta.py tests/test_gpu_parity.py -q -m gpu -s -k "delta_mode_golden or multi_site" 2>&1 | grep -E "BLAT|by sites|passed|failed" | tee gpurun_out/r1_delta_numbers.txt
echo "== 2. smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== 3. bench (defaults)"
timeout 1500 python bench.py > gpurun_out/bench_r02_final4_n1.json 2> gpurun_out/bench_r02_final4_n1.err; tail -3 gpurun_out/bench_r02_final4_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_final4_n1.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "precision_mode")}, "e2e", d["e2e"]["value"], d["clocks"])
    print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "issued_frac")}, d["roofline"]["secondary"])
    print("cats", {k: round(v["ms"] / d["steps"], 1) for k, v in d["roofline"]["kernel_ms_in_timed_region"].items()})
    for o in d["other_precision_modes"]:
        print(o["precision_mode"], round(o["value"]), o["roofline"]["frac"], o["roofline"]["issued_frac"], o["clocks"]["sm_mhz"])
    for o in d.get("other_workloads", []):
        print({k: o.get(k) for k in ("value", "seconds", "algorithmic_tflops", "frac_of_peak", "precision_mode", "error")}, o.get("config", "")[:40])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
echo "== 4. launch list + ncu of the delta kernels (inside a bench step)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_final4_bench_steps1.csv python bench.py --steps 1 --warmup 1 --no-other-workloads --no-cpu-baseline --no-other-modes > /dev/null 2>&1; wc -l gpurun_out/launches_r02_final4_bench_steps1.csv
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel<1, 1, 1>" -s 40 -c 1 -f -o gpurun_out/prof_r02_gemm_fc1_delta python bench.py --steps 1 --warmup 1 --no-other-workloads --no-cpu-baseline --no-other-modes > gpurun_out/r4_ncu_fc1.log 2>&1; tail -1 gpurun_out/r4_ncu_fc1.log
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel<2, 1, 1>" -s 81 -c 1 -f -o gpurun_out/prof_r02_gemm_fc2_delta python bench.py --steps 1 --warmup 1 --no-other-workloads --no-cpu-baseline --no-other-modes > gpurun_out/r4_ncu_fc2.log 2>&1; tail -1 gpurun_out/r4_ncu_fc2.log
echo "== done"
