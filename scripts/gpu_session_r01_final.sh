#!/bin/bash
# One-shot GPU session (round 1, last GPU minutes): default-path verification first, then the opt-in experiments, each bounded.
mkdir -p gpurun_out
echo "== 1. full GPU suite (defaults)"; timeout 420 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/s1_suite.log
echo "== 2. bench (defaults)"; timeout 240 python bench.py > gpurun_out/bench_final_default.json 2> gpurun_out/bench_final_default.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_final_default.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["other_precision_mode"]["value"], d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
echo "== 3. attention impl 4 qualification"; timeout 150 python scripts/check_attention_impl.py 4 2>&1 | tail -28 | tee gpurun_out/s3_attn4.log
echo "== 4. model-level parity with PG_ATTN_INPLACE=1"; PG_ATTN_INPLACE=1 timeout 240 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tranception.py -m gpu -q -k "golden or model_matches or fast_mode or score_mutants_matches or sequence_logprobs or boundary" 2>&1 | tail -4 | tee gpurun_out/s4_inplace_tests.log
echo "== 5. bench with PG_ATTN_INPLACE=1"; PG_ATTN_INPLACE=1 timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_final_inplace.json 2> gpurun_out/bench_final_inplace.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_final_inplace.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["other_precision_mode"]["value"], d["clocks"])
    print("attention ms x3:", d["roofline"]["kernel_ms_in_timed_region"]["attention"], " f16:", d["other_precision_mode"]["roofline"]["kernel_ms_in_timed_region"]["attention"])
except Exception as e:
    print("bench parse failed", e)
PY
echo "== 6. Tranception tests with PG_CONV_V2=1"; PG_CONV_V2=1 timeout 200 python -m pytest tests/test_gpu_tranception.py tests/test_gpu_trancepteve.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/s6_conv2_tests.log
echo "== 7. Tranception bench with PG_CONV_V2=1 (+ PG_ATTN_INPLACE=1)"; PG_CONV_V2=1 PG_BENCH_CASES=subs_L512_1000,subs_L1500_windowed_300 timeout 150 python scripts/bench_tranception.py 2>&1 | grep "^{" | tee gpurun_out/s7_tranception_conv2.jsonl
PG_CONV_V2=1 PG_ATTN_INPLACE=1 PG_BENCH_CASES=subs_L512_1000 timeout 100 python scripts/bench_tranception.py 2>&1 | grep "^{" | tee gpurun_out/s7_tranception_conv2_inplace.jsonl
echo "== done"
