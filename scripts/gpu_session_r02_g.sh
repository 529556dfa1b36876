#!/bin/bash
# Round 2, GPU session G (2 GPUs): the 2-GPU tests (position partition, Tranception row sharding over NCCL) and the bench at N = 2.
mkdir -p gpurun_out
nvidia-smi -L
echo "== 1. two-GPU tests"; timeout 900 python -m pytest tests -m gpu -q -s -k "two_gpus" 2>&1 | grep -v "^$" | tail -6 | tee gpurun_out/g1_two_gpu_tests.log
echo "== 2. checks, verbose"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29681 scripts/check_position_partition.py 650m 2>&1 | grep "position-partition" | tee gpurun_out/g2_position_partition.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29682 scripts/check_tranception_sharding.py L 2>&1 | grep "tranception-row" | tee gpurun_out/g2_tranception_sharding.txt
echo "== 3. bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29683 bench.py --gpus 2 --no-other-modes > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; tail -2 gpurun_out/bench_r02_n2.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_n2.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "per_rank_ms")}, "e2e", d["e2e"]["value"], d["clocks"])
    for o in d.get("other_workloads", []):
        print({k: o.get(k) for k in ("value", "seconds", "per_rank_ms", "frac_of_peak", "precision_mode")}, o.get("config", "")[:40])
except Exception as e:
    print("bench parse failed", e)
PY
echo "== done"
