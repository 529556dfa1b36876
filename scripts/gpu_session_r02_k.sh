#!/bin/bash
# round 2 session K (2 GPUs): 2-GPU tests incl. the MSA position partition, bench --gpus 2 with the strong-scaling leg
mkdir -p gpurun_out
nvidia-smi -L
echo "== 1. two-GPU tests"
timeout 900 python -m pytest tests -q -m gpu -k "two_gpu or 2_gpu or sharding_2gpu or two_gpus" 2>&1 | tail -4 | tee gpurun_out/k1_two_gpu_tests.log
echo "== 2. MSA position partition (MSA-1b size)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29675 scripts/check_msa_partition.py msa1b 2>&1 | grep "msa-position" | tee gpurun_out/k2_msa_partition.txt
echo "== 3. bench N=2"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29677 bench.py --gpus 2 > gpurun_out/bench_r02_final_n2.json 2> gpurun_out/bench_r02_final_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r02_final_n2.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','per_rank_ms')}, 'e2e', d['e2e']['value'], d.get('strong_scaling'))
for w in d.get('other_workloads',[]): print({k:w.get(k) for k in ('value','unit','per_rank_ms','frac_of_peak','error')}, w['config'][:40])
PY
echo "== done"
