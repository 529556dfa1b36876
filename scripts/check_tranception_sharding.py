"""torchrun --nproc-per-node N scripts/check_tranception_sharding.py: one Tranception assay scored with its sequence rows split over N
GPUs (TranceptionScorer.shard + one NCCL all-gather per direction, prefix reuse on) must equal the scores rank 0 computes alone to
fp32 round-off of the split summation (< 2e-6). Prints one line."""
import os
import sys

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from proteingym_b200 import synth  # noqa: E402
from proteingym_b200.tranception_engine import TranceptionScorer  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    big = len(sys.argv) > 1 and sys.argv[1] == "L"
    arch = synth.TRANCEPTION_L if big else synth.TranceptionArch(2, 256, 4, 512)
    st = {k[len("transformer."):]: v for k, v in synth.make_tranception_state(arch, 3).items() if k.startswith("transformer.")}
    cfg = {"n_embd": arch.embed_dim, "n_head": arch.heads, "n_layer": arch.layers, "n_ctx": arch.n_ctx, "n_inner": arch.ffn_dim, "vocab_size": 25}
    sc = TranceptionScorer(cfg, st, device=local)
    seq = synth.random_protein(400 if big else 300, 1)
    muts = synth.sample_mutants(seq, 400 if big else 150, 2, multi_frac=0.2)
    dms = pd.DataFrame({"mutant": muts, "mutated_sequence": [synth.apply_mutant(seq, m) for m in muts]})
    torch.cuda.synchronize()
    dist.barrier()
    sc.shard = (rank, world)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    split = sc.score_mutants(dms, seq)
    t1.record(); torch.cuda.synchronize()
    ms_split = t0.elapsed_time(t1)
    ok = True
    if rank == 0:
        sc.shard = None
        t0.record()
        alone = sc.score_mutants(dms, seq)
        t1.record(); torch.cuda.synchronize()
        d = max(np.abs(split[c].to_numpy(np.float64) - alone[c].to_numpy(np.float64)).max() for c in ("avg_score_L_to_R", "avg_score_R_to_L", "avg_score"))
        ok = bool(d < 2e-6) and list(split["mutated_sequence"]) == list(alone["mutated_sequence"])
        print(f"tranception-row-sharding world={world} mutants={len(muts)} L={len(seq)} max_abs_diff={d:.2e} ok={ok} ms_split={ms_split:.1f} "
              f"ms_alone={t0.elapsed_time(t1):.1f}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    sc.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
