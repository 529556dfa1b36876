"""torchrun --nproc-per-node N scripts/check_msa_partition.py: the masked positions of one alignment split over N GPUs
(MsaScorer.score_assay(shard=...)) must equal, bit for bit, the single-GPU scores rank 0 computes alone. Prints one line."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from proteingym_b200 import checkpoint, msa_engine, sharding, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    big = len(sys.argv) > 1 and sys.argv[1] == "msa1b"
    arch = synth.MSA_1B if big else synth.MsaArch(3, 256, 4, 512)
    conf = checkpoint.config_from_msa_synth(arch)
    state = sharding.broadcast_state(checkpoint.normalise_msa_synth_state(arch, synth.make_msa_state(arch, 0)) if rank == 0 else None, src=0,
                                     device=torch.device("cuda", local))
    seq = synth.random_protein(256 if big else 90, 1)
    rows = synth.random_alignment(seq, 128 if big else 12, 2)
    muts = synth.sample_mutants(seq, 400 if big else 300, 3, multi_frac=0.2)
    R, C = len(rows), len(seq) + 1
    sc = msa_engine.MsaScorer(conf, state, precision="f16f8", device=local, max_rows=msa_engine.default_max_rows(conf, R, C, want=4))
    torch.cuda.synchronize()
    dist.barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    split = sc.score_assay(rows, seq, muts, shard=(rank, world))
    t1.record(); torch.cuda.synchronize()
    ms_split = t0.elapsed_time(t1)
    ok = True
    if rank == 0:
        t0.record()
        alone = sc.score_assay(rows, seq, muts)
        t1.record(); torch.cuda.synchronize()
        ok = bool(np.array_equal(split, alone))
        print(f"msa-position-partition world={world} mutants={len(muts)} R={R} C={C} bit_identical={ok} ms_split={ms_split:.1f} "
              f"ms_alone={t0.elapsed_time(t1):.1f}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    sc.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
