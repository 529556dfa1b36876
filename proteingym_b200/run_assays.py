"""Multi-assay, multi-GPU driver: score every assay of a ProteinGym reference file with ESM masked-marginals on all the
GPUs of one box. Replaces the SLURM-array pattern of scripts/scoring_DMS_zero_shot/scoring_ESM*_substitutions.sh (one
``compute_fitness.py --dms_index N`` process per assay) by one process per GPU:

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m proteingym_b200.run_assays \\
        --model-location ckpt1.pt ckpt2.pt ... --model_type ESM1v --dms_mapping reference_files/DMS_substitutions.csv \\
        --dms-input DMS_ProteinGym_substitutions --dms-output out/

Weights are read on rank 0 and NCCL-broadcast. ``--partition assays`` (default when there are at least as many assays as
GPUs): assays are assigned by LPT on the analytic cost (sharding.assay_cost); each rank writes the CSVs of its own assays (same
files compute_fitness.py would write); no collective runs on the data path. ``--partition positions`` (default otherwise — few big
assays, single-assay latency): all ranks work on every assay, each on a contiguous chunk of its masked positions, and one
all-gather of the [P, 33] log-prob rows (<= 135 KB) completes the table; rank 0 writes the CSVs. Scores are bit-identical in both
modes and for any world size. Rank 0 gathers an (assay, seconds, mutants) summary."""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import pandas as pd
import torch

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None):
    from proteingym_b200 import sharding
    from proteingym_b200.checkpoint import checkpoint_column_name, load_esm_checkpoint
    from proteingym_b200.esm_engine import EsmScorer, choose_precision
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-location", nargs="+", required=True)
    ap.add_argument("--model_type", nargs="+", default=["ESM1v"])
    ap.add_argument("--dms_mapping", required=True)
    ap.add_argument("--dms-input", required=True)
    ap.add_argument("--dms-output", required=True)
    ap.add_argument("--mutation-col", default="mutant")
    ap.add_argument("--precision", default="auto", choices=["auto", "f16d", "f16f8", "f16x3", "f16"],
                    help="auto: per assay, the cheapest operand scheme that meets 1e-3 (esm_engine.choose_precision)")
    ap.add_argument("--indices", type=int, nargs="*", default=None, help="subset of dms_index values (default: all rows)")
    ap.add_argument("--partition", default="auto", choices=["auto", "assays", "positions"],
                    help="what is split across GPUs: whole assays (LPT) or the masked positions of each assay")
    a = ap.parse_args(argv)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mapping = pd.read_csv(a.dms_mapping)
    idx = list(range(len(mapping))) if a.indices is None else a.indices
    os.makedirs(a.dms_output, exist_ok=True)
    frames = {}
    summary = {}
    for ci, ckpt in enumerate(a.model_location):
        conf = state = None
        if rank == 0:
            conf, state, name = load_esm_checkpoint(ckpt)
        if dist is not None:
            box = [conf, name] if rank == 0 else [None, None]
            dist.broadcast_object_list(box, src=0)
            conf, name = box
            state = sharding.broadcast_state(state, src=0, device=torch.device("cuda", local))
        if ci == 0:
            # ONE partition for the whole run, from the first checkpoint's architecture: every column of an assay's CSV must be
            # produced by the same rank (a per-checkpoint LPT could move an assay between ranks when architectures are mixed).
            # Masked positions actually run = unique mutated positions <= min(L, number of mutants) when the mapping says how many.
            def n_pos(i):
                L = len(str(mapping["target_seq"][i]))
                if "DMS_total_number_mutants" in mapping.columns and not pd.isna(mapping["DMS_total_number_mutants"][i]):
                    return min(L, int(mapping["DMS_total_number_mutants"][i]))
                return L
            costs = [sharding.assay_cost(len(str(mapping["target_seq"][i])), conf.layers, conf.embed_dim, conf.ffn_dim, n_pos(i))
                     for i in idx]
            by_pos = world > 1 and (a.partition == "positions" or (a.partition == "auto" and len(idx) < world))
            mine = list(idx) if by_pos else [idx[j] for j in sharding.lpt_assign(costs, world)[rank]]
        scorers = {}

        def scorer_for(prec):  # at most two handles per checkpoint (f16f8 and f16x3), created on first use
            if prec not in scorers:
                scorers[prec] = EsmScorer(conf, state, precision=prec, device=local)
            return scorers[prec]
        for i in mine:
            row = mapping.iloc[i].replace(np.nan, "")
            seq = row["target_seq"].upper()
            col = row["DMS_mutant_column"] if "DMS_mutant_column" in mapping.columns else a.mutation_col
            off = int(row["start_idx"]) if "start_idx" in mapping.columns and row["start_idx"] != "" else 1
            if i not in frames:
                frames[i] = pd.read_csv(os.path.join(a.dms_input, row["DMS_filename"]))
            t0 = time.time()
            muts = list(frames[i][col])
            scorer = scorer_for(choose_precision(conf, muts, seq_len=len(seq)) if a.precision == "auto" else a.precision)
            frames[i][name] = scorer.score_assay(seq, muts, off, shard=(rank, world) if by_pos else None).astype(np.float64)
            if not by_pos or rank == 0:
                summary[(i, name)] = (time.time() - t0, len(frames[i]))
        for sc in scorers.values():
            sc.close()
        del state
    for i, df in frames.items():
        if by_pos and rank != 0:  # every rank holds the same scores; one writer
            continue
        if "ESM1v" in a.model_type:
            names = [checkpoint_column_name(c) for c in a.model_location]
            df["Ensemble_ESM1v"] = sum(df[n] for n in names) / len(names)
        df.to_csv(os.path.join(a.dms_output, str(mapping["DMS_id"][i]) + ".csv"), index=False)
    rows = [(int(i), n, float(s), int(m), rank) for (i, n), (s, m) in summary.items()]
    if dist is not None:
        allrows = [None] * world
        dist.all_gather_object(allrows, rows)
        rows = [r for rr in allrows for r in rr]
        dist.destroy_process_group()
    if rank == 0:
        rep = pd.DataFrame(rows, columns=["dms_index", "checkpoint", "seconds", "mutants", "rank"]).sort_values(["dms_index", "checkpoint"])
        rep.to_csv(os.path.join(a.dms_output, "_run_assays_summary.csv"), index=False)
        print(f"scored {rep['mutants'].sum()} mutant-checkpoint rows over {rep['dms_index'].nunique()} assays in "
              f"{rep.groupby('rank')['seconds'].sum().max():.1f} s (slowest rank)")


if __name__ == "__main__":
    main()
