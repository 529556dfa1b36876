"""Drop-in for ``proteingym/baselines/esm/compute_fitness.py``: the ESM-1b / ESM-1v / ESM2 branch and the MSA Transformer branch.

Same flags, same DMS/mapping resolution, same output CSV (all input columns + one float column per checkpoint named by the
file stem + ``Ensemble_ESM1v`` when ``"ESM1v" in --model_type``; for the MSA Transformer one ``<checkpoint>_seed<s>`` column per seed +
``<checkpoint>_ensemble``), so ``scripts/scoring_DMS_zero_shot/scoring_ESM1v_substitutions.sh``, ``scoring_ESM2_substitutions.sh`` and
``scoring_MSA_transformer_substitutions.sh`` work with only the script path changed. Reference line numbers refer to
proteingym/baselines/esm/compute_fitness.py.

Not reproduced: ``--filter-msa`` (the reference shells out to the external hhfilter binary, :78-89) and pseudo-ppl with the MSA
Transformer (:406-418); both fail loudly. Additive flags: ``--precision``, ``--device``.
"""
from __future__ import annotations

import argparse
import os
import pathlib
import sys

import numpy as np
import pandas as pd

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


# (flags, kwargs) — the reference parser's option names, types, defaults, nargs and choices (checked one by one against a dump of
# the reference's create_parser() in tests/golden/esm_cli_flags.json); help texts are ours.
_FLAGS = [
    (("--model_type",), dict(type=str, default="MSA_transformer", nargs="+", help="MSA_transformer | ESM1v | ESM1b | ESM2")),
    (("--model-location",), dict(type=str, nargs="+", help="one or more local fair-esm .pt checkpoints; one output column each")),
    (("--sequence",), dict(type=str, help="wild-type sequence (filled from the mapping file when --dms_index is given)")),
    (("--dms-input",), dict(type=pathlib.Path, help="DMS CSV, or the folder holding the DMS CSVs when --dms_index is given")),
    (("--dms_index",), dict(type=int, help="row of the assay in --dms_mapping")),
    (("--dms_mapping",), dict(type=str, help="reference file (DMS_substitutions.csv)")),
    (("--mutation-col",), dict(type=str, default="mutant", help="column with mutants written like A24G or A24G:T30S")),
    (("--dms-output",), dict(type=pathlib.Path, help="folder for <DMS_id>.csv")),
    (("--offset-idx",), dict(type=int, default=1, help="position of the first residue in the mutant numbering")),
    (("--scoring-strategy",), dict(type=str, default="wt-marginals", choices=["wt-marginals", "pseudo-ppl", "masked-marginals"], help="")),
    (("--msa-path",), dict(type=pathlib.Path, help="(MSA Transformer only)")),
    (("--msa-sampling-strategy",), dict(type=str, default="sequence-reweighting", help="(MSA Transformer only)")),
    (("--msa-samples",), dict(type=int, default=400, help="(MSA Transformer only)")),
    (("--msa-weights-folder",), dict(type=str, default=None, help="(MSA Transformer only)")),
    (("--seeds",), dict(type=int, default=1, nargs="+", help="(MSA Transformer only)")),
    (("--filter-msa",), dict(action="store_true", help="(MSA Transformer only)")),
    (("--hhfilter-min-cov",), dict(type=int, default=75, help="(MSA Transformer only)")),
    (("--hhfilter-max-seq-id",), dict(type=int, default=90, help="(MSA Transformer only)")),
    (("--hhfilter-min-seq-id",), dict(type=int, default=0, help="(MSA Transformer only)")),
    (("--path-to-hhfilter",), dict(type=str, default="/n/groups/marks/software/hhsuite/hhsuite-3.3.0", help="(MSA Transformer only)")),
    (("--scoring-window",), dict(type=str, default="optimal", help="long sequences: optimal (1024-token window per position) | overlapping")),
    (("--overwrite-prior-scores",), dict(action="store_true", help="accepted for compatibility")),
    (("--target_seq",), dict(default=None, type=str, help="wild type when no mapping file is used")),
    (("--weight_file_name",), dict(default=None, type=str, help="(MSA Transformer only)")),
    (("--MSA_start",), dict(default=None, type=int, help="(MSA Transformer only)")),
    (("--MSA_end",), dict(default=None, type=int, help="(MSA Transformer only)")),
    (("--nogpu",), dict(action="store_true", help="rejected: this scorer is the GPU path")),
]


def create_parser():
    p = argparse.ArgumentParser(description="ESM masked-marginal / wt-marginal / pseudo-ppl DMS scoring on B200 (compute_fitness.py drop-in)")
    for flags, kw in _FLAGS:
        p.add_argument(*flags, **kw)
    # additive (not in the reference)
    p.add_argument("--precision", default="auto", choices=["auto", "f16d", "f16f8", "f16x3", "f16"],
                   help="tensor-core operand precision: auto (default) picks f16d (delta operands, 1 tensor-pipe unit; ESM-1b / ESM-1v "
                        "masked-marginals on one window), f16f8 (fp16 + e4m3 cross terms, 2 units) or f16x3 (3 units) by architecture, "
                        "model width and mutation depth (esm_engine.choose_precision); all three meet the 1e-3 parity bar; f16 (1 unit, "
                        "plain single pass) is fast and does not")
    p.add_argument("--device", type=int, default=0, help="CUDA device ordinal")
    return p


def resolve_assay(args):
    """DMS / mapping resolution of the reference's ``main`` (:286-343). Returns (df, mutant_col, offset_idx); for the MSA Transformer
    it also sets args.msa_path / args.MSA_start / args.MSA_end / args.msa_weight_file and trims args.sequence to the aligned range."""
    mutant_col = args.mutation_col
    msa = "MSA_transformer" in args.model_type
    args.msa_weight_file = None
    if args.dms_index is not None:
        mapping = pd.read_csv(args.dms_mapping)
        DMS_id = mapping["DMS_id"][args.dms_index]
        print("Compute scores for DMS: " + str(DMS_id))
        row = mapping[mapping["DMS_id"] == DMS_id]
        if len(row) == 0:
            raise ValueError("No mappings found for DMS: " + str(DMS_id))
        elif len(row) > 1:
            raise ValueError("Multiple mappings found for DMS: " + str(DMS_id))
        row = row.iloc[0].replace(np.nan, "")
        args.sequence = row["target_seq"].upper()
        args.dms_input = str(args.dms_input) + os.sep + row["DMS_filename"]
        mutant_col = row["DMS_mutant_column"] if "DMS_mutant_column" in mapping.columns else mutant_col
        args.dms_output = str(args.dms_output) + os.sep + DMS_id + ".csv"
        offset = row["start_idx"] if "start_idx" in mapping.columns and row["start_idx"] != "" else 1
        if msa:  # :308-326
            end = offset + len(args.sequence)
            if row["MSA_filename"] == "":
                raise ValueError("No MSA found for DMS: " + str(DMS_id))
            args.msa_path = str(args.msa_path) + os.sep + row["MSA_filename"]
            args.MSA_start = int(row["MSA_start"]) if "MSA_start" in mapping.columns else 1
            args.MSA_end = int(row["MSA_end"]) if "MSA_end" in mapping.columns else len(args.sequence)
            if "weight_file_name" in mapping.columns and args.msa_weights_folder is not None:
                args.msa_weight_file = args.msa_weights_folder + os.sep + row["weight_file_name"]
            if offset != args.MSA_start or end != args.MSA_end:
                args.sequence = args.sequence[args.MSA_start - 1:args.MSA_end]
    else:
        DMS_id = str(args.dms_input).split(os.sep)[-1].split(".csv")[0]
        args.dms_output = str(args.dms_output) + os.sep + DMS_id + ".csv"
        offset = args.offset_idx
        args.sequence = args.target_seq.upper()
        if msa:  # :333-340
            if args.MSA_start is None or args.MSA_end is None:
                if args.msa_path:
                    print("MSA start and end not provided -- Assuming the MSA is covering the full WT sequence")
                args.MSA_start, args.MSA_end = 1, len(args.target_seq)
            if args.msa_weights_folder is not None:
                args.msa_weight_file = args.msa_weights_folder + os.sep + args.weight_file_name
    df = pd.read_csv(args.dms_input)
    if len(df) == 0:
        raise ValueError("No rows found in the dataframe")
    print(f"df shape: {df.shape}", flush=True)
    return df, mutant_col, int(offset)


def score_model(scorer, args, df, mutant_col, offset_idx) -> np.ndarray:
    """One checkpoint's column (:433-529)."""
    seq = args.sequence
    n_tokens = len(seq) + 2
    muts = list(df[mutant_col])
    if args.scoring_strategy == "masked-marginals":
        if n_tokens > 1024 and args.scoring_window == "overlapping":
            print("Overlapping not yet implemented for masked-marginals")  # :496-498
            sys.exit(0)
        window = 1024 if args.scoring_window == "optimal" else max(n_tokens, 1024)
        return scorer.score_assay(seq, muts, offset_idx, model_window=window).astype(np.float64)
    if args.scoring_strategy == "wt-marginals":
        if n_tokens > 1024 and args.scoring_window == "overlapping":
            table = scorer.wt_marginal_table_overlapping(seq)
        else:
            table = scorer.wt_marginal_table(seq)
        return scorer.score_mutants(table, muts, seq, offset_idx).cpu().numpy().astype(np.float64)
    if args.scoring_strategy == "pseudo-ppl":
        if "mutated_sequence" not in df:
            from .synth import apply_mutant
            df["mutated_sequence"] = [apply_mutant(seq, m, offset_idx) for m in muts]
        return np.asarray([scorer.pseudo_ppl(s) for s in df["mutated_sequence"]], dtype=np.float64)
    raise ValueError(args.scoring_strategy)


def main_msa_transformer(args, df, mutant_col):
    """The MSA Transformer branch of the reference's model loop (:360-425) and its ensemble column (:538-542)."""
    from proteingym_b200 import msa_engine
    from proteingym_b200.checkpoint import load_msa_checkpoint
    from proteingym_b200.esm_engine import choose_precision
    assert args.scoring_strategy in ["masked-marginals", "pseudo-ppl"], "Zero-shot scoring strategy not supported with MSA Transformer"
    if args.scoring_strategy == "pseudo-ppl":
        raise NotImplementedError("pseudo-ppl with the MSA Transformer (compute_fitness.py:406-418) is not part of the B200 path")
    seeds = [args.seeds] if isinstance(args.seeds, int) else list(args.seeds)
    # under torchrun (WORLD_SIZE > 1) the masked positions of every seed are split over the ranks (one NCCL all-gather per seed)
    from proteingym_b200 import sharding
    rank, world = sharding.rank_world()
    shard = None
    if world > 1:
        import torch
        import torch.distributed as dist
        args.device = int(os.environ.get("LOCAL_RANK", args.device))
        torch.cuda.set_device(args.device)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", args.device))
        shard = (rank, world)
    offset_idx = args.MSA_start
    name = None
    for model_location in args.model_location:
        config, state, name = load_msa_checkpoint(model_location)
        processed = None
        if args.msa_sampling_strategy == "sequence-reweighting":
            processed = msa_engine.process_msa(str(args.msa_path), args.msa_weight_file, args.filter_msa, device=args.device)
        elif args.filter_msa:
            msa_engine.process_msa(str(args.msa_path), args.msa_weight_file, True)
        scorer = None
        precision = choose_precision(config, df[mutant_col], args.scoring_strategy) if args.precision == "auto" else args.precision
        for seed in seeds:
            col = f"{name}_seed{seed}"
            if os.path.exists(args.dms_output):  # :365-372
                prior = pd.read_csv(args.dms_output)
                if col in prior.columns and not args.overwrite_prior_scores:
                    print(f"Skipping seed {seed} as it is already in the dataframe")
                    df = prior
                    continue
            rows = msa_engine.sample_msa(filename=str(args.msa_path), nseq=args.msa_samples, sampling_strategy=args.msa_sampling_strategy,
                                         random_seed=seed, weight_filename=args.msa_weight_file, processed_msa=processed,
                                         device=args.device)
            R, C = len(rows), len(rows[0][1]) + 1
            print(f"Batch sizes: torch.Size([1, {R}, {C}])")
            need = msa_engine.default_max_rows(config, R, min(C, 1024), 1 if precision == "f16" else 2)
            if scorer is None or scorer.max_rows < R * min(C, 1024):
                if scorer is not None:
                    scorer.close()
                scorer = msa_engine.MsaScorer(config, state, precision=precision, device=args.device, max_rows=need)
                print("Scoring with {} and model {} (operand precision {})".format(args.scoring_strategy, name, precision))
            df[col] = scorer.score_assay(rows, args.sequence, list(df[mutant_col]), offset_idx, shard=shard).astype(np.float64)
            if shard is not None:
                if rank != 0:
                    sharding.barrier_if_distributed()  # rank 0 finishes writing the CSV before anyone looks at it again
                    continue  # every rank holds the scores; rank 0 writes
            if os.path.exists(args.dms_output) and not args.overwrite_prior_scores:  # :419-423
                prior = pd.read_csv(args.dms_output)
                assert col not in prior.columns, f"Column {col} already exists in {args.dms_output}"
                prior = prior.merge(df[[col, "mutant"]], on="mutant")
                prior.to_csv(args.dms_output, index=False)
                df = prior
            else:
                df.to_csv(args.dms_output, index=False)
            if shard is not None:
                sharding.barrier_if_distributed()
        if scorer is not None:
            scorer.close()
    if shard is not None and rank != 0:
        return
    df[f"{name}_ensemble"] = 0.0
    for seed in seeds:
        df[f"{name}_ensemble"] += df[f"{name}_seed{seed}"]
    df[f"{name}_ensemble"] /= len(seeds)
    df.to_csv(args.dms_output, index=False)


def main(args):
    from proteingym_b200.checkpoint import load_esm_checkpoint
    from proteingym_b200.esm_engine import EsmScorer, choose_precision
    os.makedirs(args.dms_output, exist_ok=True)  # the reference's exists + mkdir (:283), safe when several ranks start together
    print("Arguments:", args)
    if args.nogpu:
        raise RuntimeError("--nogpu: this scorer is the B200 path and has no CPU fallback (use the reference script)")
    df, mutant_col, offset_idx = resolve_assay(args)
    print("Starting model scoring")
    if "MSA_transformer" in args.model_type:
        return main_msa_transformer(args, df, mutant_col)
    for model_location in args.model_location:
        config, state, name = load_esm_checkpoint(model_location)
        if config.arch != "esm2" and len(args.sequence) + 2 > 1024 and args.scoring_strategy == "wt-marginals" \
                and args.scoring_window != "overlapping":
            raise ValueError(f"Sequence length {len(args.sequence) + 2} above maximum  sequence length of 1024")  # modules.py:256-260
        precision = (choose_precision(config, df[mutant_col], args.scoring_strategy, seq_len=len(args.sequence))
                     if args.precision == "auto" else args.precision)
        scorer = EsmScorer(config, state, precision=precision, device=args.device)
        print("Scoring with {} and model {} (operand precision {})".format(args.scoring_strategy, name, precision))
        df[name] = score_model(scorer, args, df, mutant_col, offset_idx)
        scorer.close()
    if "ESM1v" in args.model_type:  # :532-537
        df["Ensemble_ESM1v"] = 0.0
        for model_location in args.model_location:
            df["Ensemble_ESM1v"] += df[model_location.split("/")[-1].split(".")[0]]
        df["Ensemble_ESM1v"] /= len(args.model_location)
    df.to_csv(args.dms_output, index=False)


if __name__ == "__main__":
    main(create_parser().parse_args())
