"""Drop-in for ``proteingym/baselines/esm/compute_fitness.py`` (ESM-1b / ESM-1v / ESM2 branch).

Same flags, same DMS/mapping resolution, same output CSV (all input columns + one float column per checkpoint named by the
file stem + ``Ensemble_ESM1v`` when ``"ESM1v" in --model_type``), so ``scripts/scoring_DMS_zero_shot/scoring_ESM1v_substitutions.sh``
and ``scoring_ESM2_substitutions.sh`` work with only the script path changed. Reference line numbers refer to
proteingym/baselines/esm/compute_fitness.py.

Not reproduced (out of scope): the MSA Transformer branch (:360-425) and hhfilter/MSA sampling — selecting
``--model_type MSA_transformer`` raises NotImplementedError. Additive flags: ``--precision``, ``--device``.
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


def create_parser():
    """Flag surface of the reference parser (:100-238); defaults are identical."""
    p = argparse.ArgumentParser(description="Label a deep mutational scan with predictions from an ensemble of ESM-1v models.")
    p.add_argument("--model_type", type=str, help="MSA_transformer Vs ESM1v Vs ESM1b", default="MSA_transformer", nargs="+")
    p.add_argument("--model-location", type=str, nargs="+",
                   help="PyTorch model file OR name of pretrained model to download (see README for models)")
    p.add_argument("--sequence", type=str, help="Base sequence to which mutations were applied")
    p.add_argument("--dms-input", type=pathlib.Path, help="CSV file containing the deep mutational scan")
    p.add_argument("--dms_index", type=int, help="Index of DMS in mapping file")
    p.add_argument("--dms_mapping", type=str, help="Location of DMS_mapping")
    p.add_argument("--mutation-col", type=str, default="mutant",
                   help="column in the deep mutational scan labeling the mutation as 'AiB'")
    p.add_argument("--dms-output", type=pathlib.Path,
                   help="Output file containing the deep mutational scan along with predictions")
    p.add_argument("--offset-idx", type=int, default=1, help="Offset of the mutation positions in `--mutation-col`")
    p.add_argument("--scoring-strategy", type=str, default="wt-marginals",
                   choices=["wt-marginals", "pseudo-ppl", "masked-marginals"], help="")
    p.add_argument("--msa-path", type=pathlib.Path, help="path to MSA (required for MSA Transformer)")
    p.add_argument("--msa-sampling-strategy", type=str, default="sequence-reweighting",
                   help="Strategy to sample sequences from MSA [sequence-reweighting|random|first_x_rows]")
    p.add_argument("--msa-samples", type=int, default=400, help="number of sequences to randomly sample from the MSA")
    p.add_argument("--msa-weights-folder", type=str, default=None,
                   help="Folder with weights to sample MSA sequences in 'sequence-reweighting' scheme")
    p.add_argument("--seeds", type=int, default=1, help="Random seed used during training", nargs="+")
    p.add_argument("--filter-msa", action="store_true", help="Whether to use hhfilter to filter input MSA before sampling")
    p.add_argument("--hhfilter-min-cov", type=int, default=75, help="minimum coverage with query (%%)")
    p.add_argument("--hhfilter-max-seq-id", type=int, default=90, help="maximum pairwise identity (%%)")
    p.add_argument("--hhfilter-min-seq-id", type=int, default=0, help="minimum sequence identity with query (%%)")
    p.add_argument("--path-to-hhfilter", type=str, default="/n/groups/marks/software/hhsuite/hhsuite-3.3.0",
                   help="Path to hhfilter binaries")
    p.add_argument("--scoring-window", type=str, default="optimal", help="Approach to handle long sequences [optimal|overlapping]")
    p.add_argument("--overwrite-prior-scores", action="store_true", help="Whether to overwrite prior scores in the dataframe")
    p.add_argument("--target_seq", default=None, type=str, help="WT sequence mutated in the assay")
    p.add_argument("--weight_file_name", default=None, type=str)
    p.add_argument("--MSA_start", default=None, type=int)
    p.add_argument("--MSA_end", default=None, type=int)
    p.add_argument("--nogpu", action="store_true", help="Do not use GPU even if available")
    # additive (not in the reference)
    p.add_argument("--precision", default="f16x3", choices=["f16x3", "f16"],
                   help="tensor-core operand precision: f16x3 meets the 1e-3 parity bar (default); f16 is ~2x faster")
    p.add_argument("--device", type=int, default=0, help="CUDA device ordinal")
    return p


def resolve_assay(args):
    """DMS / mapping resolution of the reference's ``main`` (:286-343). Returns (df, mutant_col, offset_idx)."""
    mutant_col = args.mutation_col
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
    else:
        DMS_id = str(args.dms_input).split(os.sep)[-1].split(".csv")[0]
        args.dms_output = str(args.dms_output) + os.sep + DMS_id + ".csv"
        offset = args.offset_idx
        args.sequence = args.target_seq.upper()
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


def main(args):
    from proteingym_b200.checkpoint import load_esm_checkpoint
    from proteingym_b200.esm_engine import EsmScorer
    if not os.path.exists(args.dms_output):
        os.mkdir(args.dms_output)
    print("Arguments:", args)
    if "MSA_transformer" in args.model_type:
        raise NotImplementedError("MSA Transformer scoring is outside the B200 hot path (use the reference script)")
    if args.nogpu:
        raise RuntimeError("--nogpu: this scorer is the B200 path and has no CPU fallback (use the reference script)")
    df, mutant_col, offset_idx = resolve_assay(args)
    print("Starting model scoring")
    for model_location in args.model_location:
        config, state, name = load_esm_checkpoint(model_location)
        if config.arch != "esm2" and len(args.sequence) + 2 > 1024 and args.scoring_strategy == "wt-marginals" \
                and args.scoring_window != "overlapping":
            raise ValueError(f"Sequence length {len(args.sequence) + 2} above maximum  sequence length of 1024")  # modules.py:256-260
        scorer = EsmScorer(config, state, precision=args.precision, device=args.device)
        print("Scoring with {} and model {}".format(args.scoring_strategy, name))
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
