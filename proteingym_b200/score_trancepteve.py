"""Drop-in for ``proteingym/baselines/trancepteve/score_trancepteve.py`` (flags :19-61, assay / MSA / EVE resolution :75-156, scoring
and CSV :176-190, coefficient log :201-208). Output: ``<output_scores_folder>/<DMS_id>.csv`` with ``mutated_sequence,
avg_score_L_to_R, avg_score_R_to_L, avg_score, mutant``. Retrieval for indels (Clustal Omega re-alignment) is not reproduced.
Additive flags: --precision, --device, --EVE_sampler (how a missing EVE log-prior cache is computed: the reference's weight-by-weight
draw order, or the same distribution through batched local reparameterisation — eve_prior.py)."""
from __future__ import annotations

import argparse
import os
import sys

import pandas as pd

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# option names, types and defaults of the reference parser; help texts are ours
_FLAGS = [
    ("--checkpoint", dict(type=str, help="Tranception checkpoint folder (config.json + pytorch_model.bin)")),
    ("--model_framework", dict(default="pytorch", type=str, help="accepted for compatibility")),
    ("--batch_size_inference", dict(default=20, type=int, help="accepted for compatibility (batching is sized by the workspace)")),
    ("--DMS_reference_file_path", dict(default=None, type=str, help="reference file listing the assays")),
    ("--DMS_index", dict(default=0, type=int, help="row of the assay in the reference file")),
    ("--target_seq", dict(default=None, type=str, help="wild type when no reference file is used")),
    ("--DMS_file_name", dict(default=None, type=str, help="assay CSV when no reference file is used")),
    ("--MSA_filename", dict(default=None, type=str, help="a2m alignment of the wild type")),
    ("--MSA_weight_file_name", dict(default=None, type=str, help="sequence weights (.npy) inside --MSA_weights_folder")),
    ("--MSA_start", dict(default=None, type=int, help="first position covered by the MSA, 1-based")),
    ("--MSA_end", dict(default=None, type=int, help="last position covered by the MSA, 1-based")),
    ("--UniprotID", dict(default=None, type=str, help="protein id used to find the EVE checkpoints")),
    ("--MSA_threshold_sequence_frac_gaps", dict(default=None, type=float, help="drop aligned sequences with more gaps than this fraction")),
    ("--MSA_threshold_focus_cols_frac_gaps", dict(default=None, type=float, help="columns with more gaps than this fraction are not focus columns")),
    ("--DMS_data_folder", dict(type=str, help="folder with the assay CSVs")),
    ("--output_scores_folder", dict(default="./", type=str, help="folder for <DMS_id>.csv")),
    ("--deactivate_scoring_mirror", dict(action="store_true", help="score left-to-right only")),
    ("--indel_mode", dict(action="store_true", help="rows are full mutated sequences (insertions / deletions)")),
    ("--scoring_window", dict(default="optimal", type=str, help="window selection for sequences longer than n_ctx - 2")),
    ("--num_workers", dict(default=8, type=int, help="accepted for compatibility")),
    ("--inference_time_retrieval_type", dict(default=None, type=str, help="None | Tranception | TranceptEVE")),
    ("--retrieval_weights_manual", dict(action="store_true", help="use the two weights below instead of the depth-based ones")),
    ("--retrieval_inference_MSA_weight", dict(default=0.5, type=float, help="alpha: weight of the MSA prior")),
    ("--retrieval_inference_EVE_weight", dict(default=0.5, type=float, help="beta: weight of the EVE prior")),
    ("--MSA_folder", dict(default=".", type=str, help="folder with the MSAs")),
    ("--MSA_weights_folder", dict(default=None, type=str, help="folder with the sequence-weight files")),
    ("--clustal_omega_location", dict(default=None, type=str, help="(indel retrieval only; not supported)")),
    ("--EVE_model_folder", dict(type=str, help="folder with the EVE checkpoints (<MSA name>_seed_<s> or <UniProt id>_seed_<s>)")),
    ("--EVE_seeds", dict(nargs="*", help="seeds of the EVE checkpoints to ensemble")),
    ("--EVE_num_samples_log_proba", dict(default=10, type=int, help="Monte-Carlo samples behind the EVE log prior")),
    ("--EVE_model_parameters_location", dict(default=None, type=str, help="json with the EVE encoder / decoder sizes")),
    ("--MSA_recalibrate_probas", dict(action="store_true", help="rescale the MSA prior to the transformer's temperature")),
    ("--EVE_recalibrate_probas", dict(action="store_true", help="rescale the EVE prior to the transformer's temperature")),
    ("--clinvar_scoring", dict(action="store_true", help="ClinVar input: merge the input columns back, separate coefficient log")),
]


def create_parser():
    parser = argparse.ArgumentParser(description="TranceptEVE scoring on B200 (score_trancepteve.py drop-in)")
    for flag, kw in _FLAGS:
        parser.add_argument(flag, **kw)
    parser.add_argument("--precision", default="f16f8", choices=["f16f8", "f16x3", "f16"])
    parser.add_argument("--device", default=0, type=int)
    parser.add_argument("--EVE_sampler", default="auto", choices=["auto", "stream", "local"])
    return parser


def main(argv=None):
    from proteingym_b200.tranception_engine import load_tranception_checkpoint
    from proteingym_b200.trancepteve_engine import TranceptEVEScorer
    args = create_parser().parse_args(argv)
    print(args)
    from proteingym_b200.score_tranception_proteingym import _finish_distributed, _init_distributed
    rank, world = _init_distributed(args)
    retrieval = args.inference_time_retrieval_type is not None
    kw = {}
    if args.DMS_reference_file_path:
        mapping = pd.read_csv(args.DMS_reference_file_path)
        DMS_id = mapping["DMS_id"][args.DMS_index]
        print("Compute scores for DMS: " + str(DMS_id))
        sel = mapping["DMS_id"] == DMS_id
        target_seq = mapping["target_seq"][sel].values[0].upper()
        DMS_file_name = mapping["DMS_filename"][sel].values[0]
        UniProt_ID = mapping["UniProt_ID"][sel].values[0] if "UniProt_ID" in mapping else "No ID"
        if retrieval:
            MSA_data_file = args.MSA_folder + os.sep + mapping["MSA_filename"][args.DMS_index] if args.MSA_folder is not None else None
            wfile = args.MSA_weights_folder + os.sep + mapping["weight_file_name"][sel].values[0] if args.MSA_weights_folder else None
            kw.update(MSA_start=int(mapping["MSA_start"][sel].values[0]) - 1, MSA_end=int(mapping["MSA_end"][sel].values[0]))
            seq_thr = float(mapping["MSA_threshold_sequence_frac_gaps"][sel].values[0]) if "MSA_threshold_sequence_frac_gaps" in mapping else 0.5
            col_thr = float(mapping["MSA_threshold_focus_cols_frac_gaps"][sel].values[0]) if "MSA_threshold_focus_cols_frac_gaps" in mapping else 1.0
            print("Sequence (fragment) gap threshold: " + str(seq_thr))
            print("Focus column gap threshold: " + str(col_thr))
    else:
        target_seq = args.target_seq
        DMS_file_name = args.DMS_file_name
        DMS_id = DMS_file_name.split(".")[0]
        UniProt_ID = args.UniprotID
        if retrieval:
            MSA_data_file = args.MSA_folder + os.sep + args.MSA_filename if args.MSA_folder is not None else None
            wfile = args.MSA_weights_folder + os.sep + args.MSA_weight_file_name if args.MSA_weights_folder is not None else None
            kw.update(MSA_start=args.MSA_start - 1, MSA_end=args.MSA_end)
            seq_thr, col_thr = args.MSA_threshold_sequence_frac_gaps, args.MSA_threshold_focus_cols_frac_gaps
    num_seeds = 0
    if retrieval:
        kw.update(inference_time_retrieval_type=args.inference_time_retrieval_type,
                  retrieval_aggregation_mode="aggregate_indel" if args.indel_mode else "aggregate_substitution",
                  MSA_filename=MSA_data_file, MSA_weight_file_name=wfile, MSA_threshold_sequence_frac_gaps=seq_thr,
                  MSA_threshold_focus_cols_frac_gaps=col_thr, retrieval_weights_manual=args.retrieval_weights_manual,
                  retrieval_inference_MSA_weight=args.retrieval_inference_MSA_weight,
                  retrieval_inference_EVE_weight=args.retrieval_inference_EVE_weight)
        if "TranceptEVE" in args.inference_time_retrieval_type:
            paths = []
            num_seeds = len(args.EVE_seeds)
            print("Number of distinct EVE models to be leveraged: {}".format(num_seeds))
            msa_stem = os.path.basename(MSA_data_file.split(".a2m")[0])
            for seed in args.EVE_seeds:
                if os.path.exists(f"{args.EVE_model_folder}/{msa_stem}_seed_{seed}"):
                    name = f"{msa_stem}_seed_{seed}"
                elif os.path.exists(f"{args.EVE_model_folder}/{UniProt_ID}_seed_{seed}"):
                    name = f"{UniProt_ID}_seed_{seed}"
                else:
                    print(f"No EVE Model available for {MSA_data_file} with random seed {seed} in {args.EVE_model_folder}. Exiting")
                    sys.exit(1)
                paths.append(args.EVE_model_folder + os.sep + name)
            kw.update(EVE_model_paths=paths, EVE_num_samples_log_proba=args.EVE_num_samples_log_proba,
                      EVE_model_parameters_location=args.EVE_model_parameters_location,
                      MSA_recalibrate_probas=args.MSA_recalibrate_probas, EVE_recalibrate_probas=args.EVE_recalibrate_probas)
        else:  # the reference leaves the config defaults in place here (config.py:30-31): MSA recalibration off, EVE on (no-op)
            kw.update(MSA_recalibrate_probas=False, EVE_recalibrate_probas=True)
    config, state = load_tranception_checkpoint(args.checkpoint)
    scorer = TranceptEVEScorer(config, state, full_target_seq=target_seq, scoring_window=args.scoring_window, precision=args.precision,
                               device=args.device, EVE_sampler=args.EVE_sampler, **kw)
    if world > 1:
        scorer.shard = (rank, world)
    if rank == 0 and not os.path.isdir(args.output_scores_folder):
        os.mkdir(args.output_scores_folder)
    DMS_data = pd.read_csv(args.DMS_data_folder + os.sep + DMS_file_name, low_memory=False)
    all_scores = scorer.score_mutants(DMS_data=DMS_data, target_seq=target_seq, scoring_mirror=not args.deactivate_scoring_mirror,
                                      batch_size_inference=args.batch_size_inference, num_workers=args.num_workers,
                                      indel_mode=args.indel_mode)
    if len(all_scores) > 0 and args.clinvar_scoring:
        all_scores = pd.merge(all_scores, DMS_data, how="left", on="mutant")
    if rank != 0:  # every rank holds the complete scores; one writer
        scorer.close()
        _finish_distributed(world)
        return
    all_scores.to_csv(args.output_scores_folder + os.sep + DMS_id + ".csv", index=False)
    log_name = "ClinVar_scoring_Tranception_20221130" if args.clinvar_scoring else "TranceptEVE_aggregation_coefficients_log"
    with open(log_name, "a+") as fh:  # the reference appends one line per assay to this file in the working directory
        if os.stat(log_name).st_size == 0:
            fh.write("DMS_id,num_mutants_scored,num_mutants_scored_no_na,processed_MSA_depth,retrieval_inference_MSA_weight,retrieval_inference_EVE_weight\n")
        fh.write(",".join(str(x) for x in [DMS_id, len(all_scores), len(all_scores.dropna()), scorer.MSA_processed_depth, scorer.EVE_processed_depth,
                                           scorer.retrieval_inference_MSA_weight, scorer.retrieval_inference_EVE_weight]) + "\n")
    scorer.close()
    _finish_distributed(world)


if __name__ == "__main__":
    main()
