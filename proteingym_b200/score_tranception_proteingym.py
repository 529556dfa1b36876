"""Drop-in for ``proteingym/baselines/tranception/score_tranception_proteingym.py`` (same flags, same output CSV:
``mutated_sequence, avg_score_L_to_R, avg_score_R_to_L, avg_score``; reference lines :18-45 flags, :57-77 DMS resolution,
:105-122 scoring + CSV). Inference-time retrieval for substitutions builds the unweighted MSA prior on the GPU (msa_prior.py, utils/msa_utils.py:63-138);
EVE sequence-weight files (MSA_processing) are not reproduced: pass a precomputed ``[L_full, 25]`` log-prior with ``--MSA_log_prior_npy``.
Additive flags: --precision, --device, --MSA_log_prior_npy."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import pandas as pd

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def create_parser():
    parser = argparse.ArgumentParser(description='Tranception scoring')
    parser.add_argument('--checkpoint', type=str, help='Path of Tranception model checkpoint')
    parser.add_argument('--model_framework', default='pytorch', type=str, help='Underlying framework [pytorch|JAX]')
    parser.add_argument('--batch_size_inference', default=20, type=int, help='Batch size for inference')
    parser.add_argument('--DMS_reference_file_path', default=None, type=str, help='Path to reference file with list of DMS to score')
    parser.add_argument('--DMS_index', default=0, type=int, help='Index of DMS assay in reference file')
    parser.add_argument('--target_seq', default=None, type=str, help='Full wild type sequence that is mutated in the DMS asssay')
    parser.add_argument('--DMS_file_name', default=None, type=str, help='Name of DMS assay file')
    parser.add_argument('--MSA_filename', default=None, type=str, help='Name of MSA (eg., a2m) file constructed on the wild type sequence')
    parser.add_argument('--MSA_weight_file_name', default=None, type=str, help='Weight of sequences in the MSA (optional)')
    parser.add_argument('--MSA_start', default=None, type=int, help='Sequence position that the MSA starts at (1-indexing)')
    parser.add_argument('--MSA_end', default=None, type=int, help='Sequence position that the MSA ends at (1-indexing)')
    parser.add_argument('--DMS_data_folder', type=str, help='Path to folder that contains all DMS assay datasets')
    parser.add_argument('--output_scores_folder', default='./', type=str, help='Name of folder to write model scores to')
    parser.add_argument('--deactivate_scoring_mirror', action='store_true', help='Whether to deactivate sequence scoring from both directions (Left->Right and Right->Left)')
    parser.add_argument('--indel_mode', action='store_true', help='Flag to be used when scoring insertions and deletions. Otherwise assumes substitutions')
    parser.add_argument('--scoring_window', default="optimal", type=str, help='Sequence window selection mode (when sequence length longer than model context size)')
    parser.add_argument('--num_workers', default=10, type=int, help='Number of workers for model scoring data loader')
    parser.add_argument('--inference_time_retrieval', action='store_true', help='Whether to perform inference-time retrieval')
    parser.add_argument('--retrieval_inference_weight', default=0.6, type=float, help='Coefficient (alpha) used when aggregating autoregressive transformer and retrieval')
    parser.add_argument('--MSA_folder', default='.', type=str, help='Path to MSA for neighborhood scoring')
    parser.add_argument('--MSA_weights_folder', default=None, type=str, help='Path to MSA weights for neighborhood scoring')
    parser.add_argument('--clustal_omega_location', default=None, type=str, help='Path to Clustal Omega (only needed with scoring indels with retrieval)')
    # additive
    parser.add_argument('--precision', default='f16x3', choices=['f16x3', 'f16'])
    parser.add_argument('--device', default=0, type=int)
    parser.add_argument('--MSA_log_prior_npy', default=None, type=str, help='precomputed [L_full, 25] log prior for --inference_time_retrieval')
    return parser


def main(argv=None):
    from proteingym_b200.tranception_engine import TranceptionScorer, load_tranception_checkpoint
    args = create_parser().parse_args(argv)
    MSA_start = MSA_end = None
    if args.DMS_reference_file_path:
        mapping = pd.read_csv(args.DMS_reference_file_path)
        DMS_id = mapping["DMS_id"][args.DMS_index]
        print("Compute scores for DMS: " + str(DMS_id))
        sel = mapping["DMS_id"] == DMS_id
        target_seq = mapping["target_seq"][sel].values[0].upper()
        DMS_file_name = mapping["DMS_filename"][sel].values[0]
        if args.inference_time_retrieval:
            MSA_start = int(mapping["MSA_start"][sel].values[0]) - 1
            MSA_end = int(mapping["MSA_end"][sel].values[0])
    else:
        target_seq = args.target_seq
        DMS_file_name = args.DMS_file_name
        DMS_id = DMS_file_name.split(".")[0]
        if args.inference_time_retrieval:
            MSA_start, MSA_end = args.MSA_start - 1, args.MSA_end
    log_prior = None
    if args.inference_time_retrieval:
        if args.indel_mode:
            raise NotImplementedError("retrieval for indels re-aligns the MSA with Clustal Omega (out of scope)")
        if args.MSA_log_prior_npy:
            log_prior = np.load(args.MSA_log_prior_npy)
        else:
            if args.MSA_weights_folder is not None:
                raise NotImplementedError("EVE sequence-weight files need MSA_processing (not reproduced); pass --MSA_log_prior_npy "
                                          "or omit --MSA_weights_folder for the unweighted prior")
            from proteingym_b200.msa_prior import msa_log_prior
            if args.DMS_reference_file_path:
                msa_file = args.MSA_folder + os.sep + mapping["MSA_filename"][args.DMS_index]
            else:
                msa_file = args.MSA_folder + os.sep + args.MSA_filename
            log_prior = msa_log_prior(msa_file, MSA_start, MSA_end, len(target_seq), device=args.device)  # model_pytorch.py:660-671
    config, state = load_tranception_checkpoint(args.checkpoint)
    scorer = TranceptionScorer(config, state, precision=args.precision, device=args.device)
    if not os.path.isdir(args.output_scores_folder):
        os.mkdir(args.output_scores_folder)
    DMS_data = pd.read_csv(args.DMS_data_folder + os.sep + DMS_file_name, low_memory=False)
    all_scores = scorer.score_mutants(DMS_data=DMS_data, target_seq=target_seq, scoring_mirror=not args.deactivate_scoring_mirror,
                                      indel_mode=args.indel_mode, scoring_window=args.scoring_window, log_prior=log_prior,
                                      retrieval_inference_weight=args.retrieval_inference_weight, MSA_start=MSA_start or 0, MSA_end=MSA_end)
    all_scores.to_csv(args.output_scores_folder + os.sep + DMS_id + ".csv", index=False)
    scorer.close()


if __name__ == '__main__':
    main()
