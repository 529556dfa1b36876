"""Drop-in for ``proteingym/baselines/tranception/score_tranception_proteingym.py`` (same flags, same output CSV:
``mutated_sequence, avg_score_L_to_R, avg_score_R_to_L, avg_score``; reference lines :18-45 flags, :57-77 DMS resolution,
:105-122 scoring + CSV). Inference-time retrieval for substitutions builds the MSA prior on the GPU (msa_processing.get_msa_prior =
utils/msa_utils.py:63-138), with EVE-style sequence weights read from ``--MSA_weights_folder`` through the MSA_processing mirror when
given. Additive flags: --precision, --device, --MSA_log_prior_npy (a precomputed ``[L_full, 25]`` log prior). Launched with torchrun
(``--nproc-per-node N``) the assay's sequence rows are split over N GPUs (see ``_init_distributed``)."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import pandas as pd

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


# option names, types and defaults of the reference parser (score_tranception_proteingym.py:18-45); help texts are ours
_FLAGS = [
    ("--checkpoint", dict(type=str, help="HF checkpoint folder (config.json + pytorch_model.bin)")),
    ("--model_framework", dict(default="pytorch", type=str, help="accepted for compatibility")),
    ("--batch_size_inference", dict(default=20, type=int, help="accepted for compatibility (batching is sized by the workspace)")),
    ("--DMS_reference_file_path", dict(default=None, type=str, help="reference file listing the assays")),
    ("--DMS_index", dict(default=0, type=int, help="row of the assay in the reference file")),
    ("--target_seq", dict(default=None, type=str, help="wild type when no reference file is used")),
    ("--DMS_file_name", dict(default=None, type=str, help="assay CSV when no reference file is used")),
    ("--MSA_filename", dict(default=None, type=str, help="a2m alignment of the wild type (retrieval)")),
    ("--MSA_weight_file_name", dict(default=None, type=str, help="EVE sequence weights (.npy) inside --MSA_weights_folder")),
    ("--MSA_start", dict(default=None, type=int, help="first position covered by the MSA, 1-based")),
    ("--MSA_end", dict(default=None, type=int, help="last position covered by the MSA, 1-based")),
    ("--DMS_data_folder", dict(type=str, help="folder with the assay CSVs")),
    ("--output_scores_folder", dict(default="./", type=str, help="folder for <DMS_id>.csv")),
    ("--deactivate_scoring_mirror", dict(action="store_true", help="score left-to-right only")),
    ("--indel_mode", dict(action="store_true", help="rows are full mutated sequences (insertions / deletions)")),
    ("--scoring_window", dict(default="optimal", type=str, help="optimal | sliding, for sequences longer than n_ctx - 2")),
    ("--num_workers", dict(default=10, type=int, help="accepted for compatibility")),
    ("--inference_time_retrieval", dict(action="store_true", help="fuse the MSA prior into the token log-probabilities")),
    ("--retrieval_inference_weight", dict(default=0.6, type=float, help="alpha of the fusion")),
    ("--MSA_folder", dict(default=".", type=str, help="folder with the MSAs")),
    ("--MSA_weights_folder", dict(default=None, type=str, help="folder with the EVE sequence-weight files")),
    ("--clustal_omega_location", dict(default=None, type=str, help="(indel retrieval only; not supported)")),
]


def create_parser():
    parser = argparse.ArgumentParser(description="Tranception scoring on B200 (score_tranception_proteingym.py drop-in)")
    for flag, kw in _FLAGS:
        parser.add_argument(flag, **kw)
    # additive
    parser.add_argument("--precision", default="f16f8", choices=["f16f8", "f16x3", "f16"])
    parser.add_argument("--device", default=0, type=int)
    parser.add_argument("--MSA_log_prior_npy", default=None, type=str, help="precomputed [L_full, 25] log prior for --inference_time_retrieval")
    return parser


def _init_distributed(args):
    """Under torchrun (WORLD_SIZE > 1) the sequence rows of the assay are split over the GPUs of the box: one process
    per GPU, NCCL, device = LOCAL_RANK; every rank ends with the full score vector, rank 0 writes the CSV. Single process: (0, 1)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    args.device = local
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist.get_rank(), world


def _finish_distributed(world):
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


def main(argv=None):
    from proteingym_b200.tranception_engine import TranceptionScorer, load_tranception_checkpoint
    args = create_parser().parse_args(argv)
    rank, world = _init_distributed(args)
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
            from proteingym_b200.msa_processing import get_msa_prior
            if args.DMS_reference_file_path:
                msa_file = args.MSA_folder + os.sep + mapping["MSA_filename"][args.DMS_index]
                wfile = args.MSA_weights_folder + os.sep + mapping["weight_file_name"][sel].values[0] if args.MSA_weights_folder else None
            else:
                msa_file = args.MSA_folder + os.sep + args.MSA_filename
                wfile = args.MSA_weights_folder + os.sep + args.MSA_weight_file_name if args.MSA_weights_folder is not None else None
            prior = get_msa_prior(msa_file, wfile, MSA_start, MSA_end, len(target_seq), verbose=True, return_depth=False,
                                  device=args.device)  # model_pytorch.py:660-671
            with np.errstate(divide="ignore"):
                log_prior = np.log(prior).astype(np.float32)
    config, state = load_tranception_checkpoint(args.checkpoint)
    scorer = TranceptionScorer(config, state, precision=args.precision, device=args.device)
    if world > 1:
        scorer.shard = (rank, world)
    if rank == 0 and not os.path.isdir(args.output_scores_folder):
        os.mkdir(args.output_scores_folder)
    DMS_data = pd.read_csv(args.DMS_data_folder + os.sep + DMS_file_name, low_memory=False)
    all_scores = scorer.score_mutants(DMS_data=DMS_data, target_seq=target_seq, scoring_mirror=not args.deactivate_scoring_mirror,
                                      indel_mode=args.indel_mode, scoring_window=args.scoring_window, log_prior=log_prior,
                                      retrieval_inference_weight=args.retrieval_inference_weight, MSA_start=MSA_start or 0, MSA_end=MSA_end)
    if rank == 0:
        all_scores.to_csv(args.output_scores_folder + os.sep + DMS_id + ".csv", index=False)
    scorer.close()
    _finish_distributed(world)


if __name__ == '__main__':
    main()
