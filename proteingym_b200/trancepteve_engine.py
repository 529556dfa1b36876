"""TranceptEVE on the B200 scorer: ``TrancepteveLMHeadModel`` (proteingym/baselines/trancepteve/trancepteve/model_pytorch.py:659-1260)
minus the transformer, which is the Tranception stack behind ``pg_ar_loglik_fused``.

What lives here is the per-assay retrieval logic around the forward:
  * constructor (:659-765): MSA log prior with EVE-style sequence weights (msa_processing.get_msa_prior), EVE log prior
    (eve_prior.eve_log_prior, cached where the reference caches it), aggregation weights alpha / beta from the processed depths;
  * ``get_transformer_log_softmax`` (:821-874), ``recalibrate_MSA_probas`` / ``recalibrate_EVE_probas`` (:876-905);
  * ``score_mutants`` (:1180-1235): both directions, WT delta per window, ``mutant`` column merged back.
The three-way fusion itself ((1-beta)*((1-alpha)*logp + alpha*MSA) + beta*EVE on the amino-acid columns, :1100-1133) runs inside the
head kernel; the host only does the slice / flip index arithmetic (tranception_engine.prior_rows).

Not reproduced: ``aggregate_indel`` retrieval (re-aligns every mutated sequence with Clustal Omega, utils/msa_utils.py:142-216) —
indels are scored without retrieval, like the reference does when no MSA is given."""
from __future__ import annotations

import bisect

import numpy as np
import pandas as pd
import torch

from . import eve_prior
from .msa_processing import MSAProcessing, get_msa_prior
from .tranception_engine import TranceptionScorer, apply_substitutions


def retrieval_weights(inference_time_retrieval_type, retrieval_aggregation_mode, MSA_processed_depth, EVE_processed_depth,
                      retrieval_weights_manual=False, retrieval_inference_MSA_weight=0.5, retrieval_inference_EVE_weight=0.5):
    """(alpha, beta) as the constructor picks them (:720-763)."""
    if retrieval_weights_manual:
        return retrieval_inference_MSA_weight, retrieval_inference_EVE_weight
    if inference_time_retrieval_type == "Tranception":
        return 0.6, 0.0
    if inference_time_retrieval_type != "TranceptEVE":
        raise ValueError("inference_time_retrieval_type not recognized")
    if retrieval_aggregation_mode == "aggregate_indel":
        return (0.0, 0.0) if MSA_processed_depth < 10 else (0.5, 0.1)

    rung = bisect.bisect_right([10, 10 ** 2, 10 ** 3, 10 ** 4, 10 ** 5], MSA_processed_depth)
    alpha = [0.0, 0.1, 0.3, 0.4, 0.4, 0.5][rung]
    rung = bisect.bisect_right([10, 10 ** 2, 10 ** 3, 10 ** 4, 10 ** 5], EVE_processed_depth)
    beta = [0.0, 0.3, 0.6, 0.7, 0.7, 0.8][rung]
    return alpha, beta


def iterative_recalibrations(x: torch.Tensor, target, distance_stop_criterion=0.001, max_steps=1000) -> torch.Tensor:
    """Temperature iteration of :876-887: rescale the log-probabilities until their mean matches ``target``."""
    loss = abs(x.mean() - target)
    step = 0
    while loss > distance_stop_criterion:
        T = x.mean() / target
        x = torch.log_softmax(x / T, dim=-1)
        loss = abs(x.mean() - target)
        step += 1
        if step > max_steps:
            break
    return x


class TranceptEVEScorer(TranceptionScorer):
    def __init__(self, config: dict, state: dict, full_target_seq: str, inference_time_retrieval_type=None,
                 retrieval_aggregation_mode=None, MSA_filename=None, MSA_weight_file_name=None, MSA_start=None, MSA_end=None,
                 MSA_threshold_sequence_frac_gaps=None, MSA_threshold_focus_cols_frac_gaps=None, EVE_model_paths=None,
                 EVE_num_samples_log_proba=10, EVE_model_parameters_location=None, MSA_recalibrate_probas=False,
                 EVE_recalibrate_probas=True, retrieval_weights_manual=False, retrieval_inference_MSA_weight=0.5,
                 retrieval_inference_EVE_weight=0.5, scoring_window="optimal", precision="f16f8", device=0, max_rows=0,
                 EVE_sampler="auto"):
        super().__init__(config, state, precision=precision, device=device, max_rows=max_rows)
        self.full_target_seq = full_target_seq
        self.full_protein_length = len(full_target_seq)
        self.scoring_window = scoring_window
        self.inference_time_retrieval_type = inference_time_retrieval_type
        self.retrieval_aggregation_mode = retrieval_aggregation_mode
        self.MSA_recalibrate_probas = MSA_recalibrate_probas
        self.EVE_recalibrate_probas = EVE_recalibrate_probas
        self.MSA_threshold_focus_cols_frac_gaps = MSA_threshold_focus_cols_frac_gaps
        self.MSA_log_prior = self.EVE_log_prior = None
        self.MSA_processed_depth = self.EVE_processed_depth = 0
        self.retrieval_inference_MSA_weight = self.retrieval_inference_EVE_weight = 0.0
        if inference_time_retrieval_type is None:
            print("Model only uses autoregressive inference")
            return
        if retrieval_aggregation_mode == "aggregate_indel":
            raise NotImplementedError("retrieval for indels re-aligns each sequence with Clustal Omega; score indels without retrieval")
        print("Model leverages both autoregressive and retrieval inference (Type: {})".format(inference_time_retrieval_type))
        self.MSA_filename, self.MSA_start, self.MSA_end = MSA_filename, MSA_start, MSA_end
        dev = self.device.index
        if inference_time_retrieval_type.startswith("Trancept"):
            prior, self.MSA_processed_depth = get_msa_prior(
                MSA_data_file=MSA_filename, MSA_weight_file_name=MSA_weight_file_name, MSA_start=MSA_start, MSA_end=MSA_end,
                len_target_seq=self.full_protein_length, retrieval_aggregation_mode=retrieval_aggregation_mode, filter_MSA=True,
                threshold_sequence_frac_gaps=MSA_threshold_sequence_frac_gaps,
                threshold_focus_cols_frac_gaps=MSA_threshold_focus_cols_frac_gaps, verbose=True, device=dev)
            with np.errstate(divide="ignore"):
                self.MSA_log_prior = torch.log(torch.tensor(prior).float())
        if inference_time_retrieval_type == "TranceptEVE":
            assert (EVE_model_paths is not None) and len(EVE_model_paths) >= 1, "Could not find a reference for EVE model"
            if MSA_threshold_focus_cols_frac_gaps != 1.0:
                print("threshold_focus_cols_frac_gaps not 1.0. Only well-covered positions are factored in the EVE retrieval aggregation.")
            self.EVE_MSA = MSAProcessing(MSA_location=MSA_filename, use_weights=True, weights_location=MSA_weight_file_name,
                                         threshold_sequence_frac_gaps=MSA_threshold_sequence_frac_gaps,
                                         threshold_focus_cols_frac_gaps=MSA_threshold_focus_cols_frac_gaps, device=dev)
            self.EVE_log_prior = eve_prior.eve_log_prior(EVE_model_paths, EVE_model_parameters_location, self.EVE_MSA,
                                                         self.full_protein_length, MSA_start, EVE_num_samples_log_proba,
                                                         device=self.device, sampler=EVE_sampler).cpu()
            self.EVE_processed_depth = len(self.EVE_MSA.seq_name_to_sequence.keys())
        self.retrieval_inference_MSA_weight, self.retrieval_inference_EVE_weight = retrieval_weights(
            inference_time_retrieval_type, retrieval_aggregation_mode, self.MSA_processed_depth, self.EVE_processed_depth,
            retrieval_weights_manual, retrieval_inference_MSA_weight, retrieval_inference_EVE_weight)
        if not retrieval_weights_manual and inference_time_retrieval_type == "TranceptEVE":
            print("Aggregation weights of retrieved MSA & EVE model are based on processed MSA depth: MSA({}) and EVE({})".format(
                self.retrieval_inference_MSA_weight, self.retrieval_inference_EVE_weight))

    # ------------------------------------------------------------------------------------------------------------------
    def fusion_kwargs(self, inference_time_retrieval_type=None):
        """Arguments of ``sequence_logprobs`` for one retrieval type (forward's branches, :1100-1133)."""
        kind = inference_time_retrieval_type if inference_time_retrieval_type is not None else self.inference_time_retrieval_type
        if kind is None or self.MSA_log_prior is None:
            return {}
        kw = dict(prior=self.MSA_log_prior.numpy(), alpha=self.retrieval_inference_MSA_weight, msa_start=self.MSA_start,
                  msa_end=self.MSA_end, first_col=5)
        if kind == "TranceptEVE":
            kw.update(prior2=self.EVE_log_prior.numpy(), beta=self.retrieval_inference_EVE_weight,
                      nonfocus_fallback=self.MSA_threshold_focus_cols_frac_gaps < 1.0)
        elif kind != "Tranception":
            raise ValueError("inference_time_retrieval_type not recognized")
        return kw

    def get_transformer_log_softmax(self, sequence, inference_time_retrieval_type="Tranception"):
        """(fused log-softmax [len + 1, vocab], shifted labels) of a full sequence cut into consecutive context windows (:821-874).
        As in the reference no ``flip`` is passed, so a reversed ``sequence`` still meets the priors in forward order."""
        ctx = self.n_ctx - 2
        num_windows = 1 + int(len(sequence) / ctx)
        pieces = [sequence[w * ctx:(w + 1) * ctx] for w in range(num_windows)]
        windows = [(w * ctx, min(len(sequence), (w + 1) * ctx)) for w in range(num_windows)]
        _, rows = self.sequence_logprobs(pieces, windows=windows, flip=False, return_rows=True,
                                         **self.fusion_kwargs(inference_time_retrieval_type))
        out = np.zeros((len(sequence) + 1, self.vocab), dtype=np.float32)
        if num_windows > 1:
            start = 0
            for w in range(num_windows):
                if w < num_windows - 1:
                    out[start:start + ctx] = rows[w][:ctx]
                else:
                    out[start:] = rows[w][:len(sequence) + 1 - start]
                start += ctx
        else:
            out[:] = rows[0][:len(sequence) + 1]
        from .tranception_engine import tokenize, replace_ambiguous
        labels = np.asarray(tokenize(replace_ambiguous(sequence))[1:], dtype=np.int64)
        return torch.from_numpy(out), labels

    def recalibrate_MSA_probas(self):
        # the reference passes inference_time_retrieval_type=None here, which forward() resolves to the model's own type (:1048)
        lr, _ = self.get_transformer_log_softmax(self.full_target_seq, inference_time_retrieval_type=None)
        rl, _ = self.get_transformer_log_softmax(self.full_target_seq[::-1], inference_time_retrieval_type=None)
        s, e = self.MSA_start, self.MSA_end
        target = (lr[s:e, 5:].mean() + rl[s:e, 5:].mean()) / 2.0
        print("Optimal temperature for MSA proba recalibration: {}".format(self.MSA_log_prior[s:e, 5:].mean() / target))
        self.MSA_log_prior[s:e, 5:] = iterative_recalibrations(self.MSA_log_prior[s:e, 5:], target)

    def recalibrate_EVE_probas(self):
        lr, _ = self.get_transformer_log_softmax(self.full_target_seq)
        rl, _ = self.get_transformer_log_softmax(self.full_target_seq[::-1])
        rows = [self.MSA_start + c for c in self.EVE_MSA.focus_cols]
        target = (lr[rows, 5:].mean() + rl[rows, 5:].mean()) / 2.0
        print("Optimal temperature for EVE proba recalibration: {}".format(self.EVE_log_prior[rows, 5:].mean() / target))
        self.EVE_log_prior[rows, 5:] = iterative_recalibrations(self.EVE_log_prior[rows, 5:], target)

    # ------------------------------------------------------------------------------------------------------------------
    def score_mutants(self, DMS_data: pd.DataFrame, target_seq=None, scoring_mirror=True, batch_size_inference=10, num_workers=10,
                      indel_mode=False) -> pd.DataFrame:
        """Same signature as the reference (:1180); ``batch_size_inference`` / ``num_workers`` are accepted and ignored (batches are
        sized by the library's workspace, tokenisation is vectorised on the host)."""
        df = DMS_data.copy()
        if self.MSA_recalibrate_probas and self.MSA_log_prior is not None:
            self.recalibrate_MSA_probas()
        if self.EVE_recalibrate_probas and self.EVE_log_prior is not None:
            self.recalibrate_EVE_probas()
        if ("mutated_sequence" not in df) and (not indel_mode):
            df["mutated_sequence"] = df["mutant"].apply(lambda m: apply_substitutions(target_seq, m))
        assert "mutated_sequence" in df, "DMS file to score does not have mutated_sequence column"
        if "mutant" not in df:
            df["mutant"] = df["mutated_sequence"]
        df = df[["mutated_sequence", "mutant"]].reset_index(drop=True)
        if target_seq is None:
            raise ValueError("target_seq is required (scores are deltas against the wild type scored in the same window)")
        window = self.scoring_window
        sl = self.slices(df, target_seq, window, indel_mode)
        pk = self.fusion_kwargs()
        print("Scoring sequences from left to right")
        out = self._direction(sl, target_seq, "avg_score_L_to_R", window, False, pk)
        if scoring_mirror:
            print("Scoring sequences from right to left")
            rl = self._direction(sl, target_seq, "avg_score_R_to_L", window, True, pk)
            out = pd.merge(out, rl, on="mutated_sequence", how="left", suffixes=("", "_R_to_L"))
            out["avg_score"] = (out["avg_score_L_to_R"] + out["avg_score_R_to_L"]) / 2.0
        else:
            out["avg_score"] = out["avg_score_L_to_R"]
        if target_seq in df.mutated_sequence.values:  # the WT scores 0 by definition (:1224-1229)
            cols = ["mutated_sequence", "avg_score_L_to_R", "avg_score_R_to_L", "avg_score"] if scoring_mirror \
                else ["mutated_sequence", "avg_score_L_to_R", "avg_score"]
            out = pd.concat([out, pd.DataFrame([[target_seq] + [0] * (len(cols) - 1)], columns=cols)], ignore_index=True)
        if len(out) > 0 and not indel_mode:
            out = pd.merge(out, df, how="left", on="mutated_sequence")
        return out
