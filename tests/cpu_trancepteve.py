"""CPU stand-in for the device half of TranceptEVEScorer, for host-logic tests: the retrieval logic of
proteingym_b200.trancepteve_engine (windows, prior-row arithmetic, recalibration, score assembly) runs unchanged while
``sequence_logprobs`` is answered by the oracle's CPU forward applying the fusion semantics of include/pgscore.h. Test code only."""
from __future__ import annotations

import numpy as np
import torch

from oracle import tranception_oracle as O
from proteingym_b200.trancepteve_engine import TranceptEVEScorer, retrieval_weights
from proteingym_b200.tranception_engine import prior_rows


class CpuTranceptEVE(TranceptEVEScorer):
    def __init__(self, arch, state, full_target_seq, kind, MSA_log_prior, EVE_log_prior, MSA_start, MSA_end, depths, focus_cols,
                 col_thr, msa_recal, eve_recal, scoring_window="optimal"):
        self.st, self.arch = state, arch
        self.n_ctx, self.vocab = arch.n_ctx, arch.vocab
        self.full_target_seq, self.full_protein_length = full_target_seq, len(full_target_seq)
        self.scoring_window = scoring_window
        self.inference_time_retrieval_type = kind
        self.retrieval_aggregation_mode = "aggregate_substitution"
        self.MSA_recalibrate_probas, self.EVE_recalibrate_probas = msa_recal, eve_recal
        self.MSA_threshold_focus_cols_frac_gaps = col_thr
        self.MSA_log_prior = torch.tensor(MSA_log_prior).clone()
        self.EVE_log_prior = torch.tensor(EVE_log_prior).clone() if EVE_log_prior is not None else None
        self.MSA_start, self.MSA_end = MSA_start, MSA_end
        self.MSA_processed_depth, self.EVE_processed_depth = depths
        self.EVE_MSA = type("M", (), {"focus_cols": focus_cols})()
        self.retrieval_inference_MSA_weight, self.retrieval_inference_EVE_weight = retrieval_weights(kind, "aggregate_substitution", *depths)

    def close(self):
        pass

    def sequence_logprobs(self, seqs, prior=None, windows=None, flip=False, alpha=0.6, msa_start=0, msa_end=None, chunk_rows=0,
                          prior2=None, beta=0.0, first_col=0, nonfocus_fallback=False, return_rows=False):
        out = np.zeros(len(seqs), dtype=np.float32)
        rows_out = []
        nonfocus = None
        if prior2 is not None and nonfocus_fallback and beta > 0:
            nonfocus = np.asarray(prior2)[:, 5:].min(axis=1) == -np.inf
        cache = {}
        for k, s in enumerate(seqs):
            T = len(s) + 2
            prow = prow2 = None
            if prior is not None:
                prow = np.full(T, -1, dtype=np.int32)
                prow2 = np.full(T, -1, dtype=np.int32) if prior2 is not None else None
                prior_rows(prow, prow2, windows[k][0], windows[k][1], msa_start, prior.shape[0] if msa_end is None else msa_end, flip, nonfocus)
            key = (s, None if prow is None else prow.tobytes(), None if prow2 is None else prow2.tobytes())
            if key not in cache:
                cache[key] = O.fused_logprob_rows(self.st, s, self.arch.layers, self.arch.heads, self.arch.ln_eps, prior, prow, alpha, prior2,
                                                  prow2, beta, first_col)
            rows, total = cache[key]
            out[k] = total
            rows_out.append(rows)
        return (out, rows_out) if return_rows else out
