"""EVE-style alignment pre-processing and the weighted retrieval prior, as the TranceptEVE / Tranception constructors run them.

  * ``MSAProcessing``  mirrors ``MSA_processing`` (trancepteve/utils/msa_utils.py:218-400; the Tranception copy is identical):
    fragment filter, focus columns, indeterminate-residue filter, sequence weights read from a ``.npy`` or computed
    (1 / #{sequences within 1 - theta identity}) — the O(N^2 L) count runs in ``pg_msa_cluster_neighbors`` on the GPU instead of
    the reference's per-sequence numpy dot products. Same attribute names (focus_cols, focus_seq_trimmed, seq_len,
    seq_name_to_sequence, seq_name_to_weight, weights, Neff, num_sequences ...), so code written against the reference class reads it.
  * ``get_msa_prior``  mirrors ``get_msa_prior`` of both baselines (trancepteve/utils/msa_utils.py:63-139 returns
    ``(prior, processed_depth)``; tranception/utils/msa_utils.py:63-138 returns the prior only): hamming filter, weights looked up
    by sequence name (sequences without a weight are dropped), weighted per-column frequencies with the 1e-5 pseudo-count in
    ``pg_msa_prior``.

The arrays are processed as [N, L] byte matrices (no per-character Python loops); nothing here falls back to a CPU weights path —
computing weights needs the CUDA library."""
from __future__ import annotations

import os
import sys
from collections import defaultdict

import numpy as np

ALPHABET = "ACDEFGHIKLMNPQRSTVWY"


def _char_matrix(seqs) -> np.ndarray:
    """Equal-length strings -> [N, L] uint8 (latin-1 code points)."""
    if len(seqs) == 0:
        return np.zeros((0, 0), dtype=np.uint8)
    L = len(seqs[0])
    if any(len(s) != L for s in seqs):
        raise ValueError("alignment rows have different lengths")
    return np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8).reshape(len(seqs), L).copy()


def _upper(a: np.ndarray) -> np.ndarray:
    lower = (a >= ord("a")) & (a <= ord("z"))
    return np.where(lower, a - 32, a).astype(np.uint8)


def _lower(a: np.ndarray) -> np.ndarray:
    upper = (a >= ord("A")) & (a <= ord("Z"))
    return np.where(upper, a + 32, a).astype(np.uint8)


class MSAProcessing:
    def __init__(self, MSA_location="", theta=0.2, use_weights=True, weights_location=None, preprocess_MSA=True,
                 threshold_sequence_frac_gaps=0.5, threshold_focus_cols_frac_gaps=0.3,
                 remove_sequences_with_indeterminate_AA_in_focus_cols=True, device=0, on_missing_weights="exit"):
        self.MSA_location = MSA_location
        self.weights_location = weights_location
        self.theta = theta
        self.alphabet = ALPHABET
        self.use_weights = use_weights
        self.preprocess_MSA = preprocess_MSA
        self.threshold_sequence_frac_gaps = threshold_sequence_frac_gaps
        self.threshold_focus_cols_frac_gaps = threshold_focus_cols_frac_gaps
        self.remove_sequences_with_indeterminate_AA_in_focus_cols = remove_sequences_with_indeterminate_AA_in_focus_cols
        self.device = device
        # a weights path that does not exist: "exit" = the Tranception / TranceptEVE copies stop (msa_utils.py:363-365); "compute" = the
        # copy the ESM scripts use computes the weights and saves them there (proteingym/utils/msa_utils.py:218-241)
        self.on_missing_weights = on_missing_weights
        self.gen_alignment()

    # ------------------------------------------------------------------------------------------------------------------
    def gen_alignment(self):
        self.aa_dict = {aa: i for i, aa in enumerate(self.alphabet)}
        raw = defaultdict(str)
        name = ""
        self.focus_seq_name = None
        with open(self.MSA_location, "r") as fh:
            for i, line in enumerate(fh):
                line = line.rstrip()
                if line.startswith(">"):
                    name = line
                    if i == 0:
                        self.focus_seq_name = name
                else:
                    raw[name] += line
        names = list(raw.keys())
        mat = _char_matrix([raw[n] for n in names])

        if self.preprocess_MSA:  # msa_utils.py:283-310
            assert 0.0 <= self.threshold_sequence_frac_gaps <= 1.0, "Invalid fragment filtering parameter"
            assert 0.0 <= self.threshold_focus_cols_frac_gaps <= 1.0, "Invalid focus position filtering parameter"
            mat = _upper(np.where(mat == ord("."), ord("-"), mat).astype(np.uint8))
            focus_row = names.index(self.focus_seq_name)
            mat = mat[:, mat[focus_row] != ord("-")]               # columns that are gaps in the wild type go
            gaps = mat == ord("-")
            seq_ok = gaps.mean(axis=1) <= self.threshold_sequence_frac_gaps
            col_ok = gaps[seq_ok].mean(axis=0) <= self.threshold_focus_cols_frac_gaps
            mat = np.where(col_ok[None, :], mat, _lower(mat))      # non-focus columns in lower case
            mat = mat[seq_ok]
            names = [n for n, k in zip(names, seq_ok) if k]
        self._names_all = names
        if self.focus_seq_name in names:
            frow = mat[names.index(self.focus_seq_name)]
            self.focus_seq = frow.tobytes().decode("latin-1")
        else:                                                      # the reference's defaultdict yields '' for a filtered-out focus
            frow = np.zeros(0, dtype=np.uint8)
            self.focus_seq = ""
        is_focus = (frow == _upper(frow)) & (frow != ord("-"))
        self.focus_cols = [int(i) for i in np.nonzero(is_focus)[0]]
        fset = set(self.focus_cols)
        self.non_focus_cols = [ix for ix in range(len(self.focus_seq)) if ix not in fset]
        self.focus_seq_trimmed = [self.focus_seq[ix] for ix in self.focus_cols]
        self.seq_len = len(self.focus_cols)
        self.alphabet_size = len(self.alphabet)
        try:  # uniprot numbering from ">name/start-stop" (msa_utils.py:320-333)
            start, stop = self.focus_seq_name.split("/")[-1].split("-")
            self.focus_start_loc, self.focus_stop_loc = int(start), int(stop)
        except Exception:
            self.focus_start_loc, self.focus_stop_loc = 1, len(self.focus_seq)
        start = self.focus_start_loc
        self.uniprot_focus_col_to_wt_aa_dict = {c + start: self.focus_seq[c] for c in self.focus_cols}
        self.uniprot_focus_col_to_focus_idx = {c + start: c for c in self.focus_cols}
        self.raw_seq_name_to_sequence = {n: mat[i].tobytes().decode("latin-1") for i, n in enumerate(names)}

        fmat = mat[:, self.focus_cols] if mat.size else np.zeros((len(names), 0), dtype=np.uint8)
        fmat = _upper(np.where(fmat == ord("."), ord("-"), fmat).astype(np.uint8))
        lut = np.full(256, 255, dtype=np.uint8)                    # 0 = gap, 1..20 = residue, 255 = indeterminate
        lut[ord("-")] = 0
        for aa, k in self.aa_dict.items():
            lut[ord(aa)] = k + 1
        tok = lut[fmat]
        if self.remove_sequences_with_indeterminate_AA_in_focus_cols:
            good = ~(tok == 255).any(axis=1)
            tok, fmat = tok[good], fmat[good]
            names = [n for n, k in zip(names, good) if k]
        tok = np.where(tok == 255, 0, tok).astype(np.uint8)        # kept indeterminate letters encode as empty one-hot rows
        self.tokens = tok
        self.seq_name_to_sequence = {n: list(fmat[i].tobytes().decode("latin-1")) for i, n in enumerate(names)}

        if self.use_weights:
            if (self.weights_location is not None) and (not os.path.isfile(self.weights_location)) and self.on_missing_weights == "exit":
                print("Provided weights location is invalid")
                sys.exit(0)                                        # the reference's behaviour (msa_utils.py:363-365)
            from . import sharding
            rank, world = sharding.rank_world()

            def load_or_compute():
                try:
                    self.weights = np.load(file=self.weights_location)
                except Exception:
                    self.weights = self.compute_weights()
                    if self.weights_location is not None:  # atomic: another rank / process may be reading this path
                        sharding.atomic_write(self.weights_location, lambda tmp: np.save(file=open(tmp, "wb"), arr=self.weights))
            if world > 1 and self.weights_location is not None:
                if rank == 0:
                    load_or_compute()
                sharding.barrier_if_distributed()  # ranks > 0 read what rank 0 found or wrote
                if rank != 0:
                    load_or_compute()
            else:
                load_or_compute()
        else:
            self.weights = np.ones(tok.shape[0])
        self.Neff = np.sum(self.weights)
        self.num_sequences = tok.shape[0]
        self.seq_name_to_weight = {}
        for i, n in enumerate(names):
            self.seq_name_to_weight[n] = self.weights[i]           # by position: IndexError on a too-short file, as the reference

    def compute_weights(self) -> np.ndarray:
        """1 / #{t : <onehot(t), onehot(s)> / <onehot(s), onehot(s)> > 1 - theta}; 0 for an empty sequence (msa_utils.py:371-381)."""
        from .msa_prior import cluster_weights
        if self.tokens.shape[0] == 0:
            return np.zeros(0)
        return cluster_weights(self.tokens, 1 - self.theta, empty_value=0, device=self.device)

    @property
    def one_hot_encoding(self) -> np.ndarray:
        oh = np.zeros((self.tokens.shape[0], self.tokens.shape[1], len(self.alphabet)))
        n, l = np.nonzero(self.tokens)
        oh[n, l, self.tokens[n, l] - 1] = 1.0
        return oh


def get_msa_prior(MSA_data_file, MSA_weight_file_name, MSA_start, MSA_end, len_target_seq, vocab=None,
                  retrieval_aggregation_mode="aggregate_substitution", filter_MSA=True, verbose=False,
                  threshold_sequence_frac_gaps=None, threshold_focus_cols_frac_gaps=None, return_depth=True, device=0):
    """``(msa_prior [len_target_seq, 25] float64, processed_MSA_depth)`` — trancepteve/utils/msa_utils.py:63-139. With
    ``return_depth=False`` and the threshold arguments left None this is the Tranception signature (MSA_processing defaults
    0.5 / 0.3). ``vocab`` is accepted for signature compatibility; the tokenizer vocabulary is fixed (Basic_tokenizer)."""
    from . import msa_prior as mp
    msa = mp.read_a2m(MSA_data_file)
    if verbose:
        print("Target seq len is {}, MSA length is {}, start position is {}, end position is {} and vocab size is {}".format(
            len_target_seq, MSA_end - MSA_start, MSA_start, MSA_end, mp.VOCAB_SIZE))
    weights = None
    if MSA_weight_file_name is not None:
        kw = {}
        if return_depth:  # the TranceptEVE copy passes the fragment threshold through and always keeps every column (:96-101)
            kw = dict(threshold_sequence_frac_gaps=threshold_sequence_frac_gaps, threshold_focus_cols_frac_gaps=1.0)
        else:
            assert os.path.exists(MSA_weight_file_name), "Weights file not located on disk."
        eve = MSAProcessing(MSA_location=MSA_data_file, use_weights=True, weights_location=MSA_weight_file_name, device=device, **kw)
        weights = {n: eve.seq_name_to_weight[n] for n in eve.seq_name_to_sequence}
    if retrieval_aggregation_mode in ("aggregate_substitution", "aggregate_indel"):
        prior, depth = mp.msa_prior(msa, MSA_start, MSA_end, len_target_seq, weights, filter_MSA=filter_MSA, device=device, return_depth=True)
    else:
        depth = mp.filtered_depth(msa, weights, filter_MSA)
        prior = np.ones((len_target_seq, mp.VOCAB_SIZE)) / mp.VOCAB_SIZE
    return (prior, depth) if return_depth else prior
