"""MSA pre-processing for Tranception's inference-time retrieval, on the GPU.

  * ``msa_log_prior``     -> the ``MSA_log_prior`` [L_full, 25] tensor the reference builds at model init
    (tranception/model_pytorch.py:660-671 calling utils/msa_utils.py:63-138 ``get_msa_prior``): a2m parsing and the
    hamming filter on the host, the weighted per-column frequency reduction in ``pg_msa_prior``.
  * ``cluster_weights``   -> EVE-style sequence weights 1/|cluster| (proteingym/utils/weights.py:13-53 ``calc_weights_fast``),
    the O(N^2 L) pairwise identity count in ``pg_msa_cluster_neighbors`` (numba on CPU in the reference).

``MSA_processing`` (focus-column / fragment filtering, the name->weight map read from or written to a ``.npy``) and the
reference-signature ``get_msa_prior`` live in msa_processing.py; here weights arrive as a ``{sequence name: weight}`` dict."""
from __future__ import annotations

from collections import defaultdict

import numpy as np
import torch

from . import _lib
from .tranception_engine import TOK

VOCAB_SIZE = 25


def read_a2m(path: str):
    """process_msa_data (msa_utils.py:28-40): name line -> concatenated upper-cased sequence, insertion order kept."""
    msa = defaultdict(str)
    name = ""
    with open(path) as fh:
        for line in fh:
            line = line.rstrip()
            if line.startswith(">"):
                name = line
            else:
                msa[name] += line.upper()
    return msa


def encode(seqs, unknown=255) -> np.ndarray:
    """[N, L] uint8 token ids of equal-length aligned strings; letters outside the tokenizer vocabulary -> ``unknown``."""
    lut = np.full(256, unknown, dtype=np.uint8)
    for ch, i in TOK.items():
        if len(ch) == 1:
            lut[ord(ch)] = i
    arr = np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8).reshape(len(seqs), -1)
    return lut[arr]


def _select(msa: dict, weights: dict | None, filter_MSA: bool):
    """Sequences get_msa_prior keeps (hamming filter, then those that have a weight) -> (names, tokens [N, L], weights [N])."""
    names = list(msa.keys())
    tok = encode([msa[n] for n in names])
    if filter_MSA:  # hamming similarity to the first sequence over its in-vocabulary positions (msa_utils.py:84-92)
        ref = tok[0]
        valid = ref != 255
        sim = ((tok == ref) & valid).sum(axis=1) / valid.sum()
        keep = ~(sim < 0.2)
        names = [n for n, k in zip(names, keep) if k]
        tok = tok[keep]
    if weights is not None:  # sequences without a weight are dropped (msa_utils.py:103-110)
        keep = np.array([n in weights for n in names], dtype=bool)
        tok = tok[keep]
        names = [n for n, k in zip(names, keep) if k]
        w = np.array([weights[n] for n in names], dtype=np.float64)
    else:
        w = np.ones(len(names), dtype=np.float64)
    return names, tok, w


def filtered_depth(msa: dict, weights: dict | None = None, filter_MSA: bool = True) -> int:
    """``processed_MSA_depth`` of the TranceptEVE get_msa_prior (trancepteve/utils/msa_utils.py:117)."""
    return len(_select(msa, weights, filter_MSA)[0])


def msa_prior(msa: dict, MSA_start: int, MSA_end: int, len_target_seq: int, weights: dict | None = None, filter_MSA: bool = True,
              device: int = 0, return_depth: bool = False):
    """get_msa_prior(..., retrieval_aggregation_mode="aggregate_substitution") -> float64 [len_target_seq, 25]
    (and the number of sequences that went into it when ``return_depth``)."""
    names, tok, w = _select(msa, weights, filter_MSA)
    Lm = MSA_end - MSA_start
    cols = np.full((len(names), Lm), 255, dtype=np.uint8)  # one_hots has MSA_end-MSA_start columns; sequences fill from column 0
    n = min(Lm, tok.shape[1])
    cols[:, :n] = tok[:, :n]
    if tok.shape[1] > Lm:
        raise IndexError("MSA sequences are longer than MSA_end - MSA_start")  # the reference's one_hots[i, j, k] indexing fails too
    lib = _lib.load()
    dev = torch.device("cuda", device)
    d_tok = torch.from_numpy(np.ascontiguousarray(cols.T)).to(dev)
    d_w = torch.from_numpy(w).to(dev)
    out = torch.empty((Lm, VOCAB_SIZE), dtype=torch.float64, device=dev)
    _lib.check(lib.pg_msa_prior(d_tok.data_ptr(), d_w.data_ptr(), len(names), Lm, VOCAB_SIZE, 1e-5, out.data_ptr(),
                                torch.cuda.current_stream(dev).cuda_stream))
    prior = np.zeros((len_target_seq, VOCAB_SIZE))
    prior[MSA_start:MSA_end, :] = out.cpu().numpy()
    return (prior, len(names)) if return_depth else prior


def msa_log_prior(msa_file: str, MSA_start: int, MSA_end: int, len_target_seq: int, weights: dict | None = None, device: int = 0):
    """float32 log prior as stored in ``TranceptionLMHeadModel.MSA_log_prior`` (model_pytorch.py:660-671)."""
    p = msa_prior(read_a2m(msa_file), MSA_start, MSA_end, len_target_seq, weights, device=device)
    with np.errstate(divide="ignore"):
        return torch.log(torch.tensor(p)).float().numpy()


def cluster_weights(matrix_mapped: np.ndarray, identity_threshold: float, empty_value: int = 0, device: int = 0) -> np.ndarray:
    """calc_weights_fast (weights.py:13-53): ``matrix_mapped`` [N, L] small ints (``empty_value`` = gap). Sequences that are all
    gaps get weight 0; the others 1 / #{sequences within the identity threshold, counted over their own non-gap length}."""
    m = np.asarray(matrix_mapped)
    assert m.ndim == 2, f"Matrix must be 2D; shape={m.shape}"
    N = m.shape[0]
    tok = np.where(m == empty_value, 0, m.astype(np.int64) + (1 if empty_value != 0 else 0)).astype(np.uint8)
    empty = np.all(tok == 0, axis=1)
    act = np.ascontiguousarray(tok[~empty])
    weights = np.zeros(N)
    if act.shape[0] == 0:
        return weights
    L = float(m.shape[1])
    Lng = L - (act == 0).sum(axis=1)
    # smallest integer match count m with (m / L_non_gaps > identity_threshold) under the reference's float64 arithmetic
    guess = np.floor(identity_threshold * Lng).astype(np.int64)
    need = np.empty(len(Lng), dtype=np.int32)
    for k in range(len(Lng)):
        c = max(int(guess[k]) - 1, 0)
        while not (c / Lng[k] > identity_threshold):
            c += 1
        need[k] = c
    lib = _lib.load()
    dev = torch.device("cuda", device)
    d_tok = torch.from_numpy(act).to(dev)
    d_need = torch.from_numpy(need).to(dev)
    d_out = torch.empty(act.shape[0], dtype=torch.int32, device=dev)
    _lib.check(lib.pg_msa_cluster_neighbors(d_tok.data_ptr(), act.shape[1], act.shape[0], act.shape[1], d_need.data_ptr(), d_out.data_ptr(),
                                            torch.cuda.current_stream(dev).cuda_stream))
    weights[~empty] = 1.0 / d_out.cpu().numpy().astype(np.float64)
    return weights
