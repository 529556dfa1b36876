"""Model-object seam: what ``proteingym/baselines/esm/compute_fitness.py`` touches on the objects returned by
``pretrained.load_model_and_alphabet`` (esm/pretrained.py:24-28), so that the reference's own Python loop (:433-529) can drive the
B200 forward one call at a time:

    model, alphabet = load_model_and_alphabet(path); model.eval(); model.cuda()
    batch_converter = alphabet.get_batch_converter(); _, _, batch_tokens = batch_converter([("protein1", sequence)])
    token_probs = torch.log_softmax(model(batch_tokens_masked.cuda())["logits"], dim=-1)

``model(tokens)["logits"]`` holds the per-token LOG-PROBABILITIES (the library applies the log-softmax itself); the reference's
``torch.log_softmax`` maps them to themselves, so its loop needs no change. This seam is for parity debugging — batch-1 calls cannot
reach the throughput of ``EsmScorer.score_assay`` (one batched pass) — and supports what that loop sends: at most ONE ``<mask>``
per sequence (the token-dropout rescale depends on the mask count; the C-ABI takes a single masked position)."""
from __future__ import annotations

import numpy as np
import torch

from .alphabet import ALPHABET, Alphabet
from .checkpoint import load_esm_checkpoint


class BatchConverter:
    """``alphabet.get_batch_converter()`` (esm/data.py:262-297) for the ESM-1b/1v/ESM2 alphabets: [(label, sequence)] ->
    (labels, sequences, int64 tokens [B, max_len + 2] padded with ``<pad>``)."""

    def __init__(self, alphabet: Alphabet):
        self.alphabet = alphabet

    def __call__(self, raw_batch):
        labels, strs = zip(*raw_batch) if len(raw_batch) else ((), ())
        enc = [self.alphabet.tokenize_sequence(s) for s in strs]
        width = max((len(e) for e in enc), default=0)
        tokens = torch.full((len(enc), width), self.alphabet.padding_idx, dtype=torch.int64)
        for i, e in enumerate(enc):
            tokens[i, :len(e)] = torch.from_numpy(e.astype(np.int64))
        return list(labels), list(strs), tokens


def _get_batch_converter(self):
    return BatchConverter(self)


Alphabet.get_batch_converter = _get_batch_converter


def mask_position(row: torch.Tensor, mask_idx: int) -> int:
    """Index of the single ``<mask>`` in a token row, -1 if there is none; more than one is outside this seam."""
    pos = torch.nonzero(row == mask_idx).flatten()
    if pos.numel() > 1:
        raise NotImplementedError("more than one <mask> per sequence: use EsmScorer (the batched API) instead of the model-object seam")
    return int(pos[0]) if pos.numel() else -1


class B200EsmModel:
    def __init__(self, config, state, name, precision="f16x3", device=0):
        self.config, self._state, self.name = config, state, name
        self.precision, self._device = precision, device
        self.scorer = None

    def eval(self):
        return self

    def cuda(self, device=None):
        from .esm_engine import EsmScorer
        if self.scorer is None:
            if device is not None:
                self._device = device if isinstance(device, int) else torch.device(device).index or 0
            self.scorer = EsmScorer(self.config, self._state, precision=self.precision, device=self._device)
            self._state = None
        return self

    def __call__(self, tokens: torch.Tensor, **kwargs):
        if self.scorer is None:
            raise RuntimeError("B200EsmModel: call .cuda() first (there is no CPU forward)")
        if tokens.dim() != 2:
            raise ValueError("tokens must be [B, T]")
        if (tokens == ALPHABET.padding_idx).any():
            raise NotImplementedError("padded batches are outside this seam (the scoring loop sends one sequence at a time)")
        B, T = tokens.shape
        out = torch.empty((B, T, self.config.vocab), dtype=torch.float32, device=self.scorer.device)
        for b in range(B):
            row = tokens[b]
            out[b] = self.scorer._forward_window(row.to(self.scorer.device, torch.int32).contiguous(), 0, T, mask_position(row, ALPHABET.mask_idx))
        return {"logits": out}


def load_model_and_alphabet(model_location: str, precision: str = "f16x3", device: int = 0):
    config, state, name = load_esm_checkpoint(model_location)
    return B200EsmModel(config, state, name, precision, device), ALPHABET
