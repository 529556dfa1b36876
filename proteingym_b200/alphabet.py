"""ESM alphabet and tokenisation (host side).

Mirrors the reference's ``Alphabet.from_architecture("ESM-1b"/"roberta_large")`` and
``BatchConverter.__call__`` for a single sequence
(reference: proteingym/baselines/esm/esm/data.py:142-174 vocabulary, :262-297 batch conversion):
33 tokens, ``<cls> seq <eos>``; ids are int32 on our side (the reference uses int64).
"""
from __future__ import annotations

import numpy as np

# proteinseq_toks["toks"] of the reference (esm/constants.py), in order.
_STANDARD = ["L", "A", "G", "V", "S", "E", "R", "T", "I", "D", "P", "K", "Q", "N", "F", "Y", "M", "H", "W",
             "C", "X", "B", "U", "Z", "O", ".", "-"]


class Alphabet:
    """ESM-1b / ESM-1v / ESM2 alphabet: prepend ``<cls> <pad> <eos> <unk>``, pad to a multiple of 8 with
    ``<null_i>``, append ``<mask>`` (data.py:106-112,151-157)."""

    def __init__(self):
        toks = ["<cls>", "<pad>", "<eos>", "<unk>"] + list(_STANDARD)
        for i in range((8 - (len(toks) % 8)) % 8):
            toks.append(f"<null_{i + 1}>")
        toks.append("<mask>")
        self.all_toks = toks
        self.tok_to_idx = {t: i for i, t in enumerate(toks)}
        self.cls_idx = self.tok_to_idx["<cls>"]
        self.padding_idx = self.tok_to_idx["<pad>"]
        self.eos_idx = self.tok_to_idx["<eos>"]
        self.unk_idx = self.tok_to_idx["<unk>"]
        self.mask_idx = self.tok_to_idx["<mask>"]
        self.prepend_bos = True
        self.append_eos = True

    def __len__(self):
        return len(self.all_toks)

    def get_idx(self, tok: str) -> int:
        # data.py:127-128: unknown tokens map to <unk> (used by label_row for wt/mt letters)
        return self.tok_to_idx.get(tok, self.unk_idx)

    def get_tok(self, ind: int) -> str:
        return self.all_toks[ind]

    def encode(self, seq: str) -> list:
        """Per-character encoding. The reference's ``encode`` (data.py:256-257) indexes ``tok_to_idx`` directly,
        so a character outside the alphabet raises ``KeyError`` there; we keep that behaviour."""
        return [self.tok_to_idx[c] for c in seq]

    def tokenize_sequence(self, seq: str) -> np.ndarray:
        """``[cls] + encode(seq) + [eos]`` as int32 (BatchConverter.__call__, data.py:262-297, batch of one)."""
        ids = [self.cls_idx] + self.encode(seq) + [self.eos_idx]
        return np.asarray(ids, dtype=np.int32)


ALPHABET = Alphabet()
