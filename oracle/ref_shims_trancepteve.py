"""ORACLE SUPPORT — TEST INFRASTRUCTURE ONLY (build container only: needs /root/reference).

Makes the reference's ``trancepteve.model_pytorch.TrancepteveLMHeadModel`` constructible under transformers 5.x (the reference
pins 4.32.1): on top of the import shims of ref_shims_tranception, ``PreTrainedModel.__init__`` is reduced to ``nn.Module.__init__``
while the model is built, ``init_weights`` is a no-op (weights come from a state dict anyway) and ``get_head_mask`` (dropped from
transformers 5) returns the all-None mask it always returned on this path. Everything else — the constructor's retrieval set-up
(get_msa_prior, MSA_processing, EVE VAE, depth-based weights), forward with the prior fusion, recalibration, score_mutants and
scoring_utils — runs UNMODIFIED. Nothing is copied from the reference."""
from __future__ import annotations

import contextlib
import os
import sys

from . import ref_shims_tranception as RT

REF = os.path.join(os.environ.get("PG_REFERENCE_ROOT", "/root/reference"), "proteingym", "baselines", "trancepteve")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "trancepteve", "model_pytorch.py"))


def install():
    RT.install()  # transformers / Bio shims
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import trancepteve.model_pytorch as mp
    from trancepteve import config as cfg
    mp.TrancepteveLMHeadModel.init_weights = lambda self: None
    mp.TranceptionModel.init_weights = lambda self: None
    mp.TranceptionModel.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n
    _positional_series_fallback()
    return mp, cfg


def _positional_series_fallback():
    """MSA_processing.gen_alignment (utils/msa_utils.py:310) reads ``series[int]`` on a string-indexed Series, which the pinned pandas
    (1.x) resolved by position; current pandas raises KeyError. Restore that fallback for integer keys on non-integer indexes."""
    import numbers

    import pandas as pd
    if getattr(pd.Series, "_pg_positional_fallback", False):
        return
    orig = pd.Series.__getitem__

    def getitem(self, key):
        try:
            return orig(self, key)
        except KeyError:
            if isinstance(key, numbers.Integral) and not pd.api.types.is_integer_dtype(self.index.dtype):
                return self.iloc[key]
            raise

    pd.Series.__getitem__ = getitem
    pd.Series._pg_positional_fallback = True


@contextlib.contextmanager
def light_pretrained_init():
    import torch
    import transformers
    orig = transformers.PreTrainedModel.__init__

    def light(self, config, *a, **k):
        torch.nn.Module.__init__(self)
        self.config = config

    transformers.PreTrainedModel.__init__ = light
    try:
        yield
    finally:
        transformers.PreTrainedModel.__init__ = orig
