"""ORACLE SUPPORT — TEST INFRASTRUCTURE ONLY (build container only: needs /root/reference).

Import shims that make the reference's Tranception modules importable under transformers 5.x (the reference pins 4.32.1,
environments/proteingym_env.txt:119): the plain-nn.Module pieces (TranceptionBlock, SpatialDepthWiseConvolution,
get_slopes) and tranception/utils/scoring_utils.py then run unmodified. ``TranceptionLMHeadModel`` itself cannot be
constructed (GPT2PreTrainedModel API drift), so golden vectors come from a hybrid: reference blocks + reference scoring
arithmetic around a thin wrapper (oracle/gen_golden_tranception.py). Nothing is copied from the reference."""
from __future__ import annotations

import os
import sys
import types

REF = os.path.join(os.environ.get("PG_REFERENCE_ROOT", "/root/reference"), "proteingym", "baselines", "tranception")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "tranception", "model_pytorch.py"))


def install():
    if "tranception.model_pytorch" in sys.modules:
        return sys.modules["tranception.model_pytorch"], sys.modules["tranception.utils.scoring_utils"]
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    if not hasattr(mu, "Conv1D"):
        mu.Conv1D = pu.Conv1D
    for n in ("find_pruneable_heads_and_indices", "prune_conv1d_layer"):
        if not hasattr(mu, n):
            setattr(mu, n, getattr(pu, n, lambda *a, **k: (_ for _ in ()).throw(NotImplementedError(n))))
    if not hasattr(mu, "SequenceSummary"):
        mu.SequenceSummary = None
    import transformers.file_utils as fu
    for n in ("add_code_sample_docstrings", "add_start_docstrings", "add_start_docstrings_to_model_forward", "replace_return_docstrings"):
        if not hasattr(fu, n):
            setattr(fu, n, lambda *a, **k: (lambda f: f))
    if "transformers.utils.model_parallel_utils" not in sys.modules:
        m = types.ModuleType("transformers.utils.model_parallel_utils")
        m.assert_device_map = lambda *a, **k: None
        m.get_device_map = lambda *a, **k: None
        sys.modules["transformers.utils.model_parallel_utils"] = m
    for name in ("Bio", "Bio.Align", "Bio.Align.Applications", "Bio.SeqIO", "Bio.Seq", "Bio.SeqRecord"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["Bio.Align.Applications"].ClustalOmegaCommandline = object
    sys.modules["Bio"].SeqIO = sys.modules["Bio.SeqIO"]
    sys.modules["Bio.SeqRecord"].SeqRecord = object
    sys.modules["Bio.Seq"].Seq = object
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import tranception.model_pytorch as mp
    from tranception.utils import scoring_utils as su
    return mp, su


def tokenizer():
    from transformers import PreTrainedTokenizerFast
    return PreTrainedTokenizerFast(tokenizer_file=os.path.join(REF, "tranception", "utils", "tokenizers", "Basic_tokenizer"),
                                   unk_token="[UNK]", sep_token="[SEP]", pad_token="[PAD]", cls_token="[CLS]", mask_token="[MASK]")
