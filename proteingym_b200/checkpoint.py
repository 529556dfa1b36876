"""fair-esm checkpoint reader for the B200 scorer (host side, replaces the model-construction half of
``pretrained.load_model_and_alphabet``; reference: proteingym/baselines/esm/esm/pretrained.py:24-28,67-77,85-99,
162-181,184-218).

Returns a flat fp32 state (reference key names with the ``encoder.`` / ``encoder.sentence_encoder.`` prefixes removed)
plus the architecture description the C-ABI needs. Nothing here builds an nn.Module.
"""
from __future__ import annotations

import argparse
import os
import re
from dataclasses import dataclass
from pathlib import Path

import torch

MASK_IDX = 32


@dataclass
class EsmConfig:
    arch: str            # "esm1b" (ESM-1b / ESM-1v, learned positions) | "esm2" (rotary) | "msa" (MSA Transformer)
    layers: int
    embed_dim: int
    heads: int
    ffn_dim: int
    token_dropout: bool
    emb_layer_norm_before: bool
    max_positions: int = 1024
    vocab: int = 33


def _strip_v1(name: str) -> str:
    # pretrained.py:90-93: drop everything up to and including "encoder." / "sentence_encoder."
    if "sentence_encoder." in name:
        name = "".join(name.split("sentence_encoder.")[1:])
    if "encoder." in name:
        name = "".join(name.split("encoder.")[1:])
    return name


def checkpoint_column_name(model_location: str) -> str:
    """Name of the score column a checkpoint produces: the reference's ``model_location.split("/")[-1].split(".")[0]``
    (compute_fitness.py:350), i.e. everything before the FIRST dot of the file name (``esm1v.v2.pt`` -> ``esm1v``)."""
    return os.path.basename(str(model_location)).split(".")[0]


def load_esm_checkpoint(model_location: str):
    """-> (EsmConfig, state: dict[str, fp32 CPU tensor], score column name)."""
    path = Path(model_location)
    if not str(model_location).endswith(".pt"):
        raise ValueError("only local .pt checkpoints are supported (no network): " + str(model_location))
    with torch.serialization.safe_globals([argparse.Namespace]):
        data = torch.load(str(path), map_location="cpu", weights_only=True)
    name = path.stem
    if name.startswith("esm2"):  # pretrained.py:187
        cfg = data["cfg"]["model"]
        pat = re.compile("^" + "|".join(["encoder.sentence_encoder.", "encoder."]))
        raw = {pat.sub("", k): v for k, v in data["model"].items()}
        conf = EsmConfig("esm2", int(cfg.encoder_layers), int(cfg.encoder_embed_dim), int(cfg.encoder_attention_heads),
                         4 * int(cfg.encoder_embed_dim), bool(cfg.token_dropout), False)
    else:
        args = data["args"]
        if args.arch != "roberta_large":
            raise ValueError(f"unsupported fair-esm architecture {args.arch!r} (ESM-1b/ESM-1v/ESM2 only)")
        raw = {_strip_v1(k): v for k, v in data["model"].items()}
        get = lambda n, dflt=None: getattr(args, n, getattr(args, "encoder_" + n, dflt))
        # pretrained.py:97: zero the <mask> embedding in place ("for token drop"); aliasing with lm_head.weight is kept
        raw["embed_tokens.weight"][MASK_IDX].zero_()
        conf = EsmConfig("esm1b", int(get("layers")), int(get("embed_dim")), int(get("attention_heads")),
                         int(get("ffn_embed_dim")), bool(getattr(args, "token_dropout", False)),
                         any(k.startswith("emb_layer_norm_before") for k in raw), int(get("max_positions", 1024)))
    # The tied output matrix: load_state_dict copies "lm_head.weight" into the shared parameter after
    # "embed_tokens.weight" (lm_head is registered last, esm1.py:97-101 / esm2.py:70-74), so it wins.
    tied = raw["lm_head.weight"] if "lm_head.weight" in raw else raw["embed_tokens.weight"]
    state = {k: v.detach().to(torch.float32).contiguous() for k, v in raw.items()
             if not k.startswith("contact_head") and k != "lm_head.weight"}
    state["embed_tokens.weight"] = tied.detach().to(torch.float32).contiguous()
    return conf, state, checkpoint_column_name(model_location)


def _msa_model_key(name: str) -> str:
    """File key -> model key of an msa_transformer checkpoint (pretrained.py:107-115): the loader swaps "row" and "column" in every
    name (the first release stored the two attention blocks under each other's names), then strips the encoder prefixes."""
    if "row" in name:
        name = name.replace("row", "column")
    else:
        name = name.replace("column", "row")
    return _strip_v1(name)


def load_msa_checkpoint(model_location: str):
    """-> (EsmConfig with arch "msa", state: dict[str, fp32 CPU tensor], score column name) for a fair-esm ``msa_transformer`` file.
    ``msa_position_embedding`` is returned as [1024, embed_dim] (the first release's [1, 1024, 1, 1] table broadcasts over the width,
    msa_transformer.py:166-168); absent when the checkpoint was trained without it."""
    path = Path(model_location)
    if not str(model_location).endswith(".pt"):
        raise ValueError("only local .pt checkpoints are supported (no network): " + str(model_location))
    with torch.serialization.safe_globals([argparse.Namespace]):
        data = torch.load(str(path), map_location="cpu", weights_only=True)
    args = data["args"]
    if args.arch != "msa_transformer":
        raise ValueError(f"not an msa_transformer checkpoint: {args.arch!r}")
    raw = {_msa_model_key(k): v for k, v in data["model"].items()}
    get = lambda n, dflt=None: getattr(args, n, getattr(args, "encoder_" + n, dflt))
    conf = EsmConfig("msa", int(get("layers")), int(get("embed_dim")), int(get("attention_heads")), int(get("ffn_embed_dim")),
                     False, True, int(get("max_positions", 1024)))
    tied = raw["lm_head.weight"] if "lm_head.weight" in raw else raw["embed_tokens.weight"]
    state = {k: v.detach().to(torch.float32).contiguous() for k, v in raw.items()
             if not k.startswith("contact_head") and k != "lm_head.weight"}
    state["embed_tokens.weight"] = tied.detach().to(torch.float32).contiguous()
    if get("embed_positions_msa", False) and "msa_position_embedding" in state:
        state["msa_position_embedding"] = normalise_row_positions(state["msa_position_embedding"], conf.embed_dim)
    else:
        state.pop("msa_position_embedding", None)
    return conf, state, checkpoint_column_name(model_location)


def normalise_row_positions(t: torch.Tensor, d: int) -> torch.Tensor:
    t = t.to(torch.float32).reshape(1024, -1)
    return t.expand(1024, d).contiguous()


def config_from_msa_synth(arch) -> EsmConfig:
    return EsmConfig("msa", arch.layers, arch.embed_dim, arch.heads, arch.ffn_dim, False, True, arch.max_positions, arch.vocab)


def normalise_msa_synth_state(arch, st: dict) -> dict:
    """Apply to an in-memory ``synth.make_msa_state`` dict what ``load_msa_checkpoint`` does to a file."""
    st = {k: v.to(torch.float32).contiguous() for k, v in st.items() if k != "lm_head.weight"}
    if "msa_position_embedding" in st:
        st["msa_position_embedding"] = normalise_row_positions(st["msa_position_embedding"], arch.embed_dim)
    return st


def config_from_synth(arch) -> EsmConfig:
    """EsmConfig for a ``synth.EsmArch`` (tests / bench build states in memory instead of writing 2.6 GB files)."""
    return EsmConfig("esm2" if arch.kind == "esm2" else "esm1b", arch.layers, arch.embed_dim, arch.heads, arch.ffn_dim,
                     arch.token_dropout, arch.emb_layer_norm_before, arch.max_positions, arch.vocab)


def normalise_synth_state(arch, st: dict) -> dict:
    """Apply to an in-memory ``synth.make_esm_state`` dict what ``load_esm_checkpoint`` does to a file."""
    st = {k: v.clone() for k, v in st.items()}
    if arch.kind != "esm2":
        st["embed_tokens.weight"][MASK_IDX].zero_()
    st.pop("lm_head.weight", None)  # tied alias of embed_tokens.weight in synth states
    return {k: v.to(torch.float32).contiguous() for k, v in st.items()}


def rotary_tables(inv_freq: torch.Tensor, T: int):
    """cos/sin [T, 32] exactly as the reference builds them (esm/rotary_embedding.py:46-61): fp32 outer product of
    arange(T) with the checkpoint's ``inv_freq`` buffer, then cos / sin. Both halves of the 64-wide table are equal."""
    t = torch.arange(T).type_as(inv_freq)
    freqs = torch.einsum("i,j->ij", t, inv_freq.float())
    return freqs.cos().contiguous(), freqs.sin().contiguous()
