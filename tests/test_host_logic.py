"""CPU: host-side mirror of the reference interface (tokenisation, windows, mutant parsing, checkpoint reading) and that
the C-ABI library loads and exports every symbol include/pgscore.h declares (no compute calls: no GPU here)."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import esm_oracle as O
from proteingym_b200 import _lib, checkpoint, synth
from proteingym_b200.alphabet import ALPHABET
from proteingym_b200.mutants import parse_mutants
from proteingym_b200.windows import optimal_window_starts


def test_alphabet_matches_reference_vocabulary():
    assert ALPHABET.all_toks == O.VOCAB and len(ALPHABET) == 33
    assert (ALPHABET.cls_idx, ALPHABET.padding_idx, ALPHABET.eos_idx, ALPHABET.unk_idx, ALPHABET.mask_idx) == (0, 1, 2, 3, 32)
    seq = synth.random_protein(57, 1)
    assert ALPHABET.tokenize_sequence(seq).tolist() == O.tokenize(seq).tolist()
    assert ALPHABET.tokenize_sequence("").tolist() == [0, 2]
    with pytest.raises(KeyError):
        ALPHABET.tokenize_sequence("MKJ")  # data.py:256-257 indexes tok_to_idx directly
    assert ALPHABET.get_idx("J") == 3  # label_row's get_idx falls back to <unk> (data.py:127-128)


@pytest.mark.parametrize("n", [3, 100, 1024, 1025, 1102, 1535, 1536, 1537, 3425])
def test_windows_match_get_optimal_window(n):
    pos = np.arange(n)
    starts, T = optimal_window_starts(pos, n)
    for i in pos:
        s, e = O.get_optimal_window(int(i), n, 1024)
        assert (starts[i], starts[i] + T) == (s, e)
        assert starts[i] <= i < starts[i] + T


def test_parse_mutants_csr_and_reference_errors():
    seq = "MKVLAAGIC"
    rows, wts, mts, offs = parse_mutants(["M1A", "K2C:V3L", "C9W"], seq)
    assert rows.tolist() == [1, 2, 3, 9] and offs.tolist() == [0, 1, 3, 4]
    assert wts.tolist() == [ALPHABET.get_idx(c) for c in "MKVC"] and mts.tolist() == [ALPHABET.get_idx(c) for c in "ACLW"]
    r2, *_ = parse_mutants(["M5A"], seq, offset_idx=5)  # --offset-idx / start_idx column (compute_fitness.py:307,427)
    assert r2.tolist() == [1]
    with pytest.raises(AssertionError, match="The listed wildtype does not match the provided sequence"):
        parse_mutants(["A1G"], seq)
    with pytest.raises(ValueError):
        parse_mutants(["MxA"], seq)
    with pytest.raises(IndexError):
        parse_mutants(["M99A"], seq)
    e = parse_mutants([], seq)
    assert e[3].tolist() == [0] and len(e[0]) == 0


@pytest.mark.parametrize("kind,fname", [("esm1v", "esm1v_t3_x.pt"), ("esm2", "esm2_t3_x.pt")])
def test_checkpoint_reader_round_trip(tmp_path, kind, fname):
    ffn = 256 if kind == "esm2" else 128  # ESM2 hard-codes ffn = 4 * embed_dim (esm2.py:52)
    arch = synth.EsmArch(kind, 2, 64, 1, ffn, emb_layer_norm_before=(kind == "esm1v"))
    st = synth.write_esm_checkpoint(str(tmp_path / fname), arch, seed=2)
    conf, state, name = checkpoint.load_esm_checkpoint(str(tmp_path / fname))
    assert name == fname[:-3]
    assert (conf.arch, conf.layers, conf.embed_dim, conf.heads, conf.ffn_dim) == ("esm2" if kind == "esm2" else "esm1b", 2, 64, 1, ffn)
    assert conf.emb_layer_norm_before == (kind == "esm1v") and conf.token_dropout
    want = checkpoint.normalise_synth_state(arch, st)
    assert set(want) == set(state)
    for k in want:
        assert torch.equal(want[k], state[k]), k
    # v1 zeroes the <mask> embedding row (pretrained.py:97) and the tied head sees it; v2 does not
    assert bool((state["embed_tokens.weight"][32] == 0).all()) == (kind == "esm1v")
    ostate = O.load_state(st, kind)
    assert torch.equal(ostate["lm_head.weight"], state["embed_tokens.weight"])


def test_checkpoint_reader_rejects_unknown_arch(tmp_path):
    import argparse
    p = tmp_path / "weird.pt"
    torch.save({"args": argparse.Namespace(arch="protein_bert_base"), "model": {}}, str(p))
    with pytest.raises(ValueError):
        checkpoint.load_esm_checkpoint(str(p))
    with pytest.raises(ValueError):
        checkpoint.load_esm_checkpoint("esm1v_t33_650M_UR90S_1")  # hub names need the network


def test_rotary_tables_match_reference_formula():
    inv = 1.0 / (10000 ** (torch.arange(0, 64, 2).float() / 64))
    cos, sin = checkpoint.rotary_tables(inv, 300)
    c2, s2 = O.rotary_tables(300, 64, torch.float32)
    assert torch.equal(cos, c2[:, :32]) and torch.equal(sin, s2[:, :32]) and torch.equal(c2[:, :32], c2[:, 32:])


def test_library_loads_and_exports_every_declared_symbol():
    from __graft_entry__ import build
    build()
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "pgscore.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pg_abi_version() == 5


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "proteingym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "esm_oracle" not in src, f


def test_no_cuda_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from proteingym_b200.esm_engine import EsmScorer
    arch = synth.EsmArch("esm1v", 1, 64, 1, 64)
    with pytest.raises(_lib.PgError):
        EsmScorer(checkpoint.config_from_synth(arch), checkpoint.normalise_synth_state(arch, synth.make_esm_state(arch)))


def test_cli_flag_surface_matches_reference():
    """tests/golden/esm_cli_flags.json was dumped from the reference's create_parser() (compute_fitness.py:100-238)."""
    import json
    from proteingym_b200.compute_fitness import create_parser

    def dump(p):
        out = {}
        for a in p._actions:
            if not a.option_strings or a.dest == "help":
                continue
            out[a.dest] = {"opts": sorted(a.option_strings), "default": str(a.default), "nargs": str(a.nargs),
                           "type": getattr(a.type, "__name__", str(a.type)), "choices": list(a.choices) if a.choices else None,
                           "const": str(a.const)}
        return out
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "esm_cli_flags.json")))
    mine = dump(create_parser())
    assert set(mine) - set(ref) == {"precision", "device"}  # additive flags only
    for k, v in ref.items():
        assert mine[k] == v, k


def test_cli_resolve_assay_errors_match_reference(tmp_path):
    from proteingym_b200.compute_fitness import create_parser, resolve_assay
    synth.write_mapping_csv(str(tmp_path / "map.csv"), [("A1", "a1.csv", "mkv"), ("A1", "a1b.csv", "MKV"), ("B2", "b2.csv", "MKVL")])
    import pandas as pd
    pd.DataFrame({"mutant": []}).to_csv(tmp_path / "b2.csv", index=False)
    base = ["--dms_mapping", str(tmp_path / "map.csv"), "--dms-input", str(tmp_path), "--dms-output", str(tmp_path / "out"), "--model_type", "ESM1v"]
    with pytest.raises(KeyError, match="MSA_filename"):  # the parser's default model type is the MSA Transformer, which needs the MSA columns (:311)
        resolve_assay(create_parser().parse_args(base[:-2] + ["--dms_index", "2"]))
    with pytest.raises(ValueError, match="Multiple mappings found"):
        resolve_assay(create_parser().parse_args(base + ["--dms_index", "0"]))
    with pytest.raises(ValueError, match="No rows found"):
        resolve_assay(create_parser().parse_args(base + ["--dms_index", "2"]))
    synth.write_dms_csv(str(tmp_path / "b2.csv"), "MKVL", ["M1A", "K2C"])
    a = create_parser().parse_args(base + ["--dms_index", "2"])
    df, col, off = resolve_assay(a)
    assert (len(df), col, off, a.sequence) == (2, "mutant", 1, "MKVL") and a.dms_output.endswith("B2.csv")


# ------------------------------------------------------------------------------------------------ Tranception host logic
def test_tranception_slopes_taps_tokens_slices_match_oracle():
    import pandas as pd
    from oracle import tranception_oracle as TO
    from proteingym_b200 import tranception_engine as TE
    for heads in (4, 8, 12, 20, 24):
        assert np.allclose(TE.alibi_slopes(heads), TO.get_slopes(heads), rtol=0, atol=0)
    assert TE.VOCAB == TO.VOCAB and TE.tokenize("MSIQ") == [1, 15, 20, 12, 18, 2] == TO.tokenize("MSIQ")
    # conv taps reproduce the reference-style depthwise conv
    arch = synth.TranceptionArch(1, 256, 4, 256)
    st = {k[len("transformer."):]: v for k, v in synth.make_tranception_state(arch, 3).items() if k.startswith("transformer.")}
    taps = TE.conv_taps(st, 0, 4).view(3, 4, 64, 8)
    x = torch.randn(1, 1, 30, 64)
    for gi, k in enumerate((3, 5, 7)):
        ref = TO.depthwise_causal_conv(x, st[f"h.0.attn.key_depthwiseconv.{gi}.conv.weight"], st[f"h.0.attn.key_depthwiseconv.{gi}.conv.bias"])
        got = torch.zeros_like(x) + taps[1, gi + 1, :, 7]
        for o in range(7):
            shifted = torch.cat([torch.zeros(1, 1, o, 64), x[:, :, :30 - o]], dim=2)
            got = got + shifted * taps[1, gi + 1, :, o]
        assert (got - ref).abs().max() < 1e-5
    assert torch.equal(taps[:, 0, :, 0], torch.ones(3, 64)) and taps[:, 0, :, 1:].abs().sum() == 0
    # slices: same rows, same order as the oracle restatement of get_sequence_slices
    sc = object.__new__(TE.TranceptionScorer)
    sc.n_ctx = 64
    seq = synth.random_protein(150, 2)
    muts = synth.sample_mutants(seq, 40, 1, multi_frac=0.3)
    df = pd.DataFrame({"mutant": muts, "mutated_sequence": [synth.apply_mutant(seq, m) for m in muts]})
    for mode in ("optimal", "sliding"):
        a = sc.slices(df, seq, mode)
        b = TO.sequence_slices(df, seq, 62, scoring_window=mode)
        assert a.equals(b), mode
    ind = synth.random_indels(seq[:50], 20, 3)
    dfi = pd.DataFrame({"mutant": ind, "mutated_sequence": ind})
    assert sc.slices(dfi, seq[:50], "optimal", indel_mode=True).equals(TO.sequence_slices(dfi, seq[:50], 62, indel_mode=True))
    with pytest.raises(AssertionError, match="Invalid from_AA or mutant position"):
        TE.apply_substitutions("MKV", "A1G")


def test_tranception_checkpoint_round_trip(tmp_path):
    from proteingym_b200 import tranception_engine as TE
    arch = synth.TranceptionArch(2, 256, 4, 512)
    st = synth.write_tranception_checkpoint(str(tmp_path / "Tranception_tiny"), arch, seed=2)
    cfg, state = TE.load_tranception_checkpoint(str(tmp_path / "Tranception_tiny"))
    assert (cfg["n_embd"], cfg["n_head"], cfg["n_layer"], cfg["n_inner"]) == (256, 4, 2, 512)
    assert "lm_head.weight" not in state and torch.equal(state["wte.weight"], st["transformer.wte.weight"])
    assert torch.equal(state["h.1.mlp.c_fc.weight"], st["transformer.h.1.mlp.c_fc.weight"])


def test_model_object_seam_host_side():
    """proteingym_b200.pretrained: batch converter = the reference's for one or several sequences (data.py:262-297); mask detection;
    no CPU forward behind the model object."""
    from proteingym_b200 import pretrained
    from proteingym_b200.alphabet import ALPHABET
    labels, strs, tok = ALPHABET.get_batch_converter()([("protein1", "MKV"), ("p2", "ACDEF")])
    assert labels == ["protein1", "p2"] and strs == ["MKV", "ACDEF"] and tok.dtype == torch.int64 and tuple(tok.shape) == (2, 7)
    assert tok[0].tolist() == [0, 20, 15, 7, 2, 1, 1] and tok[1].tolist() == [0, 5, 23, 13, 9, 18, 2]
    row = tok[1].clone()
    assert pretrained.mask_position(row, ALPHABET.mask_idx) == -1
    row[3] = ALPHABET.mask_idx
    assert pretrained.mask_position(row, ALPHABET.mask_idx) == 3
    row[4] = ALPHABET.mask_idx
    with pytest.raises(NotImplementedError):
        pretrained.mask_position(row, ALPHABET.mask_idx)
    m = pretrained.B200EsmModel(None, {}, "x")
    assert m.eval() is m
    with pytest.raises(RuntimeError, match="no CPU forward"):
        m(tok[:1])


def test_tranception_tokenize_batch_equals_per_sequence_tokenize():
    from proteingym_b200.tranception_engine import PAD, tokenize, tokenize_batch
    rng = np.random.RandomState(0)
    seqs = ["", "M", "ACDEFGHIKLMNPQRSTVWY", "MKU*-xZ", "B" * 5] + ["".join(rng.choice(list("ACDEFGHIKLMNPQRSTVWYXBJZ"), size=n)) for n in (3, 17, 64, 200)]
    T = max(len(s) for s in seqs) + 2
    ids, lens = tokenize_batch(seqs, T)
    assert ids.dtype == np.int32 and ids.shape == (len(seqs), T) and lens.dtype == np.int32
    for r, s in enumerate(seqs):
        t = tokenize(s)
        assert lens[r] == len(t) and ids[r, :len(t)].tolist() == t and (ids[r, len(t):] == PAD).all()
    e, l0 = tokenize_batch([], 4)
    assert e.shape == (0, 4) and l0.shape == (0,)


def test_auto_precision_rule():
    """esm_engine.choose_precision: f16d (delta operands) only where it was measured inside the 1e-3 bar — ESM-1b / ESM-1v
    masked-marginals, 192..1022 residues (one shared window), at most two sites per mutant."""
    from proteingym_b200.checkpoint import EsmConfig
    from proteingym_b200.esm_engine import PRECISIONS, choose_precision
    e1 = EsmConfig("esm1b", 33, 1280, 20, 5120, True, False)
    e2 = EsmConfig("esm2", 33, 1280, 20, 5120, True, False)
    e3b = EsmConfig("esm2", 36, 2560, 40, 10240, True, False)
    single, double, triple = ["A5G", "C9D"], ["A5G:C9D", "E3F"], ["A5G:C9D:E3F"]
    assert choose_precision(e1, single, seq_len=512) == "f16d" and choose_precision(e1, double, seq_len=192) == "f16d"
    assert choose_precision(e1, single, seq_len=1022) == "f16d" and choose_precision(e1, single, seq_len=1023) == "f16f8"  # windows
    assert choose_precision(e1, single, seq_len=191) == "f16f8" and choose_precision(e1, single) == "f16f8"  # short / unknown length
    assert choose_precision(e1, triple, seq_len=512) == "f16x3"
    assert choose_precision(e1, single, "wt-marginals", seq_len=512) == "f16f8" and choose_precision(e1, single, "pseudo-ppl", seq_len=512) == "f16x3"
    assert choose_precision(e2, single, seq_len=512) == "f16f8" and choose_precision(e3b, single, seq_len=512) == "f16x3"
    assert set(PRECISIONS) == {"f16", "f16x3", "f16f8", "f16d"}
