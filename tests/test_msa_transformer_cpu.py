"""CPU checks of the MSA Transformer path: the oracle against the unmodified reference's outputs, and the host logic (alignment
sampling, tokenisation, checkpoint key mapping) against what the reference did with the same files."""
import os

import numpy as np
import pytest
import torch

from msa_transformer_cases import SMALL_CASES, case_dir, have, load_case
from oracle import msa_oracle as MO
from proteingym_b200 import checkpoint, msa_engine, synth


def oracle_scores(table_rows, positions, df, sequence, offset):
    """label_row (compute_fitness.py:240-250) over the rows the oracle computed."""
    row_of = {int(p): i for i, p in enumerate(positions)}
    out = []
    for m in df["mutant"]:
        s = 0.0
        for mut in m.split(":"):
            wt, idx, mt = mut[0], int(mut[1:-1]) - offset, mut[-1]
            assert sequence[idx] == wt
            r = table_rows[row_of[1 + idx]]
            s += float(r[MO.TOK[mt]] - r[MO.TOK[wt]])
        out.append(s)
    return np.asarray(out)


@pytest.mark.parametrize("name", SMALL_CASES)
def test_oracle_matches_reference_table_and_scores(name):
    c = load_case(name)
    arch, meta = c["arch"], c["meta"]
    st = synth.make_msa_state(arch, meta["seed"], qk_gain=c["qk_gain"])
    toks = MO.tokenize_alignment(c["rows"])
    tab = MO.masked_marginal_table(st, toks, arch.layers, arch.heads, positions=meta["table_positions"]).numpy()
    assert np.abs(tab - c["table"]).max() < 5e-5
    if name == "window":
        return  # 1101 forwards of the score columns: the table rows above cover the window arithmetic
    col = f"{meta['column']}_seed{meta['seeds'][0]}"
    df = c["df"].iloc[:12 if name == "batched" else 40]  # one oracle forward per mutated column: keep the CPU suite in minutes
    need = sorted({1 + int(mut[1:-1]) - meta["MSA_start"] for m in df["mutant"] for mut in m.split(":")})
    rows = MO.masked_marginal_table(st, toks, arch.layers, arch.heads, positions=need).numpy()
    got = oracle_scores(rows, need, df, c["sequence"], meta["MSA_start"])
    assert np.abs(got - df[col].to_numpy()).max() < 1e-4


def test_sampling_first_rows_and_random_match_reference():
    for name in ("tiny", "batched"):
        c = load_case(name)
        m = c["meta"]
        rows = msa_engine.sample_msa(os.path.join(c["dir"], "alignment.a2m"), m["msa_samples"], m["strategy"], m["seeds"][0])
        assert [tuple(r) for r in rows] == [tuple(r) for r in c["rows"]]
    toks = msa_engine.tokenize_alignment(load_case("tiny")["rows"])
    assert toks.dtype == np.int32 and np.array_equal(toks, MO.tokenize_alignment(load_case("tiny")["rows"]).numpy())
    assert (toks[1:, 8:10] == 29).all()  # '.' columns stay tokens (compute_fitness.py:69-70 upper-cases, nothing is removed)


def test_weighted_sampling_matches_reference_given_its_weights():
    """sequence-reweighting (:39-64): with the reference's weights file the same seed draws the same rows; the ensemble column is the
    mean of the seed columns (:538-542)."""
    c = load_case("weights")
    m = c["meta"]

    class Processed:  # the attributes sample_msa reads from MSA_processing
        pass
    names, seqs = [], []
    for n, s in msa_engine.read_fasta_records(os.path.join(c["dir"], "alignment.a2m")):
        names.append(">" + n)
        seqs.append(s)
    P = Processed()
    P.focus_seq_name = names[0]
    P.raw_seq_name_to_sequence = dict(zip(names, seqs))
    P.seq_name_to_sequence = dict(zip(names, [s.replace(".", "-").upper() for s in seqs]))
    assert len(c["weights"]) == len(names)  # no row of this alignment is filtered out
    P.seq_name_to_weight = dict(zip(names, c["weights"]))
    rows = msa_engine.sample_msa(os.path.join(c["dir"], "alignment.a2m"), m["msa_samples"], "sequence-reweighting", m["seeds"][0],
                                 processed_msa=P)
    assert [tuple(r) for r in rows] == [tuple(r) for r in c["rows"]]
    df = c["df"]
    cols = [f"{m['column']}_seed{s}" for s in m["seeds"]]
    assert np.allclose(df[f"{m['column']}_ensemble"], df[cols].mean(axis=1), atol=1e-12)


def test_checkpoint_loader_maps_file_keys_like_the_reference(tmp_path):
    arch = synth.MsaArch(2, 128, 2, 256, msa_pos_dim=1)
    path = str(tmp_path / "msa_t.pt")
    st = synth.write_msa_checkpoint(path, arch, seed=4)
    blob = torch.load(path, weights_only=False)
    assert "encoder.sentence_encoder.layers.0.column_self_attention.layer.q_proj.weight" in blob["model"]  # stored swapped
    conf, state, name = checkpoint.load_msa_checkpoint(path)
    assert (conf.arch, conf.layers, conf.embed_dim, conf.heads, conf.ffn_dim) == ("msa", 2, 128, 2, 256) and name == "msa_t"
    assert torch.equal(state["layers.0.row_self_attention.layer.q_proj.weight"], st["layers.0.row_self_attention.layer.q_proj.weight"])
    assert torch.equal(state["layers.1.column_self_attention.layer_norm.bias"], st["layers.1.column_self_attention.layer_norm.bias"])
    assert state["msa_position_embedding"].shape == (1024, 128)
    assert torch.equal(state["msa_position_embedding"][:, 5], st["msa_position_embedding"].reshape(1024))
    assert "lm_head.weight" not in state
    same = checkpoint.normalise_msa_synth_state(arch, synth.make_msa_state(arch, 4))
    assert set(same) == set(state) and all(torch.equal(same[k], state[k]) for k in state)


def test_window_arithmetic_matches_reference_slices():
    """C > 1024: window starts / widths used by the engine = get_optimal_window(i, L + 2, 1024) clipped at C (compute_fitness.py:387-391)."""
    from proteingym_b200.windows import optimal_window_starts
    C = 1101
    pos = np.arange(C)
    starts, _ = optimal_window_starts(pos, C + 1, 1024)
    for i in (0, 1, 511, 512, 513, 588, 589, 590, 1000, 1100):
        s, e = MO.optimal_window(i, C + 1, 1024)
        assert starts[i] == s and min(starts[i] + 1024, C) == min(e, C)


def test_oracle_true_size_msa1b_rows_match_reference():
    """12 x 768 (MSA-1b) on 48 x 130 tokens: two rows of the reference's table (a forward is ~6 s of CPU)."""
    if not have("msa1b"):
        pytest.skip("true-size fixture not generated")
    c = load_case("msa1b")
    arch, meta = c["arch"], c["meta"]
    st = synth.make_msa_state(arch, meta["seed"], qk_gain=c["qk_gain"])
    toks = MO.tokenize_alignment(c["rows"])
    pos = meta["table_positions"][5:7]
    tab = MO.masked_marginal_table(st, toks, arch.layers, arch.heads, positions=pos).numpy()
    assert np.abs(tab - c["table"][5:7]).max() < 1e-4
