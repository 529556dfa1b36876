"""CPU: the oracle restatement (oracle/esm_oracle.py) against golden vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py ran /root/reference's compute_fitness.main + fair-esm modules on the same seeded checkpoints)."""
import json
import os

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import GOLDEN, GOLDEN_SMALL, GOLDEN_WINDOW, load_golden
from oracle import esm_oracle as O
from proteingym_b200 import synth


def _kind(arch):
    return "esm2" if arch.kind == "esm2" else "esm1v"


@pytest.mark.parametrize("name", GOLDEN_SMALL)
def test_oracle_table_and_scores_match_reference(name):
    g = load_golden(name)
    arch, seq = g["arch"], g["seq"]
    st = O.load_state(g["state"](), _kind(arch))
    table = O.masked_marginal_table(st, seq, _kind(arch), arch.layers, arch.heads, arch.token_dropout)
    assert table.shape == g["table"].shape == (len(seq) + 2, 33)
    assert np.abs(table.numpy() - g["table"]).max() < 5e-5  # fp32 summation-order noise only
    col = g["meta"]["ckpt_names"][0].split(".")[0]
    got = O.score_mutants(g["df"]["mutant"], seq, table)
    assert np.abs(got - g["df"][col].to_numpy()).max() < 1e-4


def test_oracle_ensemble_column_matches_reference():
    # compute_fitness.py:530-537: Ensemble_ESM1v = mean of the per-checkpoint columns
    g = load_golden("tiny_esm1v")
    arch, seq, df = g["arch"], g["seq"], g["df"]
    cols = [c.split(".")[0] for c in g["meta"]["ckpt_names"]]
    assert "Ensemble_ESM1v" in df.columns and len(cols) == 2
    acc = np.zeros(len(df))
    for c, seed in zip(cols, (g["meta"]["seed"], g["meta"]["extra_seed"])):
        st = O.load_state(g["state"](seed), "esm1v")
        t = O.masked_marginal_table(st, seq, "esm1v", arch.layers, arch.heads)
        s = O.score_mutants(df["mutant"], seq, t)
        assert np.abs(s - df[c].to_numpy()).max() < 1e-4
        acc += s
    assert np.abs(acc / 2 - df["Ensemble_ESM1v"].to_numpy()).max() < 1e-4


@pytest.mark.parametrize("name", GOLDEN_WINDOW)
def test_oracle_windowed_rows_match_reference(name):
    # L+2 = 1102 > 1024: per-position optimal windows (compute_fitness.py:492-495); check a spread of positions
    g = load_golden(name)
    arch, seq = g["arch"], g["seq"]
    st = O.load_state(g["state"](), _kind(arch))
    pos = [0, 1, 300, 511, 512, 513, 560, 589, 590, 591, 700, 1100, 1101]
    table = O.masked_marginal_table(st, seq, _kind(arch), arch.layers, arch.heads, positions=pos)
    assert np.abs(table[pos].numpy() - g["table"][pos]).max() < 5e-5


def test_oracle_true_size_rows_match_reference():
    # BASELINE config 1 at true ESM-1v 650M size: a few rows of the 288-row table the reference produced
    g = load_golden("blat_esm1v_650m")
    arch, seq = g["arch"], g["seq"]
    st = O.load_state(g["state"](), "esm1v")
    pos = [24, 150, 286]
    table = O.masked_marginal_table(st, seq, "esm1v", arch.layers, arch.heads, positions=pos, batch=3)
    assert np.abs(table[pos].numpy() - g["table"][pos]).max() < 2e-4


def test_oracle_fp64_agrees_with_fp32():
    g = load_golden("tiny_esm2")
    arch, seq = g["arch"], g["seq"]
    t32 = O.masked_marginal_table(O.load_state(g["state"](), "esm2"), seq, "esm2", arch.layers, arch.heads, positions=range(1, 20))
    t64 = O.masked_marginal_table(O.load_state(g["state"](), "esm2", torch.float64), seq, "esm2", arch.layers, arch.heads,
                                  dtype=torch.float64, positions=range(1, 20))
    assert (t32[1:20].double() - t64[1:20]).abs().max() < 5e-5


# ------------------------------------------------------------------------------------------------ Tranception
def _load_tranception(name):
    import json, os
    import pandas as pd
    from conftest import GOLDEN
    from proteingym_b200 import synth
    meta = json.load(open(os.path.join(GOLDEN, f"{name}_meta.json")))
    arch = synth.TranceptionArch(**meta["arch"])
    return dict(meta=meta, arch=arch, state=synth.make_tranception_state(arch, meta["seed"]),
                dms=pd.read_csv(os.path.join(GOLDEN, f"{name}_dms.csv")),
                scores=pd.read_csv(os.path.join(GOLDEN, f"{name}_reference_scores.csv")),
                logits=np.load(os.path.join(GOLDEN, f"{name}_padded_batch_logits.npy")))


@pytest.mark.parametrize("name", ["tranception_subs", "tranception_indels", "tranception_long"])
def test_tranception_oracle_matches_reference_hybrid(name):
    from oracle import tranception_oracle as TO
    g = _load_tranception(name)
    arch, seq = g["arch"], g["meta"]["target_seq"]
    # forward: the reference scored a right-padded batch of two sequences; unpadded evaluation must agree on real tokens
    for row, s in enumerate((seq[:min(len(seq), 60)], seq[:23])):
        ids = torch.tensor([TO.tokenize(s)])
        lg = TO.forward(g["state"], ids, arch.layers, arch.heads, arch.ln_eps)[0].numpy()
        assert np.abs(lg - g["logits"][row, :len(s) + 2]).max() < 2e-4
    got = TO.score_mutants(g["state"], g["dms"], seq, arch.layers, arch.heads, arch.n_ctx, indel_mode=g["meta"]["indel_mode"],
                           scoring_window=g["meta"]["scoring_window"])
    ref = g["scores"]
    assert list(got.columns) == list(ref.columns) and len(got) == len(ref)
    key = got.columns[0] if got.columns[0] in ("mutated_sequence", "mutant") else "mutated_sequence"
    m = pd.merge(ref, got, on="mutated_sequence", suffixes=("_ref", "")) if "mutated_sequence" in ref else None
    for c in ("avg_score_L_to_R", "avg_score_R_to_L", "avg_score"):
        a = ref[c].to_numpy(dtype=np.float64)
        b = got[c].to_numpy(dtype=np.float64)
        assert np.abs(a - b).max() < 2e-5, c  # same row order as the reference


def test_tranception_slopes_match_reference_list():
    from oracle import tranception_oracle as TO
    assert TO.get_slopes(20) == [0.25, 0.0625, 0.015625, 0.00390625, 0.5] * 4  # SURVEY.md hard-parts note; model_pytorch.py:59-71
    assert len(TO.get_slopes(12)) == 12 and TO.get_slopes(4) == [2 ** -8] * 4


def test_msa_prior_and_weights_oracle_match_reference():
    import json, os
    from conftest import GOLDEN
    from oracle import tranception_oracle as TO
    from proteingym_b200 import synth
    meta = json.load(open(os.path.join(GOLDEN, "msa_meta.json")))
    msa = synth.synthetic_msa(meta["target_seq"], meta["msa_n"], seed=meta["msa_seed"])
    prior = TO.msa_prior(msa, meta["MSA_start"], meta["MSA_end"], meta["len_target_seq"])
    ref = np.load(os.path.join(GOLDEN, "msa_prior_reference.npy"))
    assert prior.shape == ref.shape and np.abs(prior - ref).max() < 1e-15
    g = np.load(os.path.join(GOLDEN, "msa_weights_reference.npz"))
    w = TO.cluster_weights(g["matrix"].astype(np.int64), meta["identity_threshold"])
    assert np.array_equal(w, g["weights"]) and w[5] == 0


def test_tranception_oracle_retrieval_vs_real_reference_class(tmp_path):
    """tests/golden/tranception_retrieval comes from the reference's real TranceptionLMHeadModel (oracle/gen_golden_trancepteve.py
    tranception): weighted MSA log prior fused with alpha = 0.6 on every vocabulary column (model_pytorch.py:806-830)."""
    from trancepteve_cases import make_inputs
    from oracle import tranception_oracle as TO
    gd = os.path.join(GOLDEN, "tranception_retrieval")
    meta = json.load(open(os.path.join(gd, "meta.json")))
    case = meta["case"]
    a = case["arch"]
    arch = synth.TranceptionArch(a[0], a[1], a[2], a[3], n_ctx=a[4])
    st = synth.make_tranception_state(arch, meta["tranception_seed"])
    inp = make_inputs(case, str(tmp_path))
    lp = torch.tensor(np.load(os.path.join(gd, "msa_log_prior.npy")))
    out = TO.score_mutants(st, inp["dms"], meta["target_seq"], arch.layers, arch.heads, n_ctx=arch.n_ctx, log_prior=lp, alpha=0.6,
                           msa_start=case["msa"][0], msa_end=case["msa"][1])
    ref = pd.read_csv(os.path.join(gd, "reference_scores.csv"))
    assert list(out.mutated_sequence) == list(ref.mutated_sequence)
    for c in ("avg_score_L_to_R", "avg_score_R_to_L", "avg_score"):
        assert np.abs(out[c].values - ref[c].values).max() < 2e-5, c
