"""GPU (-m gpu): Tranception autoregressive scoring through the C-ABI (pg_ar_loglik) against the oracle and the
hybrid-reference golden vectors (oracle/gen_golden_tranception.py). Tolerance 1e-3 abs on scores (parity mode)."""
import json
import os

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import GOLDEN
from oracle import tranception_oracle as TO
from proteingym_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-3


def make_scorer(arch, raw_state, precision="f16f8", max_rows=32768):
    from proteingym_b200.tranception_engine import TranceptionScorer
    cfg = {"n_embd": arch.embed_dim, "n_head": arch.heads, "n_layer": arch.layers, "n_ctx": arch.n_ctx, "n_inner": arch.ffn_dim,
           "vocab_size": arch.vocab, "layer_norm_epsilon": arch.ln_eps, "activation_function": "squared_relu"}
    st = {k[len("transformer."):]: v for k, v in raw_state.items() if k.startswith("transformer.")}
    return TranceptionScorer(cfg, st, precision=precision, max_rows=max_rows)


def load(name):
    meta = json.load(open(os.path.join(GOLDEN, f"{name}_meta.json")))
    arch = synth.TranceptionArch(**meta["arch"])
    return meta, arch, synth.make_tranception_state(arch, meta["seed"]), pd.read_csv(os.path.join(GOLDEN, f"{name}_dms.csv")), \
        pd.read_csv(os.path.join(GOLDEN, f"{name}_reference_scores.csv"))


@pytest.mark.parametrize("precision,tol", [("f16f8", 2e-4), ("f16x3", 2e-4), ("f16", 5e-2)])
def test_sequence_logprobs_match_oracle(precision, tol):
    arch = synth.TranceptionArch(2, 256, 4, 512)
    st = synth.make_tranception_state(arch, 9)
    sc = make_scorer(arch, st, precision)
    seqs = [synth.random_protein(L, L) for L in (1, 5, 33, 64, 65, 130, 131, 257)]  # ragged batch -> right padding
    got = sc.sequence_logprobs(seqs)
    sc.close()
    want = [TO.sequence_logprob(st, s, arch.layers, arch.heads, dtype=torch.float64) for s in seqs]
    rel = np.abs(got - np.asarray(want)) / np.maximum(1.0, np.abs(want) / 50)
    assert rel.max() < tol, (got, want)


@pytest.mark.parametrize("name", ["tranception_subs", "tranception_indels", "tranception_long"])
def test_score_mutants_matches_reference_hybrid_golden(name):
    meta, arch, st, dms, ref = load(name)
    sc = make_scorer(arch, st, max_rows=65536)
    got = sc.score_mutants(dms, meta["target_seq"], indel_mode=meta["indel_mode"], scoring_window=meta["scoring_window"])
    sc.close()
    assert list(got.columns) == list(ref.columns) and len(got) == len(ref)
    assert list(got[got.columns[0]]) == list(ref[ref.columns[0]])  # same rows in the same order
    for c in ("avg_score_L_to_R", "avg_score_R_to_L", "avg_score"):
        assert np.abs(got[c].to_numpy(dtype=np.float64) - ref[c].to_numpy(dtype=np.float64)).max() < TOL, c


@pytest.mark.parametrize("name", ["tranception_L_subs", "tranception_L_indels", "tranception_L_long"])
def test_true_size_tranception_l_matches_reference_class(name):
    """BASELINE config 4 architecture at TRUE SIZE (36 x 1280, 20 heads): scores of the unmodified TranceptionLMHeadModel
    (oracle/gen_golden_trancepteve.py truesize) — substitutions incl. multi-mutants, ragged indels, a > n_ctx protein."""
    meta, arch, st, dms, ref = load(name)
    for reuse in (True, False):
        sc = make_scorer(arch, st, max_rows=32768)
        sc.prefix_reuse = reuse
        got = sc.score_mutants(dms, meta["target_seq"], indel_mode=meta["indel_mode"], scoring_window=meta["scoring_window"])
        rows = sc.reuse_rows
        sc.close()
        assert list(got[got.columns[0]]) == list(ref[ref.columns[0]])
        err = max(np.abs(got[c].to_numpy(dtype=np.float64) - ref[c].to_numpy(dtype=np.float64)).max()
                  for c in ("avg_score_L_to_R", "avg_score_R_to_L", "avg_score"))
        print(f"\n{name} (prefix reuse {reuse}, token rows run / plain = {rows}): max |score - reference| = {err:.2e}")
        assert err < TOL


@pytest.mark.parametrize("precision", ["f16f8", "f16"])
def test_prefix_reuse_equals_plain_path(precision):
    """Exact wild-type-prefix reuse (pg_ar_prefix_begin / pg_ar_loglik_prefix): same scores as the plain path — the rows are
    bit-identical, only the per-sequence summation is split — and fewer token rows run. With and without a retrieval prior."""
    arch = synth.TranceptionArch(2, 256, 4, 512)
    st = synth.make_tranception_state(arch, 11)
    seq = synth.random_protein(420, 6)  # T = 422: group starts 0, 128, 256, 384
    muts = synth.sample_mutants(seq, 120, 4, multi_frac=0.25) + [f"{seq[127]}128{'A' if seq[127] != 'A' else 'C'}",
                                                                  f"{seq[126]}127{'A' if seq[126] != 'A' else 'C'}"]  # first change at / next to a tile edge
    dms = pd.DataFrame({"mutant": muts, "mutated_sequence": [synth.apply_mutant(seq, m) for m in muts]})
    rng = np.random.RandomState(1)
    prior = np.log(rng.dirichlet(np.ones(25), size=420)).astype(np.float32)
    for kw in ({}, dict(log_prior=prior, retrieval_inference_weight=0.6, MSA_start=30, MSA_end=400)):
        res = {}
        for reuse in (True, False):
            sc = make_scorer(arch, st, precision)
            sc.prefix_reuse = reuse
            res[reuse] = (sc.score_mutants(dms, seq, **kw), sc.reuse_rows)
            sc.close()
        a, b = res[True][0], res[False][0]
        assert list(a["mutated_sequence"]) == list(b["mutated_sequence"])
        for c in ("avg_score_L_to_R", "avg_score_R_to_L", "avg_score"):
            assert np.abs(a[c].to_numpy(dtype=np.float64) - b[c].to_numpy(dtype=np.float64)).max() < 2e-6, c
        full, run = res[True][1]
        assert run < 0.8 * full and res[False][1][1] == res[False][1][0]


def test_prefix_reuse_rejects_bad_calls():
    import ctypes as C
    from proteingym_b200 import _lib
    arch = synth.TranceptionArch(1, 256, 4, 256)
    sc = make_scorer(arch, synth.make_tranception_state(arch, 2))
    ids = torch.ones(300, dtype=torch.int32, device="cuda")
    out = torch.zeros(300, device="cuda")
    lens = torch.full((2,), 44, dtype=torch.int32, device="cuda")
    lib = sc.lib
    assert lib.pg_ar_loglik_prefix(sc.handle, ids.data_ptr(), lens.data_ptr(), 2, 44, 128, None, out.data_ptr(), None) == 3  # nothing recorded
    _lib.check(lib.pg_ar_prefix_begin(sc.handle, ids.data_ptr(), 300, None, out.data_ptr(), None), sc.handle)
    for start in (0, 100, 384):  # not positive / not a multiple of 128 / not below the recorded length
        assert lib.pg_ar_loglik_prefix(sc.handle, ids.data_ptr(), lens.data_ptr(), 2, 44, start, None, out.data_ptr(), None) == 1
    assert lib.pg_ar_loglik_prefix(sc.handle, ids.data_ptr(), lens.data_ptr(), 2, 44, 256, None, out.data_ptr(), None) == 0
    torch.cuda.synchronize()
    sc.close()


def test_row_sharding_two_gpus_matches_single_gpu():
    """SURVEY.md §8e for Tranception: the sequence rows of one assay split over 2 GPUs + one NCCL all-gather per direction give the
    single-GPU scores. Needs 2 devices (skipped on the 1-GPU box; runs under `gpurun --gpus 2`)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29673", os.path.join(root, "scripts", "check_tranception_sharding.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ok=True" in r.stdout


def test_retrieval_fusion_matches_oracle():
    arch = synth.TranceptionArch(1, 256, 4, 256)
    st = synth.make_tranception_state(arch, 4)
    seq = synth.random_protein(80, 5)
    muts = synth.sample_mutants(seq, 25, 2, multi_frac=0.2)
    dms = pd.DataFrame({"mutant": muts, "mutated_sequence": [synth.apply_mutant(seq, m) for m in muts]})
    rng = np.random.RandomState(0)
    prior = np.log(rng.dirichlet(np.ones(25), size=80)).astype(np.float32)  # [L_full, 25]
    sc = make_scorer(arch, st)
    got = sc.score_mutants(dms, seq, log_prior=prior, retrieval_inference_weight=0.6, MSA_start=10, MSA_end=70)
    sc.close()
    want = TO.score_mutants(st, dms, seq, arch.layers, arch.heads, arch.n_ctx, dtype=torch.float64, log_prior=torch.from_numpy(prior).double(),
                            alpha=0.6, msa_start=10, msa_end=70)
    for c in ("avg_score_L_to_R", "avg_score_R_to_L", "avg_score"):
        assert np.abs(got[c].to_numpy(dtype=np.float64) - want[c].to_numpy(dtype=np.float64)).max() < TOL, c


def test_cli_matches_golden(tmp_path):
    from proteingym_b200 import score_tranception_proteingym as cli
    meta, arch, st, dms, ref = load("tranception_subs")
    synth.write_tranception_checkpoint(str(tmp_path / "Tranception_tiny"), arch, state=st)
    dms.to_csv(tmp_path / "assay.csv", index=False)
    synth.write_mapping_csv(str(tmp_path / "map.csv"), [("A0", "x.csv", "MKV"), ("ASSAY", "assay.csv", meta["target_seq"])])
    cli.main(["--checkpoint", str(tmp_path / "Tranception_tiny"), "--DMS_reference_file_path", str(tmp_path / "map.csv"), "--DMS_index", "1",
              "--DMS_data_folder", str(tmp_path), "--output_scores_folder", str(tmp_path / "out")])
    got = pd.read_csv(tmp_path / "out" / "ASSAY.csv")
    assert list(got.columns) == ["mutated_sequence", "avg_score_L_to_R", "avg_score_R_to_L", "avg_score"]
    assert np.abs(got["avg_score"].to_numpy() - ref["avg_score"].to_numpy()).max() < TOL


def test_true_size_tranception_l_properties():
    """Tranception-L architecture (36 x 1280, 20 heads): no CPU oracle at this size in the suite's time budget; check
    determinism, batching independence and the WT row convention on a short protein."""
    arch = synth.TRANCEPTION_L
    st = synth.make_tranception_state(arch, 0)
    seq = synth.random_protein(120, 1)
    muts = synth.sample_mutants(seq, 60, 2)
    dms = pd.DataFrame({"mutant": muts, "mutated_sequence": [synth.apply_mutant(seq, m) for m in muts]})
    sc = make_scorer(arch, st, max_rows=16384)
    a = sc.score_mutants(dms, seq)
    b = sc.score_mutants(dms, seq)
    assert a.equals(b) and len(a) == len(muts) and np.isfinite(a["avg_score"]).all()
    lp1 = sc.sequence_logprobs([seq, seq[:50], seq])
    assert lp1[0] == lp1[2] and np.isfinite(lp1).all()
    sc.close()
    sc2 = make_scorer(arch, st, max_rows=1024)  # forces one sequence per pass
    lp2 = sc2.sequence_logprobs([seq, seq[:50], seq])
    sc2.close()
    assert np.array_equal(lp1, lp2)


def test_msa_prior_and_cluster_weights_kernels_match_reference(tmp_path):
    """pg_msa_prior / pg_msa_cluster_neighbors against the unmodified reference functions' outputs (oracle/gen_golden_msa.py)
    and, at a larger size, against the numpy oracle."""
    from proteingym_b200 import msa_prior as MP
    meta = json.load(open(os.path.join(GOLDEN, "msa_meta.json")))
    msa = synth.synthetic_msa(meta["target_seq"], meta["msa_n"], seed=meta["msa_seed"])
    synth.write_a2m(str(tmp_path / "m.a2m"), msa)
    got = MP.msa_prior(MP.read_a2m(str(tmp_path / "m.a2m")), meta["MSA_start"], meta["MSA_end"], meta["len_target_seq"])
    ref = np.load(os.path.join(GOLDEN, "msa_prior_reference.npy"))
    assert np.abs(got - ref).max() < 1e-14
    lp = MP.msa_log_prior(str(tmp_path / "m.a2m"), meta["MSA_start"], meta["MSA_end"], meta["len_target_seq"])
    assert lp.dtype == np.float32 and np.isneginf(lp[0]).all() and np.isfinite(lp[meta["MSA_start"]:meta["MSA_end"]]).all()
    g = np.load(os.path.join(GOLDEN, "msa_weights_reference.npz"))
    w = MP.cluster_weights(g["matrix"].astype(np.int64), meta["identity_threshold"])
    assert np.array_equal(w, g["weights"])
    # weighted prior + bigger, ragged-size problems vs the numpy oracle (bit-exact integer counts)
    seq = synth.random_protein(333, 3)
    big = synth.synthetic_msa(seq, 700, seed=9)
    names = list(big)
    wts = {n: 0.05 + (i % 7) / 7.0 for i, n in enumerate(names) if i % 5}
    a = MP.msa_prior(big, 0, 333, 333, weights=wts)
    b = TO.msa_prior(big, 0, 333, 333, weights=wts)
    assert np.abs(a - b).max() < 1e-13
    mat = MP.encode([big[n] for n in names], unknown=0).astype(np.int64)  # '-' and 'X' -> 0 = gap
    for thr in (0.8, 0.99, 0.3):
        assert np.array_equal(MP.cluster_weights(mat, thr), TO.cluster_weights(mat, thr)), thr
