"""GPU (-m gpu): the retrieval / TranceptEVE rows end to end through the C-ABI (pg_ar_loglik_fused, pg_msa_prior,
pg_msa_cluster_neighbors) against outputs of the reference's UNMODIFIED TrancepteveLMHeadModel / TranceptionLMHeadModel classes
(tests/golden/trancepteve_*, tests/golden/tranception_retrieval; oracle/gen_golden_trancepteve.py). Score tolerance 1e-3 abs."""
import json
import os
import pickle

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import GOLDEN
from trancepteve_cases import CASES, make_inputs

from proteingym_b200 import eve_prior, synth

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _setup(name, tmp_path, with_weights=True, with_cache=True):
    """Files of a case as the reference's launcher would find them: MSA, weights .npy, Tranception checkpoint folder, EVE
    checkpoints (+ the per-model log-prior caches holding the reference's own samples)."""
    case = CASES[name]
    meta = json.load(open(os.path.join(GOLDEN, name, "meta.json")))
    gd = os.path.join(GOLDEN, name)
    inp = make_inputs(case, str(tmp_path), weights_file=os.path.join(gd, "msa_weights.npy") if with_weights else None)
    ck = str(tmp_path / "ckpt")
    synth.write_tranception_checkpoint(ck, inp["arch"], meta["tranception_seed"])
    paths = []
    for i, sd in enumerate(case["eve_seeds"]):
        pth = os.path.join(inp["eve_dir"], f"TARGET_msa_seed_{sd}")
        torch.save({"model_state_dict": synth.make_eve_state(meta["focus_seq_len"], seed=100 + sd)}, pth)
        paths.append(pth)
        if with_cache:
            loc = eve_prior.cache_location(pth, case["n_samples"])
            os.makedirs(os.path.dirname(loc), exist_ok=True)
            with open(loc, "wb") as fh:
                pickle.dump(torch.from_numpy(np.load(os.path.join(gd, f"eve_log_prior_model{i}.npy"))), fh)
    return case, meta, gd, inp, ck, paths


def _scorer(case, meta, inp, ck, paths, precision="f16f8"):
    from proteingym_b200.tranception_engine import load_tranception_checkpoint
    from proteingym_b200.trancepteve_engine import TranceptEVEScorer
    config, state = load_tranception_checkpoint(ck)
    return TranceptEVEScorer(config, state, full_target_seq=meta["target_seq"], inference_time_retrieval_type=case["kind"],
                             retrieval_aggregation_mode="aggregate_substitution", MSA_filename=inp["msa_file"],
                             MSA_weight_file_name=inp["weights_file"], MSA_start=case["msa"][0], MSA_end=case["msa"][1],
                             MSA_threshold_sequence_frac_gaps=case["seq_thr"], MSA_threshold_focus_cols_frac_gaps=case["col_thr"],
                             EVE_model_paths=paths or None, EVE_num_samples_log_proba=case["n_samples"],
                             EVE_model_parameters_location=inp["params_file"], MSA_recalibrate_probas=case["msa_recal"],
                             EVE_recalibrate_probas=case["eve_recal"], precision=precision, max_rows=32768)


def _same_prior(mine, want, tol):
    mine, want = np.asarray(mine), np.asarray(want)
    assert np.array_equal(np.isfinite(mine), np.isfinite(want))
    assert np.abs(np.nan_to_num(mine, neginf=0) - np.nan_to_num(want, neginf=0)).max() < tol


@pytest.mark.parametrize("name", list(CASES))
def test_trancepteve_constructor_recalibration_and_scores(name, tmp_path):
    case, meta, gd, inp, ck, paths = _setup(name, tmp_path)
    sc = _scorer(case, meta, inp, ck, paths)
    assert (sc.MSA_processed_depth, sc.EVE_processed_depth) == (meta["MSA_processed_depth"], meta["EVE_processed_depth"])
    assert (sc.retrieval_inference_MSA_weight, sc.retrieval_inference_EVE_weight) == \
        (meta["retrieval_inference_MSA_weight"], meta["retrieval_inference_EVE_weight"])
    _same_prior(sc.MSA_log_prior, np.load(os.path.join(gd, "msa_log_prior_init.npy")), 1e-5)   # pg_msa_prior (fp64) + weights by name
    if case["kind"] == "TranceptEVE":
        _same_prior(sc.EVE_log_prior, np.load(os.path.join(gd, "eve_log_prior_init.npy")), 1e-6)  # ensemble mean of the caches
    rows, labels = sc.get_transformer_log_softmax(meta["target_seq"])                          # out_logprobs of the fused call
    assert list(labels) == meta["wt_shift_labels"]
    assert np.abs(rows.numpy() - np.load(os.path.join(gd, "wt_log_softmax_msa_fused.npy"))).max() < TOL
    got = sc.score_mutants(inp["dms"], meta["target_seq"])
    ref = pd.read_csv(os.path.join(gd, "reference_scores.csv"))
    assert list(got.columns) == list(ref.columns) and list(got["mutated_sequence"]) == list(ref["mutated_sequence"])
    assert list(got["mutant"]) == list(ref["mutant"])
    for c in ("avg_score_L_to_R", "avg_score_R_to_L", "avg_score"):
        assert np.abs(got[c].to_numpy(dtype=np.float64) - ref[c].to_numpy(dtype=np.float64)).max() < TOL, c
    _same_prior(sc.MSA_log_prior, np.load(os.path.join(gd, "msa_log_prior_final.npy")), 5e-3)   # after recalibration
    if case["kind"] == "TranceptEVE":
        _same_prior(sc.EVE_log_prior, np.load(os.path.join(gd, "eve_log_prior_final.npy")), 5e-3)
    sc.close()


def test_fast_precision_keeps_ranking(tmp_path):
    from scipy.stats import spearmanr
    case, meta, gd, inp, ck, paths = _setup("trancepteve_subs", tmp_path)
    sc = _scorer(case, meta, inp, ck, paths, precision="f16")
    got = sc.score_mutants(inp["dms"], meta["target_seq"])
    sc.close()
    ref = pd.read_csv(os.path.join(gd, "reference_scores.csv"))
    assert np.abs(got["avg_score"].values - ref["avg_score"].values).max() < 2e-2
    assert spearmanr(got["avg_score"].values, ref["avg_score"].values).correlation > 0.999


@pytest.mark.parametrize("name", ["trancepteve_subs", "trancepteve_nonfocus"])
def test_sequence_weights_computed_on_gpu_match_reference(name, tmp_path):
    """No readable weights file -> MSAProcessing computes 1/|cluster| with pg_msa_cluster_neighbors and saves it; the reference's
    numpy loop wrote tests/golden/<case>/msa_weights.npy (same sequence order)."""
    from proteingym_b200.msa_processing import MSAProcessing
    case = CASES[name]
    inp = make_inputs(case, str(tmp_path))
    open(inp["weights_file"], "wb").close()
    m = MSAProcessing(inp["msa_file"], weights_location=inp["weights_file"], threshold_sequence_frac_gaps=case["seq_thr"],
                      threshold_focus_cols_frac_gaps=case["col_thr"])
    want = np.load(os.path.join(GOLDEN, name, "msa_weights.npy"))
    assert m.weights.shape == want.shape and np.abs(m.weights - want).max() < 1e-12
    assert np.array_equal(np.load(inp["weights_file"]), m.weights) and abs(m.Neff - want.sum()) < 1e-9


def test_eve_prior_on_device_matches_host_arithmetic(tmp_path):
    """The sampler's tensor algebra on the GPU against the same function on the CPU, with the posterior variances switched off
    (log-variance -80) so the two generators' different streams cannot matter; plus the stochastic run's support and scale."""
    L = 23
    focus = list(synth.random_protein(L, 3))
    cols = list(range(L))
    st0 = synth.make_eve_state(L, seed=7, log_var=-80.0)
    st0["encoder.fc_log_var.bias"] = torch.full_like(st0["encoder.fc_log_var.bias"], -80.0)
    st0["encoder.fc_log_var.weight"] = torch.zeros_like(st0["encoder.fc_log_var.weight"])
    a = eve_prior.eve_log_prior_single(st0, synth.EVE_TINY_PARAMS, focus, cols, L + 4, 2, 3, device="cpu")
    b = eve_prior.eve_log_prior_single(st0, synth.EVE_TINY_PARAMS, focus, cols, L + 4, 2, 3, device="cuda").cpu()
    assert torch.equal(torch.isfinite(a), torch.isfinite(b))
    fin = torch.isfinite(a)
    assert (a[fin] - b[fin]).abs().max() < 1e-4
    assert torch.isinf(b[:2]).all() and torch.isinf(b[2 + L:]).all() and torch.isinf(b[:, :5]).all()
    st1 = synth.make_eve_state(L, seed=7)
    c = eve_prior.eve_log_prior_single(st1, synth.EVE_TINY_PARAMS, focus, cols, L + 4, 2, 3000, device="cuda").cpu()
    d = eve_prior.eve_log_prior_single(st1, synth.EVE_TINY_PARAMS, focus, cols, L + 4, 2, 3000, device="cpu")
    assert torch.equal(torch.isfinite(c), torch.isfinite(d)) and (c[fin] - d[fin]).abs().mean() < 0.2  # Monte-Carlo agreement (3000 draws; two streams differ by ~0.08)


def test_eve_dense_products_run_on_the_library_kernels():
    """eve_prior._Gemm (pg_pack_weight fmt 1 + pg_gemm nseg 3 with the fp32 reduce-add epilogue) and pg_eve_output_conv against fp64,
    on the awkward shapes of the EVE decoder: K not a multiple of 64, one-row products, variance-sized magnitudes."""
    from proteingym_b200 import _lib
    lib = _lib.load()
    n0 = lib.pg_launch_count()
    g = torch.Generator(device="cuda").manual_seed(5)
    mm = eve_prior._Gemm()
    for (M, K, N, mag) in [(1, 50, 300, 1.0), (37, 300, 1000, 1.0), (200, 2000, 460, 30.0), (64, 40, 20, 1e-6), (3, 4000, 500, 1e-5)]:
        x = torch.randn(M, K, device="cuda", generator=g) * mag
        w = torch.randn(N, K, device="cuda", generator=g) * (mag if mag < 1 else 1.0) / K ** 0.5
        got = mm(x, w, key=f"w{K}x{N}")
        again = mm(x, w, key=f"w{K}x{N}")   # second call: packed weight from the cache
        ref = x.double() @ w.double().T
        scale = (x.double().abs() @ w.double().abs().T).max().item()
        assert torch.equal(got, again)
        assert (got.double() - ref).abs().max().item() < 2e-6 * scale, (M, K, N, mag)
    S, J, A, Cd = 33, 100, 20, 40
    x = torch.randn(S, J * A, device="cuda", generator=g)
    conv = torch.randn(S, Cd, A, device="cuda", generator=g)
    y = eve_prior._output_conv(x, conv, J, A, Cd)
    ref = torch.einsum("sja,sca->sjc", x.double().reshape(S, J, A), conv.double()).reshape(S, J * Cd)
    assert (y.double() - ref).abs().max().item() < 1e-5
    assert lib.pg_launch_count() > n0   # the library's kernels did the work


def test_eve_local_sampler_on_device_matches_host_arithmetic():
    """The batched (local-reparameterisation) sampler through the library's kernels against the same function with CPU tensors:
    variances switched off, three batches (the packed-weight cache is reused), then a stochastic run's scale."""
    import copy
    P = copy.deepcopy(synth.EVE_TINY_PARAMS)
    P["decoder_parameters"]["hidden_layers_sizes"] = [24, 32, 60]   # alphabet | last hidden size
    L = 11
    focus, cols = list(synth.random_protein(L, 3)), list(range(L))
    st0 = synth.make_eve_state(L, P, seed=7, log_var=-80.0)
    st0["encoder.fc_log_var.bias"].fill_(-80.0)
    st0["encoder.fc_log_var.weight"].zero_()
    a = eve_prior.eve_log_prior_single(st0, P, focus, cols, L, 0, 70, device="cpu", sampler="local", batch=32)
    b = eve_prior.eve_log_prior_single(st0, P, focus, cols, L, 0, 70, device="cuda", sampler="local", batch=32).cpu()
    fin = torch.isfinite(a)
    assert torch.equal(fin, torch.isfinite(b)) and (a[fin] - b[fin]).abs().max() < 1e-4
    st = synth.make_eve_state(L, P, seed=7, log_var=-3.0)
    c = eve_prior.eve_log_prior_single(st, P, focus, cols, L, 0, 4000, device="cuda", sampler="local").cpu()
    d = eve_prior.eve_log_prior_single(st, P, focus, cols, L, 0, 4000, device="cpu", sampler="local")
    assert torch.equal(torch.isfinite(c), torch.isfinite(d)) and (c[fin] - d[fin]).abs().mean() < 0.2


def test_cli_end_to_end(tmp_path, monkeypatch):
    """score_trancepteve.py drop-in, manual-fields mode: CSV columns / rows / values as the reference's, coefficient log appended."""
    from proteingym_b200 import score_trancepteve
    name = "trancepteve_subs"
    case, meta, gd, inp, ck, paths = _setup(name, tmp_path)
    dms_dir, out_dir = tmp_path / "dms", tmp_path / "out"
    dms_dir.mkdir()
    inp["dms"].to_csv(dms_dir / "ASSAY1.csv", index=False)
    monkeypatch.chdir(tmp_path)
    score_trancepteve.main([
        "--checkpoint", ck, "--target_seq", meta["target_seq"], "--DMS_file_name", "ASSAY1.csv", "--DMS_data_folder", str(dms_dir),
        "--output_scores_folder", str(out_dir), "--inference_time_retrieval_type", "TranceptEVE", "--MSA_folder", str(tmp_path),
        "--MSA_filename", "TARGET_msa.a2m", "--MSA_weights_folder", str(tmp_path), "--MSA_weight_file_name", "TARGET_weights.npy",
        "--MSA_start", str(case["msa"][0] + 1), "--MSA_end", str(case["msa"][1]), "--MSA_threshold_sequence_frac_gaps", str(case["seq_thr"]),
        "--MSA_threshold_focus_cols_frac_gaps", str(case["col_thr"]), "--EVE_model_folder", inp["eve_dir"], "--EVE_seeds", "0",
        "--EVE_num_samples_log_proba", str(case["n_samples"]), "--EVE_model_parameters_location", inp["params_file"],
        "--EVE_recalibrate_probas"])
    got = pd.read_csv(out_dir / "ASSAY1.csv")
    ref = pd.read_csv(os.path.join(gd, "reference_scores.csv"))
    assert list(got.columns) == list(ref.columns) and list(got["mutant"]) == list(ref["mutant"])
    assert np.abs(got["avg_score"].values - ref["avg_score"].values).max() < TOL
    log = (tmp_path / "TranceptEVE_aggregation_coefficients_log").read_text().splitlines()
    assert log[0].startswith("DMS_id,num_mutants_scored") and log[1].split(",")[:2] == ["ASSAY1", str(len(ref))]
    assert log[1].split(",")[-2:] == [str(meta["retrieval_inference_MSA_weight"]), str(meta["retrieval_inference_EVE_weight"])]


def test_tranception_cli_retrieval_with_sequence_weights(tmp_path):
    """score_tranception_proteingym.py --inference_time_retrieval with an EVE weights file, against the reference's real
    TranceptionLMHeadModel (tests/golden/tranception_retrieval): weighted MSA prior on the GPU + alpha = 0.6 fusion on all columns."""
    from proteingym_b200 import score_tranception_proteingym
    gd = os.path.join(GOLDEN, "tranception_retrieval")
    meta = json.load(open(os.path.join(gd, "meta.json")))
    case = meta["case"]
    inp = make_inputs(case, str(tmp_path), weights_file=os.path.join(gd, "msa_weights.npy"))
    ck = str(tmp_path / "ckpt")
    synth.write_tranception_checkpoint(ck, inp["arch"], meta["tranception_seed"])
    dms_dir, out_dir = tmp_path / "dms", tmp_path / "out"
    dms_dir.mkdir()
    inp["dms"].to_csv(dms_dir / "ASSAY2.csv", index=False)
    score_tranception_proteingym.main([
        "--checkpoint", ck, "--target_seq", meta["target_seq"], "--DMS_file_name", "ASSAY2.csv", "--DMS_data_folder", str(dms_dir),
        "--output_scores_folder", str(out_dir), "--inference_time_retrieval", "--MSA_folder", str(tmp_path), "--MSA_filename", "TARGET_msa.a2m",
        "--MSA_weights_folder", str(tmp_path), "--MSA_weight_file_name", "TARGET_weights.npy", "--MSA_start", str(case["msa"][0] + 1),
        "--MSA_end", str(case["msa"][1])])
    got = pd.read_csv(out_dir / "ASSAY2.csv")
    ref = pd.read_csv(os.path.join(gd, "reference_scores.csv"))
    assert list(got.columns) == list(ref.columns) and list(got["mutated_sequence"]) == list(ref["mutated_sequence"])
    for c in ("avg_score_L_to_R", "avg_score_R_to_L", "avg_score"):
        assert np.abs(got[c].values - ref[c].values).max() < TOL, c


def test_true_size_trancepteve_fusion_properties():
    """Tranception-L architecture (36 x 1280, 20 heads) with synthetic [L, 25] MSA / EVE log priors. No CPU
    oracle at this size; the fused head is checked through properties: (1) zero weights reproduce the plain scores bit for bit,
    (2) every emitted row is a normalised distribution and the label terms add up to the returned sum, (3) the fused sums equal the
    three-way mixture evaluated on the host from the unfused rows and the prior-row indices — left-to-right and flipped, window
    partially overlapping the MSA range, non-focus rows included."""
    from proteingym_b200.tranception_engine import TranceptionScorer, prior_rows, tokenize
    arch = synth.TRANCEPTION_L
    st = synth.make_tranception_state(arch, 0)
    cfg = {"n_embd": arch.embed_dim, "n_head": arch.heads, "n_layer": arch.layers, "n_ctx": arch.n_ctx, "n_inner": arch.ffn_dim,
           "vocab_size": arch.vocab, "layer_norm_epsilon": arch.ln_eps, "activation_function": "squared_relu"}
    sc = TranceptionScorer(cfg, {k[len("transformer."):]: v for k, v in st.items() if k.startswith("transformer.")}, max_rows=16384)
    rng = np.random.RandomState(0)
    Lfull = 400
    full = synth.random_protein(Lfull, 4)
    msa_lp = np.log(rng.dirichlet(np.ones(25), size=Lfull)).astype(np.float32)
    eve_lp = np.full((Lfull, 25), -np.inf, dtype=np.float32)
    eve_lp[:, 5:] = np.log(rng.dirichlet(np.ones(20), size=Lfull)).astype(np.float32)
    nonfocus_rows = [70, 71, 150, 333]
    eve_lp[nonfocus_rows, 5:] = -np.inf
    windows = [(0, 300), (100, 400), (50, 200)]
    seqs = [full[a:b] for a, b in windows]
    msa_start, msa_end, alpha, beta = 60, 350, 0.3, 0.6
    plain, rows = sc.sequence_logprobs(seqs, return_rows=True)
    for r, s in zip(rows, seqs):
        assert r.shape == (len(s) + 1, 25) and np.abs(np.log(np.exp(r.astype(np.float64)).sum(-1))).max() < 1e-4
        lab = np.asarray(tokenize(s)[1:])
        assert abs(r[np.arange(len(lab)), lab].astype(np.float64).sum() - plain[seqs.index(s)]) < 2e-3
    kw = dict(windows=windows, prior=msa_lp, msa_start=msa_start, msa_end=msa_end, prior2=eve_lp, first_col=5, nonfocus_fallback=True)
    zero = sc.sequence_logprobs(seqs, alpha=0.0, beta=0.0, **{**kw, "prior2": np.where(np.isfinite(eve_lp), eve_lp, 0).astype(np.float32)})
    assert np.array_equal(zero, plain)
    for flip in (False, True):
        strings = [s[::-1] for s in seqs] if flip else seqs
        base, rws = sc.sequence_logprobs(strings, return_rows=True)
        got = sc.sequence_logprobs(strings, alpha=alpha, beta=beta, flip=flip, **kw)
        nonfocus = eve_lp[:, 5:].min(axis=1) == -np.inf
        for k, s in enumerate(strings):
            T = len(s) + 2
            p1, p2 = np.full(T, -1, np.int32), np.full(T, -1, np.int32)
            prior_rows(p1, p2, windows[k][0], windows[k][1], msa_start, msa_end, flip, nonfocus)
            lab = np.asarray(tokenize(s)[1:])
            want = 0.0
            for t, v in enumerate(lab):
                lp = np.float32(rws[k][t, v])
                if v >= 5 and p1[t] >= 0:
                    m = np.float32(1 - alpha) * lp + np.float32(alpha) * msa_lp[p1[t], v]
                    if p2[t] >= 0:
                        m = np.float32(1 - beta) * m + np.float32(beta) * eve_lp[p2[t], v]
                    lp = m
                elif v >= 5 and p1[t] == -2:
                    lp = np.float32(1 - alpha) * lp
                want += float(lp)
            assert np.isfinite(got[k]) and abs(got[k] - want) < 2e-3 * max(1.0, abs(want) / 100), (flip, k, got[k], want)
        assert (np.abs(got - base) > 1.0).all()  # the priors really moved the scores
    sc.close()
