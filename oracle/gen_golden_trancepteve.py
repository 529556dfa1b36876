"""ORACLE SUPPORT — TEST INFRASTRUCTURE ONLY. Build-container script (needs /root/reference).

Golden vectors for the retrieval / TranceptEVE rows from the reference's UNMODIFIED
``TrancepteveLMHeadModel`` (see oracle/ref_shims_trancepteve.py for the three compatibility patches): constructor (MSA prior with
MSA_processing weights, EVE VAE log prior, depth-based aggregation weights), recalibration, fused forward and score_mutants.
Inputs (Tranception weights, MSA, EVE checkpoint, DMS) are regenerated from seeds by proteingym_b200.synth; what is stored under
tests/golden/trancepteve_<case>/ is the reference's outputs plus the sequence-weight vector it computed."""
from __future__ import annotations

import json
import os
import pickle
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402
import torch  # noqa: E402

from oracle import ref_shims_trancepteve as RE  # noqa: E402
from oracle import ref_shims_tranception as RT  # noqa: E402
from proteingym_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, os.path.join(ROOT, "tests"))
from trancepteve_cases import CASES, make_inputs  # noqa: E402  (shared with the tests: seeded inputs of each case)


def run_case(name: str, case: dict):
    mp, cfgm = RE.install()
    out = os.path.join(GOLD, name)
    os.makedirs(out, exist_ok=True)
    work = tempfile.mkdtemp(prefix="pg_te_")
    inp = make_inputs(case, work)
    arch, seq = inp["arch"], inp["seq"]
    open(inp["weights_file"], "wb").close()  # exists but unreadable: MSA_processing computes the weights and saves them here
    cfg = cfgm.TranceptEVEConfig(
        vocab_size=arch.vocab, n_positions=arch.n_ctx, n_embd=arch.embed_dim, n_layer=arch.layers, n_head=arch.heads, n_inner=arch.ffn_dim,
        activation_function="squared_relu", resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0, layer_norm_epsilon=arch.ln_eps,
        tokenizer=RT.tokenizer(), full_target_seq=seq, scoring_window="optimal", inference_time_retrieval_type=case["kind"],
        retrieval_aggregation_mode="aggregate_substitution", retrieval_weights_manual=False, MSA_filename=inp["msa_file"],
        MSA_weight_file_name=inp["weights_file"], MSA_start=case["msa"][0], MSA_end=case["msa"][1],
        MSA_threshold_sequence_frac_gaps=case["seq_thr"], MSA_threshold_focus_cols_frac_gaps=case["col_thr"],
        EVE_num_samples_log_proba=case["n_samples"], EVE_model_parameters_location=inp["params_file"],
        MSA_recalibrate_probas=case["msa_recal"], EVE_recalibrate_probas=case["eve_recal"])
    cfg.n_ctx = arch.n_ctx
    # EVE checkpoints need the focus-column count: run the reference's MSA_processing once to get it (this also writes the weights)
    from trancepteve.utils import msa_utils
    pre = msa_utils.MSA_processing(MSA_location=inp["msa_file"], use_weights=True, threshold_sequence_frac_gaps=case["seq_thr"],
                                   threshold_focus_cols_frac_gaps=case["col_thr"], weights_location=inp["weights_file"])
    # One weights file per assay, as in production: written by the EVE-style processing (focus-column threshold of the case). The
    # constructor's get_msa_prior re-reads it with all columns kept and assigns weights[i] to its i-th retained sequence.
    shutil.copy(inp["weights_file"], os.path.join(out, "msa_weights.npy"))
    paths = []
    for sd in case["eve_seeds"]:
        pth = os.path.join(inp["eve_dir"], f"TARGET_msa_seed_{sd}")
        torch.save({"model_state_dict": synth.make_eve_state(pre.seq_len, seed=100 + sd)}, pth)
        paths.append(pth)
    cfg.EVE_model_paths = paths
    # RNG harness: on a GPU the reference draws its EVE samples from the CUDA generator, which VAE_model.__init__ seeds (42) and
    # module construction (CPU) never touches. Here everything is CPU, so re-seed at the entry of the (otherwise unmodified)
    # sampling method to put the generator in that same just-seeded state.
    if not getattr(mp.TrancepteveLMHeadModel, "_pg_seeded", False):
        orig = mp.TrancepteveLMHeadModel.get_EVE_log_prior_single

        def seeded(self, *a, **k):
            torch.manual_seed(42)
            return orig(self, *a, **k)

        mp.TrancepteveLMHeadModel.get_EVE_log_prior_single = seeded
        mp.TrancepteveLMHeadModel._pg_seeded = True
    with RE.light_pretrained_init():
        model = mp.TrancepteveLMHeadModel(cfg)
    missing, unexpected = model.load_state_dict(synth.make_tranception_state(arch, 5), strict=False)
    assert not unexpected and all(k.endswith(("attn.bias", "attn.masked_bias", ".alibi")) for k in missing), (missing, unexpected)
    model.eval()
    meta = {"name": name, "case": case, "target_seq": seq, "focus_cols": [int(c) for c in pre.focus_cols], "focus_seq_len": int(pre.seq_len),
            "MSA_processed_depth": int(model.MSA_processed_depth), "EVE_processed_depth": int(model.EVE_processed_depth),
            "retrieval_inference_MSA_weight": float(model.retrieval_inference_MSA_weight),
            "retrieval_inference_EVE_weight": float(model.retrieval_inference_EVE_weight), "torch": torch.__version__, "tranception_seed": 5}
    np.save(os.path.join(out, "msa_log_prior_init.npy"), model.MSA_log_prior.numpy())
    if case["kind"] == "TranceptEVE":
        np.save(os.path.join(out, "eve_log_prior_init.npy"), model.EVE_log_prior.numpy())
        for i, pth in enumerate(paths):  # per-model cache files the reference wrote (pickled tensors)
            c = os.path.join(os.path.dirname(pth), "log_prior", "_".join([os.path.basename(pth), str(case["n_samples"]), "log_space"]))
            with open(c, "rb") as fh:
                np.save(os.path.join(out, f"eve_log_prior_model{i}.npy"), pickle.load(fh).numpy())
    with torch.no_grad():
        lr, lab = model.get_transformer_log_softmax(sequence=seq)  # default retrieval type "Tranception": MSA fusion only
        np.save(os.path.join(out, "wt_log_softmax_msa_fused.npy"), lr.numpy())
        meta["wt_shift_labels"] = [int(x) for x in lab]
        scores = model.score_mutants(DMS_data=inp["dms"], target_seq=seq, scoring_mirror=True, batch_size_inference=20, num_workers=0,
                                     indel_mode=False)
    np.save(os.path.join(out, "msa_log_prior_final.npy"), model.MSA_log_prior.numpy())
    if case["kind"] == "TranceptEVE":
        np.save(os.path.join(out, "eve_log_prior_final.npy"), model.EVE_log_prior.numpy())
    scores.to_csv(os.path.join(out, "reference_scores.csv"), index=False)
    with open(os.path.join(out, "meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    shutil.rmtree(work, ignore_errors=True)
    print(f"[gen_golden_trancepteve] {name}: {len(scores)} rows, depths MSA {meta['MSA_processed_depth']} EVE {meta['EVE_processed_depth']}, "
          f"alpha {meta['retrieval_inference_MSA_weight']} beta {meta['retrieval_inference_EVE_weight']}", flush=True)


def build_tranception(arch, seq, retrieval=None):
    """The reference's real ``tranception.model_pytorch.TranceptionLMHeadModel`` (same three compatibility patches)."""
    mp, _ = RT.install()
    from tranception import config as tcfg
    mp.TranceptionLMHeadModel.init_weights = lambda self: None
    mp.TranceptionModel.init_weights = lambda self: None
    mp.TranceptionModel.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n
    RE._positional_series_fallback()
    cfg = tcfg.TranceptionConfig(
        vocab_size=arch.vocab, n_positions=arch.n_ctx, n_embd=arch.embed_dim, n_layer=arch.layers, n_head=arch.heads, n_inner=arch.ffn_dim,
        activation_function="squared_relu", resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0, layer_norm_epsilon=arch.ln_eps,
        tokenizer=RT.tokenizer(), scoring_window="optimal", retrieval_aggregation_mode=None)
    cfg.n_ctx = arch.n_ctx
    if retrieval:
        cfg.retrieval_aggregation_mode = "aggregate_substitution"
        cfg.full_protein_length = len(seq)
        for k, v in retrieval.items():
            setattr(cfg, k, v)
    with RE.light_pretrained_init():
        model = mp.TranceptionLMHeadModel(cfg)
    missing, unexpected = model.load_state_dict(synth.make_tranception_state(arch, 5), strict=False)
    assert not unexpected and all(k.endswith(("attn.bias", "attn.masked_bias", ".alibi")) for k in missing), (missing, unexpected)
    return model.eval()


def run_tranception_cases():
    """(1) Cross-check: the real TranceptionLMHeadModel reproduces the stored tranception_subs scores (written earlier through the
    hybrid wrapper of gen_golden_tranception.py). (2) New case tranception_retrieval: the real class with inference-time retrieval
    (weighted MSA prior, alpha = 0.6 on all columns, model_pytorch.py:806-830)."""
    meta = json.load(open(os.path.join(GOLD, "tranception_subs_meta.json")))
    arch = synth.TranceptionArch(**meta["arch"])
    dms = pd.read_csv(os.path.join(GOLD, "tranception_subs_dms.csv"))
    model = build_tranception(arch, meta["target_seq"])
    with torch.no_grad():
        got = model.score_mutants(DMS_data=dms, target_seq=meta["target_seq"], scoring_mirror=True, batch_size_inference=20, num_workers=0)
    ref = pd.read_csv(os.path.join(GOLD, "tranception_subs_reference_scores.csv"))
    err = np.abs(got["avg_score"].values - ref["avg_score"].values).max()
    assert list(got["mutated_sequence"]) == list(ref["mutated_sequence"]) and err < 1e-6, err
    print(f"[gen_golden_trancepteve] real TranceptionLMHeadModel == stored tranception_subs golden (max diff {err:.2e})")

    name = "tranception_retrieval"
    case = dict(arch=(2, 256, 4, 512, 1024), L=70, msa=(8, 62), n_msa=150, gappy=[3, 30], n_mut=80, kind="Tranception")
    out = os.path.join(GOLD, name)
    os.makedirs(out, exist_ok=True)
    work = tempfile.mkdtemp(prefix="pg_tr_")
    inp = make_inputs(case, work)
    open(inp["weights_file"], "wb").close()
    model = build_tranception(inp["arch"], inp["seq"], dict(MSA_filename=inp["msa_file"], MSA_weight_file_name=inp["weights_file"],
                                                            retrieval_inference_weight=0.6, MSA_start=case["msa"][0], MSA_end=case["msa"][1]))
    shutil.copy(inp["weights_file"], os.path.join(out, "msa_weights.npy"))
    np.save(os.path.join(out, "msa_log_prior.npy"), model.MSA_log_prior.numpy())
    with torch.no_grad():
        scores = model.score_mutants(DMS_data=inp["dms"], target_seq=inp["seq"], scoring_mirror=True, batch_size_inference=20, num_workers=0)
    scores.to_csv(os.path.join(out, "reference_scores.csv"), index=False)
    with open(os.path.join(out, "meta.json"), "w") as fh:
        json.dump({"name": name, "case": case, "target_seq": inp["seq"], "tranception_seed": 5, "retrieval_inference_weight": 0.6,
                   "torch": torch.__version__}, fh, indent=1)
    shutil.rmtree(work, ignore_errors=True)
    print(f"[gen_golden_trancepteve] {name}: {len(scores)} rows")


def run_true_size_tranception_cases():
    """BASELINE config 4 architecture at TRUE SIZE (Tranception-L: 36 x 1280, 20 heads, ffn 5120, n_ctx 1024) under the real
    TranceptionLMHeadModel: substitutions incl. multi-mutants, indels (ragged lengths), and a protein longer than n_ctx - 2 (optimal
    windows). Files use the flat tranception_* layout of gen_golden_tranception.py so the same GPU test reads them."""
    arch = synth.TRANCEPTION_L
    cases = []
    seq = synth.random_protein(180, 51)
    muts = synth.sample_mutants(seq, 60, seed=7, multi_frac=0.3)
    df = pd.DataFrame({"mutant": muts, "DMS_score": 0.0})
    df["mutated_sequence"] = [synth.apply_mutant(seq, m) for m in muts]
    cases.append(("tranception_L_subs", seq, df, False))
    seq2 = synth.random_protein(70, 52)
    ind = synth.random_indels(seq2, 40, seed=8) + [seq2]
    cases.append(("tranception_L_indels", seq2, pd.DataFrame({"mutant": ind, "mutated_sequence": ind, "DMS_score": 0.0}), True))
    seq3 = synth.random_protein(1060, 53)
    muts3 = synth.sample_mutants(seq3, 6, seed=9, multi_frac=0.34)
    df3 = pd.DataFrame({"mutant": muts3, "DMS_score": 0.0})
    df3["mutated_sequence"] = [synth.apply_mutant(seq3, m) for m in muts3]
    cases.append(("tranception_L_long", seq3, df3, False))
    for name, target, dms, indel in cases:
        if len(sys.argv) > 2 and name not in sys.argv[2:]:
            continue
        model = build_tranception(arch, target)
        t0 = time.time()
        with torch.no_grad():
            scores = model.score_mutants(DMS_data=dms, target_seq=target, scoring_mirror=True, batch_size_inference=10, num_workers=0,
                                         indel_mode=indel)
        scores.to_csv(os.path.join(GOLD, f"{name}_reference_scores.csv"), index=False)
        dms.to_csv(os.path.join(GOLD, f"{name}_dms.csv"), index=False)
        with open(os.path.join(GOLD, f"{name}_meta.json"), "w") as fh:
            json.dump({"name": name, "arch": arch.__dict__, "seed": 5, "target_seq": target, "indel_mode": indel, "scoring_window": "optimal",
                       "torch": torch.__version__, "model": "tranception.model_pytorch.TranceptionLMHeadModel (unmodified)",
                       "seconds": time.time() - t0}, fh, indent=1)
        print(f"[gen_golden_trancepteve] {name}: {len(scores)} rows in {time.time() - t0:.0f} s", flush=True)
        del model


def dump_cli_flags():
    """Option surface of the reference's two Tranception-family scripts (their parsers are built inside main(): stop main() at
    parse_args and dump the parser's actions) -> tests/golden/{tranception,trancepteve}_cli_flags.json."""
    import argparse
    import importlib.util
    RE.install()

    class Stop(Exception):
        pass

    def grab(self, *a, **k):
        raise Stop(self)

    for tag, path in (("tranception", os.path.join(RT.REF, "score_tranception_proteingym.py")),
                      ("trancepteve", os.path.join(RE.REF, "score_trancepteve.py"))):
        spec = importlib.util.spec_from_file_location("ref_cli_" + tag, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        orig = argparse.ArgumentParser.parse_args
        argparse.ArgumentParser.parse_args = grab
        try:
            mod.main()
        except Stop as e:
            parser = e.args[0]
        finally:
            argparse.ArgumentParser.parse_args = orig
        out = {}
        for a in parser._actions:
            if not a.option_strings or a.dest == "help":
                continue
            out[a.dest] = {"opts": sorted(a.option_strings), "default": str(a.default), "nargs": str(a.nargs),
                           "type": getattr(a.type, "__name__", str(a.type)), "choices": list(a.choices) if a.choices else None,
                           "const": str(a.const)}
        with open(os.path.join(GOLD, tag + "_cli_flags.json"), "w") as fh:
            json.dump(out, fh, indent=1)
        print(f"[gen_golden_trancepteve] {tag}: {len(out)} flags")


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    if len(sys.argv) > 1 and sys.argv[1] == "cli":
        return dump_cli_flags()
    if len(sys.argv) > 1 and sys.argv[1] == "tranception":
        return run_tranception_cases()
    if len(sys.argv) > 1 and sys.argv[1] == "truesize":
        return run_true_size_tranception_cases()
    for name, case in CASES.items():
        if (len(sys.argv) > 1 and name not in sys.argv[1:]) or (len(sys.argv) == 1 and case.get("true_size")):
            continue  # true-size cases only on request
        run_case(name, case)


if __name__ == "__main__":
    main()
