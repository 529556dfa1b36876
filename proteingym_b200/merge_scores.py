"""Merge per-assay score files into one frame per assay and compute the leaderboard's rank metric — the two steps that follow the
scorers in ProteinGym's pipeline:

  * ``merge``   restates proteingym/merge.py:47-114 for the models of this repo: per assay, left-join every model's score column
                (config.json: ``location`` / ``input_score_name`` / ``key`` / ``directionality``) onto the DMS file, after the
                reference's de-duplication (drop_duplicates + mean per key) and with its skip rules (no overlap / proper subset /
                changed length -> the model is left out of that assay with the same warning text).
  * ``metrics`` Spearman of every score column against ``DMS_score`` per assay (performance_DMS_benchmarks.py:212,
                ``scipy.stats.spearmanr``) and, when a folder of reference score files is given, the parity numbers north_star
                asks for per assay: Spearman and max |difference| between this repo's scores and the reference's.

  python -m proteingym_b200.merge_scores --DMS_reference_file map.csv --DMS_assays_location dms/ --model_scores_location scores/ \\
         [--models ESM1v_single ESM2_3B ...] [--config_file config.json] [--reference_scores_location ref_scores/]

writes ``<model_scores_location>/<merged_scores_dir>/<DMS_id>.csv`` and ``_metrics.csv`` next to them; exits non-zero when
``--min_parity_spearman`` / ``--max_parity_abs`` are given and an assay misses them. CPU-side bookkeeping by design: one argsort per
column per assay (the GPU work is upstream)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import pandas as pd

# config.json entries (reference config.json, model_list_zero_shot_substitutions_DMS / _indels_DMS) of the models this repo scores
DEFAULT_MODELS = {
    "substitutions": {
        "ESM1v_single": {"input_score_name": "esm1v_t33_650M_UR90S_1", "location": "ESM1v", "directionality": 1, "key": "mutant"},
        "ESM1v_ensemble": {"input_score_name": "Ensemble_ESM1v", "location": "ESM1v", "directionality": 1, "key": "mutant"},
        "ESM1b": {"input_score_name": "esm1b_t33_650M_UR50S", "location": "ESM1b", "directionality": 1, "key": "mutant"},
        "ESM2_650M": {"input_score_name": "esm2_t33_650M_UR50D", "location": "ESM2/650M", "directionality": 1, "key": "mutant"},
        "ESM2_3B": {"input_score_name": "esm2_t36_3B_UR50D", "location": "ESM2/3B", "directionality": 1, "key": "mutant"},
        "Tranception_L": {"input_score_name": "avg_score", "location": "Tranception/Tranception_L", "directionality": 1, "key": "mutated_sequence"},
        "TranceptEVE_L": {"input_score_name": "avg_score", "location": "TranceptEVE/TranceptEVE_L", "directionality": 1, "key": "mutant"},
    },
    "indels": {
        "Tranception_L": {"input_score_name": "avg_score", "location": "Tranception/Tranception_L", "directionality": 1, "key": "mutated_sequence"},
        "TranceptEVE_L": {"input_score_name": "avg_score", "location": "TranceptEVE/TranceptEVE_L", "directionality": 1, "key": "mutated_sequence"},
    },
}


def load_models(config_file, mutation_type, dataset="DMS", names=None):
    if config_file:
        with open(config_file) as fh:
            config = json.load(fh)
        field = f"model_list_zero_shot_{mutation_type}_{'DMS' if dataset == 'DMS' else 'clinical'}"
        models = config[field]
    else:
        models = DEFAULT_MODELS[mutation_type]
    if names:
        missing = [n for n in names if n not in models]
        if missing:
            raise KeyError(f"models not in the configuration: {missing}")
        models = {n: models[n] for n in names}
    return models


def merge_assay(DMS_file: pd.DataFrame, DMS_id: str, models: dict, model_scores_location: str, mutation_type: str, log=print):
    """One assay of merge.py:58-102 -> merged frame (DMS columns + one column per merged model)."""
    if "mutated_sequence" not in DMS_file:
        DMS_file = DMS_file.assign(mutated_sequence=DMS_file["mutant"])
    merged = DMS_file
    n0 = len(merged)
    for model, spec in models.items():
        key = spec["key"]
        dms_col = key if mutation_type == "substitutions" else "mutated_sequence"
        path = os.path.join(model_scores_location, spec["location"], DMS_id + ".csv")
        if not os.path.exists(path):
            log(f"Warning: no score file for {DMS_id} with model {model} ({path}). Skipping")
            continue
        sf = pd.read_csv(path)
        if "sequence" in sf:
            sf["mutated_sequence"] = sf["sequence"]
        sf[model] = spec["directionality"] * sf[spec["input_score_name"]]
        sf = sf[[key, model]].drop_duplicates().groupby(key).mean().reset_index()
        have, want = set(sf[key]), set(merged[dms_col])
        if not (have & want):
            log("Warning: No overlap on mutants for {} with model {}. Skipping".format(DMS_id, model))
            continue
        if have < want:
            log("WARNING: {} and {} do not have the same mutants. Skipping.".format(model, DMS_id))
            continue
        out = pd.merge(merged, sf.rename(columns={key: dms_col}), on=dms_col, how="left")
        if len(out) != n0:
            log("WARNING: Merge on {} for {} changed length. mutant_merge_keys are likely different between them.".format(model, DMS_id))
            continue
        merged = out
    return merged


def rank(a: np.ndarray) -> np.ndarray:
    """Average ranks (ties share the mean rank), what scipy.stats.rankdata / spearmanr use."""
    a = np.asarray(a, dtype=np.float64)
    order = np.argsort(a, kind="mergesort")
    r = np.empty(len(a), dtype=np.float64)
    sa = a[order]
    bounds = np.flatnonzero(np.concatenate([[True], sa[1:] != sa[:-1], [True]]))
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        r[order[lo:hi]] = 0.5 * (lo + hi - 1) + 1.0
    return r


def spearman(x, y) -> float:
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    if len(x) < 2 or np.isnan(x).any() or np.isnan(y).any():
        return float("nan")  # spearmanr's default nan_policy propagates
    rx, ry = rank(x), rank(y)
    rx -= rx.mean()
    ry -= ry.mean()
    den = np.sqrt((rx * rx).sum() * (ry * ry).sum())
    return float((rx * ry).sum() / den) if den > 0 else float("nan")


def assay_metrics(merged: pd.DataFrame, DMS_id: str, score_cols, reference: pd.DataFrame | None = None, ref_cols=None, key="mutant"):
    rows = []
    for c in score_cols:
        if c not in merged:
            continue
        rec = {"DMS_id": DMS_id, "model": c, "n": len(merged),
               "spearman_vs_DMS": spearman(merged["DMS_score"], merged[c]) if "DMS_score" in merged else float("nan")}
        rc = (ref_cols or {}).get(c, c)
        if reference is not None and rc in reference:
            k = key if key in merged and key in reference else "mutated_sequence"
            j = pd.merge(merged[[k, c]], reference[[k, rc]].rename(columns={rc: "_ref"}).drop_duplicates(k), on=k, how="inner")
            rec.update(n_compared=len(j), parity_spearman=spearman(j[c], j["_ref"]),
                       parity_max_abs=float(np.abs(j[c].to_numpy(np.float64) - j["_ref"].to_numpy(np.float64)).max()) if len(j) else float("nan"))
        rows.append(rec)
    return rows


def main(argv=None):
    ap = argparse.ArgumentParser(description="merge per-assay score files and compute Spearman (DMS and parity vs reference scores)")
    ap.add_argument("--DMS_assays_location", required=True)
    ap.add_argument("--model_scores_location", required=True)
    ap.add_argument("--merged_scores_dir", default="merged_scores")
    ap.add_argument("--mutation_type", default="substitutions", choices=["substitutions", "indels"])
    ap.add_argument("--dataset", default="DMS", choices=["DMS", "clinical"])
    ap.add_argument("--DMS_reference_file", required=True)
    ap.add_argument("--config_file", default=None, help="the reference's config.json; default: the built-in entries of this repo's models")
    ap.add_argument("--models", nargs="*", default=None, help="subset of model names (default: all configured models)")
    ap.add_argument("--reference_scores_location", default=None,
                    help="folder laid out like --model_scores_location holding the REFERENCE implementation's score files: adds parity columns")
    ap.add_argument("--min_parity_spearman", type=float, default=None)
    ap.add_argument("--max_parity_abs", type=float, default=None)
    a = ap.parse_args(argv)
    ref_file = pd.read_csv(a.DMS_reference_file)
    models = load_models(a.config_file, a.mutation_type, a.dataset, a.models)
    out_dir = os.path.join(a.model_scores_location, a.merged_scores_dir)
    os.makedirs(out_dir, exist_ok=True)
    rows = []
    for DMS_id, DMS_filename in zip(ref_file["DMS_id"], ref_file["DMS_filename"]):
        path = os.path.join(a.DMS_assays_location, DMS_filename)
        if not os.path.exists(path):
            print("Could not find DMS file {}. Skipping.".format(path))
            continue
        merged = merge_assay(pd.read_csv(path), DMS_id, models, a.model_scores_location, a.mutation_type)
        if "DMS_total_number_mutants" in ref_file:
            expected = ref_file.loc[ref_file["DMS_id"] == DMS_id, "DMS_total_number_mutants"].values[0]
            if len(merged) != expected:
                print(f"Warning: Insufficient mutants for {DMS_id}: {len(merged)}, expected {expected}.")
        merged.to_csv(os.path.join(out_dir, f"{DMS_id}.csv"), index=False)
        reference = None
        if a.reference_scores_location:
            reference = merge_assay(pd.read_csv(path), DMS_id, models, a.reference_scores_location, a.mutation_type, log=lambda *_: None)
        key = "mutant" if a.mutation_type == "substitutions" and "mutant" in merged else "mutated_sequence"
        rows += assay_metrics(merged, DMS_id, list(models), reference, key=key)
    rep = pd.DataFrame(rows)
    rep.to_csv(os.path.join(out_dir, "_metrics.csv"), index=False)
    if len(rep):
        print(rep.groupby("model")[[c for c in ("spearman_vs_DMS", "parity_spearman", "parity_max_abs") if c in rep]].agg(["mean", "min", "max"]).to_string())
    bad = []
    if a.min_parity_spearman is not None and "parity_spearman" in rep:
        bad += list(rep.loc[~(rep["parity_spearman"] >= a.min_parity_spearman), "DMS_id"])
    if a.max_parity_abs is not None and "parity_max_abs" in rep:
        bad += list(rep.loc[~(rep["parity_max_abs"] <= a.max_parity_abs), "DMS_id"])
    if bad:
        print("parity bar missed on:", sorted(set(bad)))
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
