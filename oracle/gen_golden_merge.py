"""ORACLE SUPPORT — build-container script (needs /root/reference). Runs the UNMODIFIED reference merge step
(proteingym/merge.py::main) and its Spearman line (performance_DMS_benchmarks.py:212 = scipy.stats.spearmanr) on a small synthetic
layout and stores inputs + outputs under tests/golden/merge_case/ for tests/test_merge_cpu.py. The layout exercises the reference's
rules: duplicate rows (drop_duplicates + mean per key), a sign-flipped model (directionality -1), a model whose file lacks some
mutants (skipped), a `sequence` column, and a model keyed on mutated_sequence."""
import importlib.util
import json
import os
import shutil
import sys

import numpy as np
import pandas as pd
from scipy.stats import spearmanr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from proteingym_b200 import synth  # noqa: E402

REF = os.environ.get("PG_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "merge_case")


def main():
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(os.path.join(OUT, "dms"))
    rng = np.random.RandomState(0)
    models = {"ModelA": {"input_score_name": "colA", "location": "A", "directionality": 1, "key": "mutant", "model_type": "x"},
              "ModelNeg": {"input_score_name": "score", "location": "neg", "directionality": -1, "key": "mutant", "model_type": "x"},
              "ModelSeq": {"input_score_name": "avg_score", "location": "deep/seq", "directionality": 1, "key": "mutated_sequence", "model_type": "x"},
              "ModelShort": {"input_score_name": "s", "location": "short", "directionality": 1, "key": "mutant", "model_type": "x"}}
    with open(os.path.join(OUT, "config.json"), "w") as fh:
        json.dump({"model_list_zero_shot_substitutions_DMS": models}, fh, indent=1)
    rows = []
    for k, L in enumerate((40, 75, 23)):
        seq = synth.random_protein(L, 70 + k)
        muts = synth.sample_mutants(seq, 60, seed=k, multi_frac=0.2)
        dms = pd.DataFrame({"mutant": muts, "mutated_sequence": [synth.apply_mutant(seq, m) for m in muts],
                            "DMS_score": rng.randn(len(muts)), "DMS_score_bin": rng.randint(0, 2, len(muts))})
        dms.to_csv(os.path.join(OUT, "dms", f"assay{k}.csv"), index=False)
        rows.append({"DMS_id": f"ASSAY{k}", "DMS_filename": f"assay{k}.csv", "target_seq": seq, "DMS_total_number_mutants": len(muts) + (k == 2)})
        for loc in ("A", "neg", "deep/seq", "short"):
            os.makedirs(os.path.join(OUT, "scores", loc), exist_ok=True)
        a = pd.DataFrame({"mutant": muts, "colA": rng.randn(len(muts)), "other": 1.0})
        a = pd.concat([a, a.iloc[:5].assign(colA=lambda d: d["colA"] + 1.0), a.iloc[5:8]], ignore_index=True)  # duplicates: mean per key
        a.to_csv(os.path.join(OUT, "scores", "A", f"ASSAY{k}.csv"), index=False)
        pd.DataFrame({"mutant": muts[::-1], "score": rng.randn(len(muts))}).to_csv(os.path.join(OUT, "scores", "neg", f"ASSAY{k}.csv"), index=False)
        pd.DataFrame({"sequence": list(dms["mutated_sequence"]), "avg_score": rng.randn(len(muts))}).assign(
            mutated_sequence="ignored").to_csv(os.path.join(OUT, "scores", "deep/seq", f"ASSAY{k}.csv"), index=False)
        short = muts[:-7] if k == 1 else muts
        pd.DataFrame({"mutant": short, "s": rng.randn(len(short))}).to_csv(os.path.join(OUT, "scores", "short", f"ASSAY{k}.csv"), index=False)
    pd.DataFrame(rows).to_csv(os.path.join(OUT, "mapping.csv"), index=False)
    spec = importlib.util.spec_from_file_location("pg_ref_merge", os.path.join(REF, "proteingym", "merge.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    argv = sys.argv
    sys.argv = ["merge.py", "--DMS_assays_location", os.path.join(OUT, "dms"), "--model_scores_location", os.path.join(OUT, "scores"),
                "--merged_scores_dir", "reference_merged", "--DMS_reference_file", os.path.join(OUT, "mapping.csv"),
                "--config_file", os.path.join(OUT, "config.json")]
    try:
        mod.main()
    finally:
        sys.argv = argv
    sp = {}
    for r in rows:
        m = pd.read_csv(os.path.join(OUT, "scores", "reference_merged", r["DMS_id"] + ".csv"))
        sp[r["DMS_id"]] = {c: float(spearmanr(m["DMS_score"], m[c])[0]) for c in models if c in m}
    with open(os.path.join(OUT, "reference_spearman.json"), "w") as fh:
        json.dump(sp, fh, indent=1)
    print(sp)


if __name__ == "__main__":
    main()
