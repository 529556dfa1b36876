"""CPU: proteingym_b200.merge_scores against the UNMODIFIED reference merge step and its Spearman line
(tests/golden/merge_case/, written by oracle/gen_golden_merge.py from proteingym/merge.py::main and scipy.stats.spearmanr)."""
import json
import os
import shutil

import numpy as np
import pandas as pd

from conftest import GOLDEN
from proteingym_b200 import merge_scores as MS

CASE = os.path.join(GOLDEN, "merge_case")


def test_merge_and_spearman_match_reference(tmp_path):
    work = tmp_path / "scores"
    shutil.copytree(os.path.join(CASE, "scores"), work, ignore=shutil.ignore_patterns("reference_merged"))
    rc = MS.main(["--DMS_assays_location", os.path.join(CASE, "dms"), "--model_scores_location", str(work), "--DMS_reference_file",
                  os.path.join(CASE, "mapping.csv"), "--config_file", os.path.join(CASE, "config.json")])
    assert rc == 0
    ref_sp = json.load(open(os.path.join(CASE, "reference_spearman.json")))
    met = pd.read_csv(work / "merged_scores" / "_metrics.csv")
    for aid in ("ASSAY0", "ASSAY1", "ASSAY2"):
        got = pd.read_csv(work / "merged_scores" / f"{aid}.csv")
        ref = pd.read_csv(os.path.join(CASE, "scores", "reference_merged", f"{aid}.csv"))
        assert list(got.columns) == list(ref.columns)  # same models merged (ModelShort skipped on ASSAY1), same order
        for c in ref.columns:
            if ref[c].dtype.kind == "f":
                assert np.allclose(got[c].to_numpy(), ref[c].to_numpy(), rtol=0, atol=1e-12, equal_nan=True), (aid, c)
            else:
                assert got[c].equals(ref[c]), (aid, c)
        for model, v in ref_sp[aid].items():
            mine = met[(met.DMS_id == aid) & (met.model == model)]["spearman_vs_DMS"].values[0]
            assert abs(mine - v) < 1e-12, (aid, model)


def test_spearman_ties_and_parity_columns(tmp_path):
    from scipy.stats import spearmanr
    rng = np.random.RandomState(3)
    x = np.round(rng.randn(500), 1)  # many ties
    y = np.round(x + rng.randn(500), 1)
    assert abs(MS.spearman(x, y) - spearmanr(x, y)[0]) < 1e-12
    assert np.isnan(MS.spearman([1.0, np.nan, 2.0], [1.0, 2.0, 3.0]))
    # parity mode: the reference's score files stand in for "the reference implementation's outputs"
    work = tmp_path / "scores"
    shutil.copytree(os.path.join(CASE, "scores"), work, ignore=shutil.ignore_patterns("reference_merged"))
    noisy = tmp_path / "ours"
    shutil.copytree(work, noisy)
    f = noisy / "A" / "ASSAY0.csv"
    df = pd.read_csv(f)
    df["colA"] += 4e-4
    df.to_csv(f, index=False)
    args = ["--DMS_assays_location", os.path.join(CASE, "dms"), "--model_scores_location", str(noisy), "--DMS_reference_file",
            os.path.join(CASE, "mapping.csv"), "--config_file", os.path.join(CASE, "config.json"), "--reference_scores_location", str(work)]
    assert MS.main(args + ["--min_parity_spearman", "0.999", "--max_parity_abs", "1e-3"]) == 0
    met = pd.read_csv(noisy / "merged_scores" / "_metrics.csv")
    row = met[(met.DMS_id == "ASSAY0") & (met.model == "ModelA")].iloc[0]
    assert abs(row["parity_max_abs"] - 4e-4) < 1e-9 and row["parity_spearman"] > 0.999999
    assert MS.main(args + ["--max_parity_abs", "1e-4"]) == 1  # the bar is enforced
