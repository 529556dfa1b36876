"""Seeded inputs of the TranceptEVE golden cases (tests/golden/trancepteve_<case>/), shared by the tests and by the script that
produced the reference outputs (oracle/gen_golden_trancepteve.py). No scoring arithmetic here."""
from __future__ import annotations

import json
import os
import shutil

import pandas as pd

from proteingym_b200 import synth

CASES = {
    # name: arch, target length, MSA range (0-based start, end), retrieval type, thresholds, recalibration, EVE seeds, gappy columns
    "trancepteve_subs": dict(arch=(2, 256, 4, 512, 1024), L=70, msa=(8, 62), n_msa=400, kind="TranceptEVE", seq_thr=0.5, col_thr=1.0,
                             msa_recal=False, eve_recal=True, eve_seeds=[0], n_samples=4, gappy=[], n_mut=100),
    "trancepteve_nonfocus": dict(arch=(2, 256, 4, 512, 1024), L=70, msa=(8, 62), n_msa=150, kind="TranceptEVE", seq_thr=0.5, col_thr=0.3,
                                 msa_recal=True, eve_recal=True, eve_seeds=[0, 1], n_samples=3, gappy=[5, 17, 18, 40], n_mut=100),
    "trancepteve_msa_only": dict(arch=(2, 256, 4, 512, 1024), L=70, msa=(0, 70), n_msa=40, kind="Tranception", seq_thr=0.5, col_thr=1.0,
                                 msa_recal=False, eve_recal=False, eve_seeds=[], n_samples=0, gappy=[], n_mut=60),
    "trancepteve_long": dict(arch=(1, 256, 4, 256, 64), L=150, msa=(20, 140), n_msa=120, kind="TranceptEVE", seq_thr=0.5, col_thr=1.0,
                             msa_recal=False, eve_recal=True, eve_seeds=[0], n_samples=2, gappy=[], n_mut=50),
    # BASELINE config 5 at TRUE SIZE: Tranception-L (36 x 1280, 20 heads, ffn 5120) under the real TrancepteveLMHeadModel, both
    # recalibrations on. GPU tests only (true_size: the CPU host-logic tests run the small cases).
    "trancepteve_L_true": dict(arch=(36, 1280, 20, 5120, 1024), L=120, msa=(10, 110), n_msa=200, kind="TranceptEVE", seq_thr=0.5,
                               col_thr=0.3, msa_recal=True, eve_recal=True, eve_seeds=[0], n_samples=3, gappy=[7, 33], n_mut=48,
                               true_size=True),
}
SMALL_CASES = {n: c for n, c in CASES.items() if not c.get("true_size")}


def make_inputs(case: dict, work: str, weights_file: str | None = None):
    """Regenerate the seeded inputs of a case into ``work`` (also used by the tests). Returns paths + objects."""
    a = case["arch"]
    arch = synth.TranceptionArch(a[0], a[1], a[2], a[3], n_ctx=a[4])
    seq = synth.random_protein(case["L"], 41)
    s, e = case["msa"]
    msa = synth.synthetic_msa(seq[s:e], case["n_msa"], seed=9, gappy_cols=case["gappy"])
    msa_file = os.path.join(work, "TARGET_msa.a2m")
    synth.write_a2m(msa_file, msa)
    wfile = os.path.join(work, "TARGET_weights.npy")
    if weights_file is not None:
        shutil.copy(weights_file, wfile)
    eve_dir = os.path.join(work, "eve")
    os.makedirs(eve_dir, exist_ok=True)
    params_file = os.path.join(work, "eve_params.json")
    with open(params_file, "w") as fh:
        json.dump(synth.EVE_TINY_PARAMS, fh)
    muts = synth.sample_mutants(seq, case["n_mut"], seed=3, multi_frac=0.3)
    dms = pd.DataFrame({"mutant": muts, "DMS_score": 0.0})
    dms["mutated_sequence"] = [synth.apply_mutant(seq, m) for m in muts]
    return dict(arch=arch, seq=seq, msa_file=msa_file, weights_file=wfile, eve_dir=eve_dir, params_file=params_file, dms=dms)
