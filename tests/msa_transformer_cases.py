"""Loader of the MSA Transformer golden cases (tests/golden/msa_transformer_<case>/, written by oracle/gen_golden_msa_transformer.py
from the unmodified reference)."""
import json
import os

import numpy as np
import pandas as pd

from conftest import GOLDEN
from proteingym_b200 import synth

SMALL_CASES = ["tiny", "weights", "batched", "window"]


def case_dir(name):
    return os.path.join(GOLDEN, f"msa_transformer_{name}")


def have(name):
    return os.path.exists(os.path.join(case_dir(name), "meta.json"))


def load_case(name):
    d = case_dir(name)
    with open(os.path.join(d, "meta.json")) as fh:
        meta = json.load(fh)
    arch = synth.MsaArch(**meta["arch"])
    with open(os.path.join(d, "sampled_rows_seed%d.json" % meta["seeds"][0])) as fh:
        rows = [(n, s) for n, s in json.load(fh)]
    wpath = os.path.join(d, "reference_weights.npy")
    return dict(meta=meta, arch=arch, dir=d, rows=rows, qk_gain=meta.get("qk_gain", 2.0), df=pd.read_csv(os.path.join(d, "reference_output.csv")),
                table=np.load(os.path.join(d, "reference_table.npy")), weights=np.load(wpath) if os.path.exists(wpath) else None,
                sequence=meta["target_seq"][meta["MSA_start"] - 1:meta["MSA_end"]])
