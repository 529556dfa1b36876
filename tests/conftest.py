import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device; run with `-m gpu` on the GPU box")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """-> dict(meta, arch, state (seeded, regenerated), seq, df (reference CLI output), table (reference token_probs))."""
    import numpy as np
    import pandas as pd
    from proteingym_b200 import synth
    with open(os.path.join(GOLDEN, f"{name}_meta.json")) as fh:
        meta = json.load(fh)
    a = meta["arch"]
    arch = synth.EsmArch(a["kind"], a["layers"], a["embed_dim"], a["heads"], a["ffn_dim"], a["token_dropout"],
                         a["emb_layer_norm_before"], a["max_positions"], a["vocab"])
    df = pd.read_csv(os.path.join(GOLDEN, f"{name}_reference_output.csv"))
    tpath = os.path.join(GOLDEN, f"{name}_reference_table.npy")
    table = np.load(tpath) if os.path.exists(tpath) else None
    return dict(meta=meta, arch=arch, seq=meta["sequence"], df=df, table=table,
                state=lambda seed=meta["seed"]: synth.make_esm_state(arch, seed=seed))


GOLDEN_SMALL = ["tiny_esm1v", "tiny_esm1b", "tiny_esm2"]
GOLDEN_WINDOW = ["window_esm1v", "window_esm2"]
