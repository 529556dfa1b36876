"""ORACLE SUPPORT — TEST INFRASTRUCTURE ONLY. Build-container script (needs /root/reference).

Runs the UNMODIFIED reference (vendored fair-esm modules + compute_fitness.main through oracle/ref_shims.py) on seeded
synthetic checkpoints and writes small golden fixtures under tests/golden/. The checkpoints themselves are NOT stored:
``proteingym_b200.synth.make_esm_state(arch, seed)`` regenerates them bit-identically (torch CPU generator).

  python oracle/gen_golden.py [tiny] [window] [blat650m] [esm2_3b]
"""
from __future__ import annotations

import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402
import torch  # noqa: E402

from oracle import ref_shims  # noqa: E402
from proteingym_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
BLAT = None


def blat_sequence():
    df = pd.read_csv(os.path.join(ref_shims.REF, "reference_files", "DMS_substitutions.csv"))
    row = df[df["DMS_id"] == "BLAT_ECOLX_Stiffler_2015"].iloc[0]
    return row["target_seq"].upper()


def reference_table(mod, ckpt_path, seq, positions=None):
    """token_probs [L+2, 33] by the reference's own loop body (compute_fitness.py:486-504), run on the reference model.
    ``positions`` (token indices) restricts the loop to those rows (true-size cases: one forward per row is minutes of CPU);
    the result then holds only those rows, in that order."""
    model, alphabet = mod.pretrained.load_model_and_alphabet(ckpt_path)
    model.eval()
    _, _, toks = alphabet.get_batch_converter()([("protein1", seq)])
    rows = []
    with torch.no_grad():
        for i in (range(toks.size(1)) if positions is None else positions):
            t = toks.clone()
            t[0, i] = alphabet.mask_idx
            if toks.size(1) > 1024:
                start, end = mod.get_optimal_window(mutation_position_relative=i, seq_len_wo_special=len(seq) + 2, model_window=1024)
                t = t[:, start:end]
            else:
                start = 0
            rows.append(torch.log_softmax(model(t)["logits"], dim=-1)[:, i - start])
    return torch.cat(rows, 0).numpy()


def run_case(name, arch, seed, seq, mutants, ckpt_name, model_type, with_table=True, extra_ckpt=None, table_positions=None):
    mod = ref_shims.install()
    tmp = tempfile.mkdtemp(prefix="pg_gold_")
    try:
        ckpts = [os.path.join(tmp, ckpt_name)]
        synth.write_esm_checkpoint(ckpts[0], arch, seed=seed)
        if extra_ckpt:
            ckpts.append(os.path.join(tmp, extra_ckpt[0]))
            synth.write_esm_checkpoint(ckpts[1], arch, seed=extra_ckpt[1])
        os.makedirs(os.path.join(tmp, "dms"))
        synth.write_dms_csv(os.path.join(tmp, "dms", f"{name}.csv"), seq, mutants, seed=0)
        synth.write_mapping_csv(os.path.join(tmp, "map.csv"), [("OTHER_ASSAY", "other.csv", "MKV"), (name, f"{name}.csv", seq)])
        t0 = time.time()
        ref_shims.run_reference_cli(["--model-location", *ckpts, "--model_type", model_type, "--dms_index", "1",
                                     "--dms_mapping", os.path.join(tmp, "map.csv"), "--dms-input", os.path.join(tmp, "dms"),
                                     "--dms-output", os.path.join(tmp, "out"), "--scoring-strategy", "masked-marginals",
                                     "--scoring-window", "optimal", "--nogpu"])
        dt = time.time() - t0
        shutil.copy(os.path.join(tmp, "out", f"{name}.csv"), os.path.join(GOLD, f"{name}_reference_output.csv"))
        meta = {"name": name, "arch": vars(arch) if not hasattr(arch, "__dataclass_fields__") else arch.__dict__, "seed": seed,
                "sequence": seq, "ckpt_names": [os.path.basename(c) for c in ckpts], "extra_seed": extra_ckpt[1] if extra_ckpt else None,
                "model_type": model_type, "reference_cli_seconds": dt, "torch": torch.__version__,
                "threads": torch.get_num_threads()}
        if table_positions is not None:
            meta["table_positions"] = [int(i) for i in table_positions]
        if with_table:
            tab = reference_table(mod, ckpts[0], seq, table_positions)
            np.save(os.path.join(GOLD, f"{name}_reference_table.npy"), tab.astype(np.float32))
        with open(os.path.join(GOLD, f"{name}_meta.json"), "w") as fh:
            json.dump(meta, fh, indent=1)
        print(f"[gen_golden] {name}: reference CLI {dt:.1f}s, {len(mutants)} mutants", flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    which = set(sys.argv[1:]) or {"tiny", "window", "blat650m"}
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    if "tiny" in which:
        seq = blat_sequence()[:90]
        for name, arch, ck, mt, extra in (
                ("tiny_esm1v", synth.EsmArch("esm1v", 3, 128, 2, 256), "esm1v_tiny_1.pt", "ESM1v", ("esm1v_tiny_2.pt", 8)),
                ("tiny_esm1b", synth.EsmArch("esm1v", 2, 128, 2, 256, emb_layer_norm_before=True), "esm1b_tiny.pt", "ESM1b", None),
                ("tiny_esm2", synth.EsmArch("esm2", 3, 128, 2, 512), "esm2_tiny.pt", "ESM2", None)):
            muts = synth.sample_mutants(seq, 300, seed=4, multi_frac=0.25)
            run_case(name, arch, 7, seq, muts, ck, mt, extra_ckpt=extra)
    if "window" in which:
        seq = synth.random_protein(1100, seed=21)  # L+2 = 1102 > 1024 -> per-position optimal windows
        muts = synth.sample_mutants(seq, 400, seed=5, multi_frac=0.2)
        run_case("window_esm1v", synth.EsmArch("esm1v", 2, 64, 1, 128), 9, seq, muts, "esm1v_win.pt", "ESM1v")
        run_case("window_esm2", synth.EsmArch("esm2", 2, 64, 1, 256), 9, seq, muts, "esm2_win.pt", "ESM2")
    if "blat650m" in which:
        # BASELINE.json config 1: BLAT_ECOLX_Stiffler_2015, all 19 substitutions at positions 24..286, ESM-1v 650M arch
        seq = blat_sequence()
        muts = synth.all_single_mutants(seq, first=24, last=286)
        run_case("blat_esm1v_650m", synth.ESM1V_650M, 0, seq, muts, "esm1v_t33_650M_UR90S_1.pt", "ESM1v", with_table=True)
    if "esm2_3b" in which:
        # BASELINE.json config 3 architecture at true size (36 x 2560, 40 heads, ffn 10240, rotary): 256-residue protein, 300 mutants
        # of which ~70 % are 2-5-site multi-mutants (errors of independently masked sites add up), table rows at 40 positions
        seq = synth.random_protein(256, seed=31)
        muts = synth.sample_mutants(seq, 300, seed=6, multi_frac=0.7)
        run_case("esm2_3b_multi", synth.ESM2_3B, 2, seq, muts, "esm2_t36_3B_UR50D.pt", "ESM2", with_table=True,
                 table_positions=list(range(3, 257, 6))[:40] + [1, 256])


if __name__ == "__main__":
    main()
