"""ORACLE SUPPORT — TEST INFRASTRUCTURE ONLY. Build-container script (needs /root/reference).

Runs the UNMODIFIED reference MSA Transformer path (compute_fitness.py::main --model_type MSA_transformer, the vendored
``MSATransformer`` / ``RowSelfAttention`` / ``ColumnSelfAttention`` modules, ``utils.msa_utils.MSA_processing`` with its numba weights)
through oracle/ref_shims.py on seeded synthetic checkpoints and alignments, and writes golden fixtures under
tests/golden/msa_transformer_<case>/. Checkpoints are not stored (``synth.make_msa_state(arch, seed)`` regenerates them); the
alignment, the DMS table, the reference's sequence weights, its output CSV and rows of its masked-marginal table are.

  python oracle/gen_golden_msa_transformer.py [tiny] [weights] [batched] [window] [msa1b]
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


def reference_table(mod, ckpt, rows, positions):
    """Rows of token_probs by the reference's loop body (compute_fitness.py:383-399) on the reference model, given the sampled rows."""
    model, alphabet = mod.pretrained.load_model_and_alphabet(ckpt)
    model.eval()
    _, _, toks = alphabet.get_batch_converter()([rows])
    L = len(rows[0][1])
    out = []
    with torch.no_grad():
        for i in positions:
            t = toks.clone()
            t[0, 0, i] = alphabet.mask_idx
            start = 0
            if toks.size(-1) > 1024:
                start, end = mod.get_optimal_window(mutation_position_relative=i, seq_len_wo_special=L + 2, model_window=1024)
                t = t[:, :, start:end]
            out.append(torch.log_softmax(model(t)["logits"], dim=-1)[:, 0, i - start])
    return torch.cat(out, 0).numpy()


def run_case(name, arch, seed, target, msa_start, n_rows, n_mut, cli, strategy, seeds, table_positions, insert_cols=(), qk_gain=2.0):
    """``target``: the full target sequence; the alignment covers target[msa_start-1:] (MSA_start / MSA_end in the mapping file)."""
    mod = ref_shims.install()
    out_dir = os.path.join(GOLD, f"msa_transformer_{name}")
    os.makedirs(out_dir, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="pg_gold_msa_")
    try:
        ckpt = os.path.join(tmp, "msa_synth.pt")
        synth.write_msa_checkpoint(ckpt, arch, seed=seed, state=synth.make_msa_state(arch, seed, qk_gain=qk_gain))
        covered = target[msa_start - 1:]
        rows = synth.random_alignment(covered, n_rows, seed=seed + 1, insert_cols=insert_cols)
        os.makedirs(os.path.join(tmp, "msa"))
        os.makedirs(os.path.join(tmp, "weights"))
        synth.write_a2m(os.path.join(tmp, "msa", f"{name}.a2m"), dict(rows))
        muts = synth.sample_mutants(covered, n_mut, seed=seed + 2, multi_frac=0.3, offset=msa_start)
        os.makedirs(os.path.join(tmp, "dms"))
        score = np.random.RandomState(0).randn(len(muts))
        pd.DataFrame({"mutant": muts, "DMS_score": score, "DMS_score_bin": (score > 0).astype(int)}).to_csv(
            os.path.join(tmp, "dms", f"{name}.csv"), index=False)
        pd.DataFrame({"DMS_id": ["OTHER", name], "DMS_filename": ["other.csv", f"{name}.csv"], "target_seq": ["MKV", target],
                      "MSA_filename": ["other.a2m", f"{name}.a2m"], "MSA_start": [1, msa_start], "MSA_end": [3, len(target)],
                      "weight_file_name": ["other.npy", f"{name}.npy"]}).to_csv(os.path.join(tmp, "map.csv"), index=False)
        t0 = time.time()
        argv = ["--model-location", ckpt, "--model_type", "MSA_transformer", "--dms_index", "1", "--dms_mapping", os.path.join(tmp, "map.csv"),
                "--dms-input", os.path.join(tmp, "dms"), "--dms-output", os.path.join(tmp, "out"), "--scoring-strategy", "masked-marginals",
                "--scoring-window", "optimal", "--msa-path", os.path.join(tmp, "msa"), "--msa-weights-folder", os.path.join(tmp, "weights"),
                "--msa-sampling-strategy", strategy, "--seeds", *[str(s) for s in seeds], "--nogpu", *cli]
        ref_shims.run_reference_cli(argv)
        dt = time.time() - t0
        shutil.copy(os.path.join(tmp, "out", f"{name}.csv"), os.path.join(out_dir, "reference_output.csv"))
        shutil.copy(os.path.join(tmp, "msa", f"{name}.a2m"), os.path.join(out_dir, "alignment.a2m"))
        shutil.copy(os.path.join(tmp, "dms", f"{name}.csv"), os.path.join(out_dir, "dms.csv"))
        wfile = os.path.join(tmp, "weights", f"{name}.npy")
        if os.path.exists(wfile):
            shutil.copy(wfile, os.path.join(out_dir, "reference_weights.npy"))
        # the rows the reference sampled for the first seed, and rows of its table on them
        nseq = int(cli[cli.index("--msa-samples") + 1]) if "--msa-samples" in cli else 400
        processed = mod.process_msa(filename=os.path.join(tmp, "msa", f"{name}.a2m"), weight_filename=wfile, filter_msa=False,
                                    path_to_hhfilter="") if strategy == "sequence-reweighting" else None
        sampled = mod.sample_msa(sampling_strategy=strategy, filename=os.path.join(tmp, "msa", f"{name}.a2m"), nseq=nseq,
                                 weight_filename=wfile, processed_msa=processed, random_seed=seeds[0])
        with open(os.path.join(out_dir, "sampled_rows_seed%d.json" % seeds[0]), "w") as fh:
            json.dump(sampled, fh)
        tab = reference_table(mod, ckpt, sampled, table_positions)
        np.save(os.path.join(out_dir, "reference_table.npy"), tab.astype(np.float32))
        meta = {"name": name, "arch": arch.__dict__, "seed": seed, "qk_gain": qk_gain, "target_seq": target, "MSA_start": msa_start, "MSA_end": len(target),
                "n_rows_file": n_rows, "msa_samples": nseq, "strategy": strategy, "seeds": list(seeds), "column": "msa_synth",
                "table_positions": [int(i) for i in table_positions], "reference_cli_seconds": dt, "torch": torch.__version__}
        with open(os.path.join(out_dir, "meta.json"), "w") as fh:
            json.dump(meta, fh, indent=1)
        print(f"[gen_golden_msa] {name}: reference CLI {dt:.1f}s, {len(muts)} mutants, {len(sampled)} sampled rows", flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    which = set(sys.argv[1:]) or {"tiny", "weights", "batched", "window"}
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    if "tiny" in which:  # first rows of the file, one seed, '.' columns kept as tokens
        t = synth.random_protein(40, seed=51)
        run_case("tiny", synth.MsaArch(2, 128, 2, 256), 3, t, 1, 12, 120, ["--msa-samples", "8"], "first_x_rows", [1], list(range(41)),
                 insert_cols=(7, 8))
    if "weights" in which:  # the launcher's strategy: weighted sampling with replacement, two seeds, alignment on a sub-range, 1-wide
        t = synth.random_protein(70, seed=52)  # msa_position_embedding as in the first release
        run_case("weights", synth.MsaArch(2, 128, 2, 256, msa_pos_dim=1), 5, t, 11, 40, 150, ["--msa-samples", "10"], "sequence-reweighting",
                 [1, 2], list(range(0, 61, 3)))
    if "batched" in which:  # R * C > 2^14: the reference's row-chunked / column-chunked attention paths (axial_attention.py:82-113,226-252)
        t = synth.random_protein(450, seed=53)
        run_case("batched", synth.MsaArch(2, 128, 2, 256), 7, t, 1, 44, 200, ["--msa-samples", "40"], "random", [3], list(range(0, 451, 25)))
    if "window" in which:  # more than 1024 columns: per-position optimal windows
        t = synth.random_protein(1100, seed=54)
        run_case("window", synth.MsaArch(1, 64, 1, 128), 9, t, 1, 6, 200, ["--msa-samples", "4"], "first_x_rows", [1],
                 [0, 1, 300, 511, 512, 513, 600, 700, 1000, 1099, 1100])
    if "msa1b" in which:  # true MSA-1b size (12 x 768, 12 heads, ffn 3072): 129 columns + BOS crosses the 128-row tile boundary.
        # q/k gain 1 instead of the 2 used elsewhere: tied-attention logits are sums over (correlated) alignment rows, and at gain 2 the
        # 12-layer synthetic model is ill-conditioned — the reference's own fp32 log-probs then sit 2.3e-4 from an fp64 evaluation
        # (7e-6 at gain 1, like a trained model's), which says nothing about any implementation's arithmetic.
        t = synth.random_protein(129, seed=55)
        run_case("msa1b", synth.MSA_1B, 11, t, 1, 60, 300, ["--msa-samples", "48"], "sequence-reweighting", [1], list(range(0, 130, 8)) + [129],
                 qk_gain=1.0)


if __name__ == "__main__":
    main()
